"""Visual effects, host side (no GPU): the tables and the algorithm the CUDA kernel implements, checked against
OpenCV -- the reference's own arithmetic for this path (watsor/output/draw.py, blend.py)."""
import cv2
import numpy as np
import pytest

from oracle import effects as oracle_fx
from tests.fx_cases import random_alpha, random_rows
from tests.fx_emulation import add_weighted_lut, percent_digits, render
from watsor_b200.config.coco import COCO_CLASSES, get_coco_class
from watsor_b200.output.font import FontAtlas


@pytest.fixture(scope='module')
def atlas():
    return FontAtlas(''.join(COCO_CLASSES) + ': 0123456789%')


def test_text_metrics_equal_get_text_size(atlas):
    face = cv2.FONT_HERSHEY_DUPLEX
    for label in COCO_CLASSES:
        for pct in (0, 7, 55, 100):
            text = '%s: %d%%' % (label, pct)
            (w, h), base = cv2.getTextSize(text, face, 0.5, 1)
            assert (atlas.text_width(text), atlas.text_height, atlas.baseline) == (w, h, base), text


def test_glyph_tables_reproduce_put_text_including_the_right_border(atlas):
    rng = np.random.default_rng(0)
    face = cv2.FONT_HERSHEY_DUPLEX
    for trial in range(250):
        label = COCO_CLASSES[int(rng.integers(0, 91))]
        text = '%s: %d%%' % (label, int(rng.integers(0, 101)))
        w, h = int(rng.integers(20, 240)), int(rng.integers(43, 80))
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        org = (int(rng.integers(2, w)), int(rng.integers(16, h - 6)))      # text may run over the right border
        ref = img.copy()
        cv2.putText(ref, text, org, face, 0.5, (255, 255, 255), 1, cv2.LINE_AA)
        got = img.copy()
        atlas.draw(got, text, org)
        assert np.array_equal(ref, got), (text, w, h, org)


def test_add_weighted_is_one_float_fma_rounded_half_even():
    ramp = np.arange(256, dtype=np.uint8).reshape(1, 256, 1).repeat(3, axis=2)
    for idx in range(91):
        cls = get_coco_class(idx)
        solid = np.full(ramp.shape, cls.box_color, np.uint8)
        ref = cv2.addWeighted(ramp, cls.alpha, solid, 1 - cls.alpha, 0)
        lut = add_weighted_lut(cls.box_color, cls.alpha)
        for c in range(3):
            assert np.array_equal(ref[0, :, c], lut[c]), (idx, c)
    # and on a large (vectorised inside OpenCV) image with arbitrary colours
    rng = np.random.default_rng(1)
    img = rng.integers(0, 256, (64, 333, 3), dtype=np.uint8)
    for color in ((0, 0, 0), (255, 255, 255), (1, 128, 254), (37, 77, 13)):
        ref = cv2.addWeighted(img, 0.55, np.full(img.shape, color, np.uint8), 1 - 0.55, 0)
        lut = add_weighted_lut(color, 0.55)
        got = np.stack([lut[c][img[:, :, c]] for c in range(3)], axis=2)
        assert np.array_equal(ref, got), color


def test_percent_string_equals_python_format():
    rng = np.random.default_rng(2)
    values = [0.0, 1.0, 0.005, 0.015, 0.025, 0.125, 0.995, 0.985, 0.5, 0.49999997] + \
             [float(np.float32(v)) for v in rng.random(5000)] + [k / 200.0 for k in range(201)]
    for v in values:
        assert percent_digits(v) + '%' == '{0:.0%}'.format(v), v


@pytest.mark.parametrize('size', [(320, 240), (640, 480), (97, 61)])
def test_kernel_algorithm_equals_opencv_chain_without_mask(atlas, size):
    w, h = size
    rng = np.random.default_rng(w)
    for trial in range(4):
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        rows = random_rows(rng, w, h, n_drawn=[0, 3, 12, 40][trial])
        ref = oracle_fx.effect_chain(img, rows)
        got = render(atlas, img, rows)
        assert np.array_equal(ref, got), (size, trial, int((ref != got).sum()))


@pytest.mark.parametrize('size', [(320, 240), (200, 150)])
def test_kernel_algorithm_equals_opencv_chain_with_mask(atlas, size):
    from watsor_b200.filter.mask import find_contours
    from watsor_b200.output.effects import contour_bits
    w, h = size
    rng = np.random.default_rng(h)
    for trial in range(3):
        nz = trial + 1
        alpha = random_alpha(rng, w, h, nz)
        assert len(find_contours(alpha)) == nz
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        rows = random_rows(rng, w, h, n_drawn=10, n_zones=nz)
        ref = oracle_fx.effect_chain(img, rows, alpha)
        got = render(atlas, img, rows, alpha, contour_bits(alpha))
        assert np.array_equal(ref, got), (size, trial, int((ref != got).sum()))
