"""Pins the filter oracle on the reference's own known-answer tests
(watsor/test/test_filter.py:14-96) and checks the raster/summed-area restatement of the
mask predicate against exact polygon geometry on random masks (CPU only)."""
import os
from tempfile import NamedTemporaryFile

import cv2
import numpy as np
import pytest
from PIL import Image, ImageDraw

from oracle.filters import (AreaOracle, ConfidenceOracle, Det, MaskOracle, TrackOracle, apply_predicates,
                            find_contours, get_alpha_channel, rect_intersects_polygon)
from tests.conftest import GOLDEN_DIR


def test_confidence_known_answers():          # test_filter.py:14-22
    f = ConfidenceOracle({'detect': [{'person': {'confidence': 50}}]})
    assert f(Det(1, 0.70))
    assert not f(Det(1, 0.40))
    assert not f(Det(2, 0.70))


def test_area_known_answers():                # test_filter.py:24-36
    f = AreaOracle({'width': 100, 'height': 100, 'detect': [{'person': {'area': 50}}]})
    assert f(Det(1, 0.70, (0, 0, 100, 50)))
    assert not f(Det(1, 0.70, (0, 0, 50, 50)))
    assert not f(Det(2, 0.70, (0, 0, 100, 50)))


def half_mask(path):
    with Image.new('RGBA', (100, 100)) as image:
        with Image.new("L", image.size) as alpha:
            ImageDraw.Draw(alpha).rectangle((50, 0, alpha.width, alpha.height), fill=255)
            image.putalpha(alpha)
        image.save(path)


def test_mask_known_answers():                # test_filter.py:38-74
    with pytest.raises(AssertionError, match="Error reading mask file"):
        MaskOracle({'width': 1, 'height': 1, 'mask': 'notafile.png'})
    tmp = NamedTemporaryFile(suffix='.png', delete=False)
    try:
        with Image.new('RGB', (10, 10)) as image:
            image.save(tmp.name)
        with pytest.raises(AssertionError, match="Mask image .+ is not of 32 bit color"):
            MaskOracle({'width': 10, 'height': 10, 'mask': tmp.name})
        half_mask(tmp.name)
        with pytest.raises(AssertionError, match="The size of mask image .+ doesn't match"):
            MaskOracle({'width': 50, 'height': 50, 'mask': tmp.name})
        f = MaskOracle({'width': 100, 'height': 100, 'mask': tmp.name, 'detect': []})
    finally:
        tmp.close()
        os.unlink(tmp.name)
    assert not f(Det(1, 0.70, (20, 20, 40, 80)))
    d = Det(1, 0.70, (20, 20, 80, 80))
    assert f(d) and d.zones[0] == 1


def test_track_known_answers():               # test_filter.py:76-96
    t = TrackOracle(sensitivity=1, history=2)
    dets, sus = t([Det(1, 0.70, (50, 50, 60, 60)), Det(1, 0.70, (10, 10, 30, 30))])
    assert sus and [(d.x_min, d.y_min, d.x_max, d.y_max) for d in dets] == [(50, 50, 60, 60), (10, 10, 30, 30)]
    dets, sus = t([Det(1, 0.70, (40, 40, 55, 55)), Det(1, 0.70, (80, 80, 90, 90))])
    assert sus and [(d.x_min, d.y_min, d.x_max, d.y_max) for d in dets] == [(40, 40, 60, 60), (80, 80, 90, 90)]


def test_lazy_predicate_chain_writes_zones_only_after_earlier_filters_pass():
    cfg = {'width': 640, 'height': 480, 'mask': os.path.join(GOLDEN_DIR, 'porch.png'),
           'detect': [{'person': {'confidence': 50, 'area': 10, 'zones': []}}]}
    filters = [ConfidenceOracle(cfg), AreaOracle(cfg), MaskOracle(cfg)]
    low = Det(1, 0.2, (0, 0, 639, 479))
    ok = Det(1, 0.9, (0, 0, 639, 479))
    kept, verdicts = apply_predicates([low, ok, Det(0, 0.9, (0, 0, 639, 479))], filters)
    assert low.zones == [0] * 10 and ok.zones[:2] == [1, 2] and kept == [ok]
    assert verdicts == [1, 1 | 2 | 4 | 8, 0]


def test_porch_mask_has_two_zones_in_reference_order():
    alpha, _ = get_alpha_channel(os.path.join(GOLDEN_DIR, 'porch.png'), 640, 480)
    contours = find_contours(alpha)
    assert [len(c) for c in contours] == [171, 322]       # SURVEY.md 8(a) notes for a12


def random_mask(rng, w, h):
    alpha = np.full((h, w), 216, np.uint8)
    for _ in range(int(rng.integers(1, 5))):
        x0, y0 = int(rng.integers(0, w - 4)), int(rng.integers(0, h - 4))
        x1, y1 = int(rng.integers(x0 + 1, w)), int(rng.integers(y0 + 1, h))
        if rng.random() < 0.5:
            cv2.rectangle(alpha, (x0, y0), (x1, y1), 255, -1)
        else:
            cv2.ellipse(alpha, ((x0 + x1) // 2, (y0 + y1) // 2), (max(1, (x1 - x0) // 2), max(1, (y1 - y0) // 2)),
                        float(rng.integers(0, 180)), 0, 360, 255, -1)
    if rng.random() < 0.5:      # a hole and an island inside it: RETR_EXTERNAL must fill both
        cv2.circle(alpha, (w // 2, h // 2), min(w, h) // 6, 216, -1)
        cv2.circle(alpha, (w // 2, h // 2), max(1, min(w, h) // 16), 255, -1)
    return alpha


@pytest.mark.parametrize('seed', range(6))
def test_raster_sat_equals_exact_polygon_intersection(seed):
    """bbox `intersects` polygon (mask.py:54)  <=>  bbox covers a pixel of the filled contour."""
    rng = np.random.default_rng(seed)
    w, h = int(rng.integers(40, 160)), int(rng.integers(40, 120))
    alpha = random_mask(rng, w, h)
    contours = [c for c in find_contours(alpha) if len(c) >= 3 and cv2.moments(c)['m00'] > 0]
    for c in contours:
        raster = np.zeros((h, w), np.uint8)
        cv2.drawContours(raster, [c], -1, 1, thickness=cv2.FILLED)
        sat = np.pad(raster.astype(np.int64).cumsum(0).cumsum(1), ((1, 0), (1, 0)))
        for i in range(400):
            x0, x1 = sorted(int(v) for v in rng.integers(0, w, 2))
            y0, y1 = sorted(int(v) for v in rng.integers(0, h, 2))
            if i % 5 == 0:
                x1 = x0
            if i % 7 == 0:
                y1 = y0
            if i % 3 == 0:
                x1, y1 = min(w - 1, x0 + int(rng.integers(0, 6))), min(h - 1, y0 + int(rng.integers(0, 6)))
            cnt = sat[y1 + 1, x1 + 1] - sat[y0, x1 + 1] - sat[y1 + 1, x0] + sat[y0, x0]
            assert (cnt > 0) == rect_intersects_polygon(x0, y0, x1, y1, c[:, 0]), (seed, x0, y0, x1, y1)
