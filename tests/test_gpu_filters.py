"""Filter predicates in CUDA: the reference's known-answer tests (watsor/test/test_filter.py)
through the drop-in filter classes, and random rows against the oracle."""
import os
from tempfile import NamedTemporaryFile

import numpy as np
import pytest
from PIL import Image

from oracle.filters import AreaOracle, ConfidenceOracle, Det, MaskOracle, apply_predicates
from tests.conftest import PORCH_CONFIG
from tests.test_oracle_filters import half_mask, random_mask
from watsor_b200 import _lib
from watsor_b200.filter.area import AreaFilter
from watsor_b200.filter.confidence import ConfidenceFilter
from watsor_b200.filter.mask import MaskFilter
from watsor_b200.stream.share import BoundingBox, Detection

pytestmark = pytest.mark.gpu


def test_confidence():                      # test_filter.py:14-22
    f = ConfidenceFilter({'detect': [{'person': {'confidence': 50}}]})
    assert f(Detection(label=1, confidence=0.70))
    assert not f(Detection(label=1, confidence=0.40))
    assert not f(Detection(label=2, confidence=0.70))


def test_area():                            # test_filter.py:24-36
    f = AreaFilter({'width': 100, 'height': 100, 'detect': [{'person': {'area': 50}}]})
    assert f(Detection(label=1, confidence=0.70, bounding_box=BoundingBox(0, 0, 100, 50)))
    assert not f(Detection(label=1, confidence=0.70, bounding_box=BoundingBox(0, 0, 50, 50)))
    assert not f(Detection(label=2, confidence=0.70, bounding_box=BoundingBox(0, 0, 100, 50)))


def test_mask():                            # test_filter.py:38-74
    with pytest.raises(AssertionError, match="Error reading mask file"):
        MaskFilter({'width': 1, 'height': 1, 'mask': 'notafile.png'})
    tmp = NamedTemporaryFile(suffix='.png', delete=False)
    try:
        with Image.new('RGB', (10, 10)) as image:
            image.save(tmp.name)
        with pytest.raises(AssertionError, match="Mask image .+ is not of 32 bit color"):
            MaskFilter({'width': 10, 'height': 10, 'mask': tmp.name})
        half_mask(tmp.name)
        with pytest.raises(AssertionError, match="The size of mask image .+ doesn't match"):
            MaskFilter({'width': 50, 'height': 50, 'mask': tmp.name})
        f = MaskFilter({'width': 100, 'height': 100, 'mask': tmp.name, 'detect': []})
    finally:
        tmp.close()
        os.unlink(tmp.name)
    assert not f(Detection(label=1, confidence=0.70, bounding_box=BoundingBox(20, 20, 40, 80)))
    d = Detection(label=1, confidence=0.70, bounding_box=BoundingBox(20, 20, 80, 80))
    assert f(d) and d.zones[0] == 1


def random_rows(rng, w, h, n=100):
    rows = (Detection * n)()
    dets = []
    for r in range(n):
        x0, x1 = sorted(int(v) for v in rng.integers(0, w, 2))
        y0, y1 = sorted(int(v) for v in rng.integers(0, h, 2))
        if r % 6 == 0:
            x1, y1 = x0, y0
        if r % 9 == 0:
            x0, x1 = x1, x0                      # unordered corners
        label = int(rng.integers(0, 5))
        conf = float(np.float32(rng.random()))
        if r % 11 == 0:
            conf = 0.5                           # exactly on the threshold: >= must pass
        rows[r].label, rows[r].confidence = label, conf
        rows[r].bounding_box = BoundingBox(x0, y0, x1, y1)
        dets.append(Det(label, conf, (x0, y0, x1, y1)))
    return rows, dets


def test_fused_chain_on_porch_mask_random_rows(shapes_model):
    from watsor_b200.detection.b200 import camera_tables
    from watsor_b200.engine import Engine
    oracle_filters = [ConfidenceOracle(PORCH_CONFIG), AreaOracle(PORCH_CONFIG), MaskOracle(PORCH_CONFIG)]
    rasters, table = camera_tables(PORCH_CONFIG, 640, 480)
    assert rasters.shape == (2, 480, 640)
    with Engine(shapes_model.to_blob(), device=0, max_batch=1) as e:
        e.set_camera(3, 640, 480, rasters, table)
        for seed in range(5):
            rows, dets = random_rows(np.random.default_rng(seed), 640, 480)
            verd = e.filter_rows(3, rows)
            kept, want = apply_predicates(dets, oracle_filters)
            got = [int(v) & 15 for v in verd]
            assert got == want
            assert [bool(v & _lib.WB_V_PASS) for v in verd] == [w == 15 for w in want]
            assert [list(rows[r].zones) for r in range(100)] == [d.zones for d in dets]


@pytest.mark.parametrize('seed', range(4))
def test_mask_filter_random_masks(seed):
    import cv2
    rng = np.random.default_rng(100 + seed)
    w, h = int(rng.integers(60, 200)), int(rng.integers(60, 160))
    alpha = random_mask(rng, w, h)
    rgba = np.zeros((h, w, 4), np.uint8)
    rgba[..., 3] = alpha
    tmp = NamedTemporaryFile(suffix='.png', delete=False)
    try:
        cv2.imwrite(tmp.name, rgba)
        cfg = {'width': w, 'height': h, 'mask': tmp.name,
               'detect': [{'person': {'confidence': 0, 'area': 0, 'zones': [1]}}]}
        try:
            oracle = MaskOracle(cfg)
        except (AssertionError, ZeroDivisionError):
            pytest.skip('degenerate random mask')
        f = MaskFilter(cfg)
    finally:
        tmp.close()
        os.unlink(tmp.name)
    rows, dets = random_rows(rng, w, h, 200)
    for r in range(200):
        want = oracle(dets[r])
        assert f(rows[r]) == want and list(rows[r].zones) == dets[r].zones, (r, dets[r].key())


def test_track_filter_known_answers():      # test_filter.py:76-96
    from watsor_b200.filter.track import TrackFilter
    t = TrackFilter(sensitivity=1, history=2)
    dets, sus = t([Detection(label=1, confidence=0.70, bounding_box=BoundingBox(50, 50, 60, 60)),
                   Detection(label=1, confidence=0.70, bounding_box=BoundingBox(10, 10, 30, 30))])
    box = lambda d: [d.bounding_box.x_min, d.bounding_box.y_min, d.bounding_box.x_max, d.bounding_box.y_max]
    assert sus and [box(d) for d in dets] == [[50, 50, 60, 60], [10, 10, 30, 30]]
    dets, sus = t([Detection(label=1, confidence=0.70, bounding_box=BoundingBox(40, 40, 55, 55)),
                   Detection(label=1, confidence=0.70, bounding_box=BoundingBox(80, 80, 90, 90))])
    assert sus and [box(d) for d in dets] == [[40, 40, 60, 60], [80, 80, 90, 90]]


def test_track_filter_fused_predicates_and_sieve_match_oracle():
    """TrackFilter([Confidence, Area, Mask]) + sieve write-back vs the oracle's restatement of
    track.py / sieve.py on random rows over several frames (history and sensitivity exercised)."""
    from oracle.filters import TrackOracle
    from watsor_b200.filter.sieve import sieve_frame
    from watsor_b200.filter.track import TrackFilter
    ours = TrackFilter([ConfidenceFilter(PORCH_CONFIG), AreaFilter(PORCH_CONFIG), MaskFilter(PORCH_CONFIG)],
                       sensitivity=2, history=3)
    oracle = TrackOracle([ConfidenceOracle(PORCH_CONFIG), AreaOracle(PORCH_CONFIG), MaskOracle(PORCH_CONFIG)],
                         sensitivity=2, history=3)
    rng = np.random.default_rng(7)
    anchors = [(int(rng.integers(0, 500)), int(rng.integers(0, 380)), int(rng.integers(1, 4))) for _ in range(6)]
    for frame in range(8):
        rows, dets = random_rows(rng, 640, 480, 100)
        for i, (x, y, lab) in enumerate(anchors):                     # persistent objects that jitter a little
            dx, dy = int(rng.integers(-3, 4)), int(rng.integers(-3, 4))
            rows[i].label, rows[i].confidence = lab, 0.9
            rows[i].bounding_box = BoundingBox(x + dx, y + dy, x + 120 + dx, y + 90 + dy)
            dets[i] = Det(lab, 0.9, (x + dx, y + dy, x + 120 + dx, y + 90 + dy))
        sus = sieve_frame(rows, [ours])
        want, want_sus = oracle(dets)
        assert sus == want_sus
        got = [(rows[r].label, rows[r].confidence, rows[r].bounding_box.x_min, rows[r].bounding_box.y_min,
                rows[r].bounding_box.x_max, rows[r].bounding_box.y_max, list(rows[r].zones)) for r in range(100)]
        exp = [(d.label, d.confidence, d.x_min, d.y_min, d.x_max, d.y_max, d.zones) for d in want]
        assert got[:len(exp)] == exp and all(g == (0, 0.0, 0, 0, 0, 0, [0] * 10) for g in got[len(exp):])
