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
