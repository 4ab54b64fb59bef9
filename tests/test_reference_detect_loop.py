"""The write loop of the reference detector (watsor/detection/tensorflow_cpu.py:74-92) run here, as it is, against
the oracle's `to_detections`.  TensorFlow is not installed, so the class is imported with an
empty stand-in `tensorflow` module and instantiated without `__init__`; only `detect()` runs, with the arrays a
`sess.run` would return supplied by the test.  Two readings of `int(np.float32 * int)`:
  * legacy promotion (numpy 1.23, the reference's pin, docker/Dockerfile.base:33): the product is a float64, i.e.
    exact -- reproduced on any numpy by handing the loop float64 copies of the float32 boxes;
  * NEP 50 (numpy >= 2): the product is rounded to float32 first; it differs from the exact reading only where that
    rounding lands on an integer -- counted here, and every such case is checked to be exactly that.
CPU only; skipped where /root/reference is absent."""
import os
import sys
import types

import numpy as np
import pytest

from oracle.ssd_graph import to_detections

REF = '/root/reference'
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason='reference tree not present')


@pytest.fixture(scope='module')
def ref():
    saved = sys.modules.get('tensorflow')
    sys.modules['tensorflow'] = types.ModuleType('tensorflow')
    sys.path.insert(0, REF)
    try:
        from watsor.detection.tensorflow_cpu import TensorFlowObjectDetector
        from watsor.stream.share import Detection
        yield types.SimpleNamespace(Detector=TensorFlowObjectDetector, Detection=Detection)
    finally:
        sys.path.remove(REF)
        if saved is None:
            sys.modules.pop('tensorflow', None)
        else:
            sys.modules['tensorflow'] = saved


def run_reference_loop(ref, boxes, classes, scores, shape, n_rows=100):
    det = object.__new__(ref.Detector)
    setattr(det, '_TensorFlowObjectDetector__detect_fn', lambda image: (boxes, classes, scores))
    rows = (ref.Detection * n_rows)()
    ms = det.detect(shape, None, rows)
    assert ms >= 0.0
    return [(r.label, r.confidence, r.bounding_box.x_min, r.bounding_box.y_min, r.bounding_box.x_max,
             r.bounding_box.y_max) for r in rows]


def random_outputs(rng, n=100):
    boxes = rng.random((n, 4)).astype(np.float32)
    boxes[::7] = np.float32(1.0)                               # clipped to the window edge
    boxes[1::7] = np.float32(0.0)
    boxes[2::11, 2:] = np.nextafter(np.float32(1.0), np.float32(0.0))
    k = int(rng.integers(0, n))
    boxes[k:] = 0.0                                             # PadOrClipBoxList padding
    scores = np.sort(rng.random(n).astype(np.float32))[::-1].copy()
    scores[k:] = 0.0
    classes = rng.integers(1, 4, n).astype(np.float32)
    classes[k:] = 1.0                                           # padded rows: 0 + 1
    return boxes, classes, scores


@pytest.mark.parametrize('shape', [(480, 640, 3), (240, 320, 3), (1080, 1920, 3), (2, 2, 3), (1, 1, 3)])
def test_oracle_equals_reference_loop_under_legacy_promotion(ref, shape):
    rng = np.random.default_rng(shape[0])
    for _ in range(40):
        boxes, classes, scores = random_outputs(rng)
        got = run_reference_loop(ref, boxes.astype(np.float64), classes, scores, shape)
        assert got == to_detections(boxes, classes, scores, shape)


def test_short_outputs_leave_the_remaining_rows_untouched(ref):
    rng = np.random.default_rng(1)
    boxes, classes, scores = random_outputs(rng, 7)
    got = run_reference_loop(ref, boxes.astype(np.float64), classes, scores, (480, 640, 3))
    assert got[:7] == to_detections(boxes, classes, scores, (480, 640, 3))
    assert all(r == (0, 0.0, 0, 0, 0, 0) for r in got[7:])     # share.py:47-50 zeros stay


def test_nep50_reading_differs_only_on_float32_rounding_to_an_integer(ref):
    if int(np.__version__.split('.')[0]) < 2:
        pytest.skip('NEP 50 promotion needs numpy >= 2')
    rng = np.random.default_rng(7)
    shape = (1080, 1920, 3)
    differing = total = 0
    for _ in range(200):
        boxes, classes, scores = random_outputs(rng)
        new = run_reference_loop(ref, boxes, classes, scores, shape)
        exact = to_detections(boxes, classes, scores, shape)
        for d, (a, b) in enumerate(zip(new, exact)):
            assert a[:2] == b[:2]
            for j, (va, vb) in enumerate(zip(a[2:], b[2:])):
                total += 1
                if va != vb:
                    differing += 1
                    col = (1, 0, 3, 2)[j]                       # x_min,y_min,x_max,y_max <- boxes[:, 1,0,3,2]
                    mx = (shape[1] - 1) if j % 2 == 0 else (shape[0] - 1)
                    p32 = np.float32(boxes[d][col]) * np.float32(mx)
                    assert va == vb + 1 and float(p32) == float(va), (d, j, boxes[d][col])
    assert differing < total * 1e-3
