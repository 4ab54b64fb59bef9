"""SSD-MobileNet-v2 (the model BASELINE.json's 640x480 configs name): architecture descriptor with
seeded synthetic weights -- no v2 weights exist offline -- checked GPU-vs-oracle: inverted residual
blocks (1x1 expand, depthwise, linear projection, residual add), 24-channel tensors (K not a multiple of
16), 576/1280-wide taps."""
import numpy as np
import pytest

from tests.artist import artist_frame
from tests.gpu_util import compare_rows, rows_to_tuples
from watsor_b200.detection.b200 import B200ObjectDetector
from watsor_b200.engine import Engine
from watsor_b200.model import OP_HEAD, synthetic_ssd_mobilenet_v2
from watsor_b200.stream.share import Detection

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def v2():
    from oracle.ssd_model import SsdModelOracle
    m = synthetic_ssd_mobilenet_v2(num_classes=3, seed=2, score_thr=0.3)
    return m, SsdModelOracle(m), SsdModelOracle(m, dtype=np.float64)


@pytest.mark.parametrize('precision', [0, 2], ids=['fp32-cuda-cores', 'fp32-3xtf32-tcgen05'])
def test_v2_layer_by_layer(v2, precision):
    m, oracle, _ = v2
    pre = oracle.preprocess(artist_frame(640, 480, 1, 0))
    enc, lg, memo = oracle.raw_heads(pre, return_memo=True)
    with Engine(m.to_blob(), device=0, max_batch=2, precision=precision) as e:
        for li, layer in enumerate(m.layers):
            if layer.op == OP_HEAD:
                continue
            want = oracle.feature(memo, li)
            _, _, got = e.backbone(pre[None], stop_layer=li, layer_shape=want.shape)
            err, scale = np.abs(got[0] - want).max(), max(1.0, float(np.abs(want).max()))
            # random weights + residual adds amplify rounding differences ~5x more than the trained v1 net
            assert err <= (1e-4 if precision == 0 else 4e-4) * scale, (li, layer.name, err, scale)
        genc, glg, _ = e.backbone(pre[None])
    assert np.abs(genc[0] - enc).max() <= 1e-3 and np.abs(glg[0] - lg).max() <= 3e-3


@pytest.mark.parametrize('precision', [0, 2], ids=['fp32-cuda-cores', 'fp32-3xtf32-tcgen05'])
def test_v2_detect_rows(v2, precision):
    from oracle.ssd_graph import to_detections
    m, oracle, oracle64 = v2
    flips = 0
    with B200ObjectDetector(None, device=0, max_batch=4, precision=precision, model_blob=m.to_blob()) as det:
        for frame in range(3):
            img = artist_frame(640, 480, 2, frame)
            rows = (Detection * 100)()
            det.detect(img.shape, img, rows)
            b, cl, s, n = oracle.run(img)
            b64, cl64, s64, n64 = oracle64.run(img)
            got, want = rows_to_tuples(rows), to_detections(b, cl, s, img.shape)
            # random weights give ~100 overlapping low-margin detections: rank swaps between near-equal
            # scores are rounding ties too, so compare as sets of (label, box) with the score tolerance
            same_order = [g[0] == w[0] for g, w in zip(got, want)]
            if all(same_order) and n == n64 and [int(x) for x in cl[:n]] == [int(x) for x in cl64[:n]]:
                flips += compare_rows(got, want, b64, img.shape, conf_tol=1e-3, margin_px=1e-2)
            else:
                gs = sorted((g[0], g[2], g[3], g[4], g[5]) for g in got)
                ws = sorted((w[0], w[2], w[3], w[4], w[5]) for w in want)
                assert sum(a != b for a, b in zip(gs, ws)) <= 4
    assert flips <= 6
