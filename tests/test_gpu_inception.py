"""SSD-Inception-v2 300x300 (BASELINE.json configs[4]: 16 cameras of 1920x1080 on 8 GPUs): architecture descriptor
with seeded synthetic weights (no weights exist offline), checked GPU-vs-oracle layer by layer and end to end on
1920x1080 frames.  Exercises the ops MobileNet does not have: 7x7/s2 stem, max / average pooling with SAME padding,
channel concatenation, 3x3 stride-1 dense convolutions on 38x38 / 19x19 maps."""
import numpy as np
import pytest

from tests.artist import artist_frame
from tests.gpu_util import rows_to_tuples
from watsor_b200.detection.b200 import B200ObjectDetector
from watsor_b200.engine import Engine
from watsor_b200.model import OP_COPY, OP_HEAD, synthetic_ssd_inception_v2
from watsor_b200.stream.share import Detection

pytestmark = pytest.mark.gpu
PRECISIONS = dict(argvalues=[0, 2], ids=['fp32-cuda-cores', 'fp32-3xtf32-tcgen05'])


@pytest.fixture(scope='module')
def inception():
    from oracle.ssd_model import SsdModelOracle
    m = synthetic_ssd_inception_v2(num_classes=90, seed=0, score_thr=1e-8)
    return m, SsdModelOracle(m), SsdModelOracle(m, dtype=np.float64)


@pytest.mark.parametrize('precision', **PRECISIONS)
def test_inception_layer_by_layer(inception, precision):
    m, oracle, _ = inception
    pre = oracle.preprocess(artist_frame(1920, 1080, 1, 0))
    enc, lg, memo = oracle.raw_heads(pre, return_memo=True)
    with Engine(m.to_blob(), device=0, max_batch=2, precision=precision) as e:
        assert np.array_equal(e.preprocess([artist_frame(1920, 1080, 1, 0)])[0], pre)      # 1920x1080 resize, bit-exact
        for li, layer in enumerate(m.layers):
            if layer.op == OP_HEAD:
                continue
            if layer.op == OP_COPY and li + 1 < len(m.layers) and m.layers[li + 1].op == OP_COPY and \
                    m.layers[li + 1].out_off == layer.out_off:
                continue                      # a concat tensor is complete after its last slice
            want = oracle.feature(memo, li)
            _, _, got = e.backbone(pre[None], stop_layer=li, layer_shape=want.shape)
            err, scale = np.abs(got[0] - want).max(), max(1.0, float(np.abs(want).max()))
            assert err <= (1e-4 if precision == 0 else 4e-4) * scale, (li, layer.name, err, scale)
        genc, glg, _ = e.backbone(pre[None])
    assert np.abs(genc[0] - enc).max() <= 1e-3 and np.abs(glg[0] - lg).max() <= 3e-3


@pytest.mark.parametrize('precision', **PRECISIONS)
def test_inception_rows_on_1080p_frames(inception, precision):
    from oracle.ssd_graph import to_detections
    from oracle.ties import analyse, compare_with_ties
    from tests.test_gpu_v2 import _check_frame
    m, o32, o64 = inception
    stats = {'strict_frames': 0, 'tie_frames': 0}
    with B200ObjectDetector(None, device=0, max_batch=2, precision=precision, model_blob=m.to_blob()) as det:
        det.configure_camera(0, 1920, 1080, None)
        det.configure_camera(1, 1920, 1080, None)
        imgs = [artist_frame(1920, 1080, 3, 0), artist_frame(1920, 1080, 4, 1)]
        rows = [(Detection * 100)(), (Detection * 100)()]
        det.detect_batch(imgs, [0, 1], rows, fuse_filters=False)
        for img, r in zip(imgs, rows):
            _check_frame(rows_to_tuples(r), img, o32, o64, stats, to_detections, analyse, compare_with_ties)
    print('inception rows:', stats)
    assert stats['strict_frames'] + stats['tie_frames'] == 2
