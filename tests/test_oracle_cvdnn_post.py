"""A second, independent executor for the part of the graph that oracle/ssd_graph.py restates after the heads: box
decode, per-class greedy NMS and the global top-100.  TensorFlow is not installable here, but OpenCV 4.13's dnn module
ships `DetectionOutputLayer` -- the SSD post-processing of Caffe / OpenCV model zoos, written by other people against
the same definitions: CENTER_SIZE box coding with variances (0.1, 0.1, 0.2, 0.2) = the graph's scale factors
(10, 10, 5, 5); confidence threshold (strict >); per class: candidates by descending score, a candidate is dropped
when its Jaccard overlap with a kept box exceeds the threshold; `keep_top_k` highest scores over all classes.

What this pins (and tests/test_oracle_cvdnn.py does not): the decode formulas and their operand order, the strictness
of both thresholds, the greedy visiting order, "suppress against kept boxes only", the cross-class top-100.
What it cannot pin: TF's tie order for equal scores (lower anchor index first) -- continuous random scores have no
ties -- and float rounding at the thresholds: both executors work in float32 with different operation orders, so
every case first checks (in float64) that no decision sits within 1e-5 of a threshold.  ClipToWindow happens after NMS
in the graph (the layer's own `clip` would clip before): it is applied to OpenCV's output here.  CPU only."""
import cv2
import numpy as np
import pytest

from oracle.ssd_model import SsdModelOracle

PROTO = '''
name: "ssd_post"
input: "loc"
input_shape {{ dim: 1 dim: {n4} }}
input: "conf"
input_shape {{ dim: 1 dim: {nc} }}
input: "prior"
input_shape {{ dim: 1 dim: 2 dim: {n4} }}
layer {{
  name: "detection_out"
  type: "DetectionOutput"
  bottom: "loc"
  bottom: "conf"
  bottom: "prior"
  top: "detection_out"
  detection_output_param {{
    num_classes: {c1}
    share_location: true
    background_label_id: 0
    nms_param {{ nms_threshold: {iou} top_k: -1 }}
    code_type: CENTER_SIZE
    keep_top_k: {keep}
    confidence_threshold: {thr}
    clip: false
  }}
}}
'''


def opencv_detection_output(oracle, enc, scores_with_background):
    n, c1 = scores_with_background.shape
    proto = PROTO.format(n4=n * 4, nc=n * c1, c1=c1, iou=repr(float(oracle.iou_thr)), keep=oracle.max_total,
                         thr=repr(float(oracle.score_thr)))
    net = cv2.dnn.readNetFromCaffe(np.frombuffer(proto.encode(), np.uint8))
    a = oracle.anchors                                   # [ymin, xmin, ymax, xmax]
    prior = np.zeros((1, 2, n * 4), np.float32)
    prior[0, 0] = a[:, [1, 0, 3, 2]].reshape(-1)         # Caffe: xmin, ymin, xmax, ymax
    prior[0, 1] = np.tile(np.float32([1 / oracle.scale_x, 1 / oracle.scale_y, 1 / oracle.scale_w, 1 / oracle.scale_h]), n)
    loc = enc[:, [1, 0, 3, 2]].reshape(1, -1).astype(np.float32)       # (ty, tx, th, tw) -> (tx, ty, tw, th)
    net.setInput(loc, 'loc')
    net.setInput(scores_with_background.reshape(1, -1).astype(np.float32), 'conf')
    net.setInput(prior, 'prior')
    out = net.forward().reshape(-1, 7)                   # image, label, score, xmin, ymin, xmax, ymax
    out = out[out[:, 1] >= 1]
    boxes = np.clip(out[:, [4, 3, 6, 5]], 0.0, 1.0)      # ClipToWindow after NMS, as in the graph
    area = (boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1])
    keep = area > 0
    return boxes[keep], out[keep, 2], out[keep, 1]


def fragile(oracle64, enc, logits, margin=1e-5):
    """True when some decision of the float64 evaluation sits within `margin` of a threshold."""
    boxes = oracle64.decode(enc)
    sc = oracle64.scores(logits)
    if np.any(np.abs(sc - oracle64.score_thr) < min(margin, 0.5 * float(oracle64.score_thr))):
        return True
    for c in range(sc.shape[1]):
        cand = np.nonzero(sc[:, c] > oracle64.score_thr)[0]
        order = cand[np.argsort(-sc[cand, c], kind='stable')]
        if len(order) > 1 and np.min(np.abs(np.diff(sc[order, c]))) == 0:
            return True                                   # an exact tie: visiting order is implementation defined
        kept = []
        for i in order:
            if len(kept) >= oracle64.max_per_class:
                break
            if kept:
                iou = oracle64._iou(boxes, i, np.asarray(kept))
                if np.any(np.abs(iou - oracle64.iou_thr) < margin):
                    return True
                if np.any(iou > oracle64.iou_thr):
                    continue
            kept.append(int(i))
    return False


def heads(rng, oracle, busy):
    n, c1 = oracle.num_anchors, oracle.num_classes_p1
    enc = rng.normal(0.0, 1.2, (n, 4)).astype(np.float32)
    # everything that is not "hot" stays below the score threshold (sigmoid(-30) = 9e-14 < 1e-8): with 172 k live
    # candidates some pair always sits within rounding distance of a threshold and nothing could be asserted
    logits = rng.normal(-30.0, 1.0, (n, c1)).astype(np.float32)
    hot = rng.choice(n, size=busy, replace=False)
    logits[hot, rng.integers(1, c1, size=busy)] = rng.normal(1.0, 1.5, busy).astype(np.float32)
    # near-duplicates so that NMS has something to suppress: the neighbouring anchor is made to decode to (almost) the
    # same box as the hot one -- encode() is the inverse of the graph's decode -- with a lower score in the same class
    a = oracle.anchors.astype(np.float64)
    ha, wa = a[:, 2] - a[:, 0], a[:, 3] - a[:, 1]
    yca, xca = a[:, 0] + ha / 2, a[:, 1] + wa / 2
    sy, sx, sh, sw = (float(v) for v in (oracle.scale_y, oracle.scale_x, oracle.scale_h, oracle.scale_w))
    for k in range(0, busy, 2):
        j = int(hot[k])
        i = j + 1 if j + 1 < n else j - 1
        if i in hot:
            continue
        yc = enc[j, 0] / sy * ha[j] + yca[j]
        xc = enc[j, 1] / sx * wa[j] + xca[j]
        h, w = np.exp(enc[j, 2] / sh) * ha[j], np.exp(enc[j, 3] / sw) * wa[j]
        yc, xc = yc + rng.normal(0, 0.02) * h, xc + rng.normal(0, 0.02) * w
        h, w = h * np.exp(rng.normal(0, 0.03)), w * np.exp(rng.normal(0, 0.03))
        enc[i] = np.float32([(yc - yca[i]) / ha[i] * sy, (xc - xca[i]) / wa[i] * sx, np.log(h / ha[i]) * sh,
                             np.log(w / wa[i]) * sw])
        logits[i] = logits[j]
        live = logits[i] > -15
        logits[i, live] -= np.float32(abs(rng.normal(0.4, 0.2)) + 0.05)
    return enc, logits


@pytest.mark.parametrize('which', ['shapes', 'coco90'])
def test_decode_nms_top100_equal_opencv_detection_output(which, shapes_model, coco_model):
    model = shapes_model if which == 'shapes' else coco_model
    oracle, oracle64 = SsdModelOracle(model), SsdModelOracle(model, dtype=np.float64)
    rng = np.random.default_rng(11)
    checked = suppressed = 0
    for trial in range(12):
        enc, logits = heads(rng, oracle, busy=[40, 150, 400][trial % 3])
        if fragile(oracle64, enc, logits):
            continue
        checked += 1
        b, s, cl, num = oracle.postprocess(enc, logits)
        sig = np.zeros((oracle.num_anchors, oracle.num_classes_p1), np.float32)
        sig[:, 1:] = oracle.scores(logits)                 # the layer takes probabilities; column 0 = background
        cb, cs, cc = opencv_detection_output(oracle, enc, sig)
        assert num == len(cs), (which, trial, num, len(cs))
        if num < oracle.max_total:                         # below the cap every surviving candidate is in the output
            suppressed += int((oracle.scores(logits) > oracle.score_thr).sum()) - num
        o1 = np.lexsort((cl[:num], -s[:num].astype(np.float64)))
        o2 = np.lexsort((cc, -cs.astype(np.float64)))
        assert np.array_equal(cl[:num][o1], cc[o2]), (which, trial)
        assert np.max(np.abs(s[:num][o1] - cs[o2])) <= 1e-6, (which, trial)
        assert np.max(np.abs(b[:num][o1] - cb[o2])) <= 2e-6, (which, trial, float(np.max(np.abs(b[:num][o1] - cb[o2]))))
    assert checked >= 8 and suppressed >= 30, (checked, suppressed)        # NMS did drop boxes in these cases
