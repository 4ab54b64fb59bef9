"""Rounding-tie classification for the end-to-end parity tests.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The north star's bar is "integer box coordinates and class ids bit-exact, confidences within 1e-3".  The
frozen graph's output is a *ranking* (per-class greedy NMS in score order, then a global sort, then a cut at
100 rows): when two scores differ by less than the fp32 rounding noise of the conv stack, or an IoU lands within
rounding noise of the 0.6 threshold, two correct fp32 evaluations of the same graph (TensorFlow's, the torch-CPU
oracle's, the GPU's) may legitimately order or select rows differently.  This module evaluates the
post-processing in float64 on the float64 heads and reports, with explicit margins,

  * `rows`    every box the float64 NMS keeps whose score is >= (100th score - d_score), in final order
              (so the list extends a little past the top-100 cut), as (label, confidence, x0, y0, x1, y1, fx0..)
  * `groups`  maximal runs of consecutive rows whose neighbouring scores differ by < d_score: the only
              admissible re-orderings are permutations inside a group
  * `poison`  the first rank from which nothing can be asserted row by row, because an NMS decision above it
              is fragile (an IoU within d_iou of the threshold, a suppressor/suppressed pair within d_score of
              each other, a clipped area within 1e-9 of zero, or a score within d_score of the score threshold)

so that a parity test can demand exact rows wherever the float64 evaluation says the answer is well defined
and can name every difference it tolerates.  Follows the same graph nodes as oracle/ssd_graph.py:postprocess
(`Postprocessor/BatchMultiClassNonMaxSuppression/*`).
"""
import numpy as np


def _iou_many(b, i, js):
    bi, bj = b[i], b[js]
    ymin_i, ymax_i = min(bi[0], bi[2]), max(bi[0], bi[2])
    xmin_i, xmax_i = min(bi[1], bi[3]), max(bi[1], bi[3])
    ymin_j, ymax_j = np.minimum(bj[:, 0], bj[:, 2]), np.maximum(bj[:, 0], bj[:, 2])
    xmin_j, xmax_j = np.minimum(bj[:, 1], bj[:, 3]), np.maximum(bj[:, 1], bj[:, 3])
    area_i = (ymax_i - ymin_i) * (xmax_i - xmin_i)
    area_j = (ymax_j - ymin_j) * (xmax_j - xmin_j)
    inter = np.maximum(np.minimum(ymax_i, ymax_j) - np.maximum(ymin_i, ymin_j), 0.0) * \
        np.maximum(np.minimum(xmax_i, xmax_j) - np.maximum(xmin_i, xmin_j), 0.0)
    with np.errstate(divide='ignore', invalid='ignore'):
        iou = inter / (area_i + area_j - inter)
    return np.where((area_i <= 0) | (area_j <= 0), 0.0, iou)


def analyse(oracle64, enc, logits, image_shape, d_score=1e-4, d_iou=1e-3):
    """oracle64: an SsdGraphOracle / SsdModelOracle built with dtype=float64; enc/logits: its raw heads."""
    assert oracle64.dtype == np.float64
    boxes = oracle64.decode(enc)
    sc = oracle64.scores(logits)
    thr_s, thr_iou = float(oracle64.score_thr), float(oracle64.iou_thr)
    _, s_final, _, n_valid = oracle64.postprocess(enc, logits)
    full = n_valid >= oracle64.max_total
    cut = (float(s_final[oracle64.max_total - 1]) - d_score) if full else (thr_s - d_score)
    fragile = []           # scores at which a selection decision is not robust
    kept_all = []          # (score, class, anchor)
    max_out = min(oracle64.max_per_class, boxes.shape[0])
    for c in range(oracle64.num_classes):
        col = sc[:, c]
        cand = np.nonzero(col >= max(cut, thr_s - d_score))[0]
        if cand.size == 0:
            continue
        order = cand[np.lexsort((cand, -col[cand]))]
        kept = []
        for i in order:
            if len(kept) >= max_out:
                break
            s = float(col[i])
            if abs(s - thr_s) < d_score:
                fragile.append(s)
            if s <= thr_s:
                continue
            if kept:
                iou = _iou_many(boxes, i, np.asarray(kept))
                top = float(iou.max())
                if top > thr_iou:
                    sup = kept[int(np.argmax(iou > thr_iou))]
                    # robust only if some kept box suppresses it by a margin and is not a near-tie in score
                    strong = [k for k, v in zip(kept, iou) if v > thr_iou + d_iou and float(col[k]) - s >= d_score]
                    if not strong:
                        fragile.append(float(col[sup]))
                    continue
                if top > thr_iou - d_iou:
                    fragile.append(s)
            kept.append(int(i))
            kept_all.append((s, c, int(i)))
    # global order: score descending, ties -> lower concat index (class-major, then selection rank); the
    # stable sort of the class-ordered list reproduces it
    kept_all.sort(key=lambda t: -t[0])
    rows = []
    max_h, max_w = image_shape[0] - 1, image_shape[1] - 1
    for s, c, i in kept_all:
        b = np.clip(boxes[i], 0.0, 1.0)
        area = (b[2] - b[0]) * (b[3] - b[1])
        if area <= 0:
            if area > -1e-9:
                fragile.append(s)
            continue
        if area < 1e-9:
            fragile.append(s)
        f = (b[1] * max_w, b[0] * max_h, b[3] * max_w, b[2] * max_h)
        rows.append((c + int(oracle64.class_offset), s, int(f[0]), int(f[1]), int(f[2]), int(f[3])) + f)
    groups = []
    a = 0
    for r in range(1, len(rows) + 1):
        if r == len(rows) or rows[r - 1][1] - rows[r][1] >= d_score:
            groups.append((a, r))
            a = r
    worst = max(fragile) if fragile else None
    poison = len(rows)
    if worst is not None:
        poison = sum(1 for r in rows if r[1] > worst + d_score)
    return {'rows': rows, 'groups': groups, 'poison': poison, 'n_valid': n_valid, 'fragile_scores': sorted(fragile, reverse=True),
            'cut': cut, 'd_score': d_score, 'd_iou': d_iou}


def row_matches(got, ref, conf_tol=1e-3, margin_px=2e-3):
    """got: (label, conf, x0, y0, x1, y1); ref: a row of analyse()['rows'].  Integer coordinates must be equal,
    or differ by one where the float64 coordinate is within margin_px of the integer boundary."""
    if got[0] != ref[0] or abs(got[1] - ref[1]) > conf_tol:
        return False
    for k in range(4):
        if got[2 + k] == ref[2 + k]:
            continue
        v = ref[6 + k]
        if abs(got[2 + k] - ref[2 + k]) != 1 or abs(v - round(v)) > margin_px:
            return False
    return True


def compare_with_ties(got_rows, an, max_total=100, conf_tol=1e-3, margin_px=2e-3):
    """Asserts that the first `poison` rows of `got_rows` (label, conf, x0, y0, x1, y1) agree with the float64
    analysis up to permutations inside tie groups.  Returns a dict of counts:
    strict = rows compared one to one, in_group = rows matched inside a multi-row tie group,
    unchecked = rows at or below the poison rank."""
    rows, poison = an['rows'], an['poison']
    n_out = min(max_total, an['n_valid'])
    strict = in_group = 0
    for a, b in an['groups']:
        if a >= n_out or a >= poison:
            break
        if b > poison:
            break
        members = list(range(a, b))
        used = set()
        for r in range(a, min(b, n_out)):
            hit = next((m for m in members if m not in used and row_matches(got_rows[r], rows[m], conf_tol, margin_px)), None)
            assert hit is not None, ('row %d has no partner in its float64 tie group %s' % (r, (a, b)), got_rows[r],
                                     [rows[m][:6] for m in members])
            used.add(hit)
        if b - a == 1:
            strict += 1
        else:
            in_group += min(b, n_out) - a
    checked = strict + in_group
    return {'strict': strict, 'in_group': in_group, 'unchecked': n_out - checked, 'poison': poison, 'n_out': n_out}
