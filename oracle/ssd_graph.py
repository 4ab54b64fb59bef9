"""CPU restatement of the frozen TF Object-Detection SSD graph that
watsor/detection/tensorflow_cpu.py:104-121 runs with `sess.run`.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Every arithmetic stage is pinned by OpenCV-dnn as an independent
executor: the conv stack on the reference's own graph (tests/test_oracle_cvdnn.py), the legacy resize on the
reference's own node (tests/test_oracle_cvdnn_resize.py), decode / NMS / top-100 against DetectionOutputLayer
(tests/test_oracle_cvdnn_post.py).  "Parity unpinned" remains for the order of exactly equal scores and the position
of ClipToWindow / pruning in the top-100 assembly: the reference holds no golden vectors at the TF boundary and
TensorFlow cannot run here; this file follows the GraphDef node by node (node names quoted below) and is
sanity-pinned on the reference's behavioural test (test_detect.py:28-77).

The graph (watsor/test/model/cpu.pb) is, per image:

  Cast(u8->f32)                                                   node `Cast`
  ResizeBilinear(legacy: align_corners=0, half_pixel_centers=0)   `Preprocessor/map/while/ResizeImage/resize/ResizeBilinear`
  x * (2/255) - 1                                                 `Preprocessor/mul`, `Preprocessor/sub`
  MobileNet convs: Conv2D / DepthwiseConv2dNative (NHWC, SAME)
      + FusedBatchNormV3(is_training=false) + Relu6               `FeatureExtractor/...`
  heads: 1x1 Conv2D + BiasAdd, reshaped and concatenated          `BoxPredictor_i/...`, `concat`, `concat_1`
  anchors (constant sub-graph)                                    `Concatenate/concat`
  box decode, sigmoid, drop background column                     `Postprocessor/Decode/*`, `convert_scores`, `Slice`
  per-class NonMaxSuppressionV5, concat, sort, clip, prune, pad   `Postprocessor/BatchMultiClassNonMaxSuppression/*`
  classes + 1                                                     `add`

All arithmetic is done in `dtype` (float32 = the graph's own type; float64 is used
by the tests to measure how far a value is from a rounding boundary).
"""
import numpy as np
import torch
import torch.nn.functional as F

from .tf_graph import FrozenGraph

NMS_SCOPE = 'Postprocessor/BatchMultiClassNonMaxSuppression/map/while/MultiClassNonMaxSuppression/'


def same_pad(in_size, k, s):
    """TF `SAME` padding (asymmetric: extra pixel goes after)."""
    out = -(-in_size // s)
    total = max((out - 1) * s + k - in_size, 0)
    return out, total // 2, total - total // 2


class SsdGraphOracle:
    def __init__(self, pb_path, dtype=np.float32, num_threads=None):
        self.g = g = FrozenGraph(pb_path)
        self.dtype = np.dtype(dtype)
        self.tdtype = torch.float32 if self.dtype == np.float32 else torch.float64
        if num_threads:
            torch.set_num_threads(num_threads)

        # ---- preprocess constants
        rb = g.ops('ResizeBilinear')
        assert len(rb) == 1
        self.resize_node = rb[0]
        n = g.node(rb[0])
        assert not n.attr['align_corners'].b and not n.attr['half_pixel_centers'].b
        size = g.const(g.inputs(rb[0])[1][0])
        self.in_h, self.in_w = int(size[0]), int(size[1])
        self.pre_mul = g.const('Preprocessor/mul/x').astype(np.float32)
        self.pre_sub = g.const('Preprocessor/sub/y').astype(np.float32)

        # ---- heads: concat (box encodings), concat_1 (class logits)
        self.box_heads = [i[0] for i in g.inputs('concat')[:-1]]
        self.cls_heads = [i[0] for i in g.inputs('concat_1')[:-1]]
        self.num_classes_p1 = self._reshape_last(self.cls_heads[0])
        self.num_classes = self.num_classes_p1 - 1

        # ---- anchors: constant sub-graph, evaluated with the graph's own ops
        self.anchors = g.eval('Concatenate/concat').astype(np.float32)      # [N,4] corners
        self.num_anchors = self.anchors.shape[0]

        # ---- post-process constants
        d = 'Postprocessor/Decode/'
        self.scale_y = g.const(d + 'truediv/y').astype(np.float32)
        self.scale_x = g.const(d + 'truediv_1/y').astype(np.float32)
        self.scale_h = g.const(d + 'truediv_2/y').astype(np.float32)
        self.scale_w = g.const(d + 'truediv_3/y').astype(np.float32)
        self.logit_scale = g.const('Postprocessor/scale_logits/y').astype(np.float32)
        nms = g.ops('NonMaxSuppressionV5')
        assert len(nms) == self.num_classes, (len(nms), self.num_classes)
        ins = g.inputs(nms[0])
        self.iou_thr = np.float32(g.const(ins[3][0]))
        self.score_thr = np.float32(g.const(ins[4][0]))
        assert float(g.const(ins[5][0])) == 0.0, 'soft-NMS not restated'
        self.max_per_class = int(g.const(NMS_SCOPE + 'Minimum/x'))
        self.max_total = int(g.const(NMS_SCOPE + 'Minimum_3/x'))
        self.class_offset = np.float32(g.const('add/y'))
        self._w = {}

    def _reshape_last(self, reshape_node):
        """Last dim of a BoxPredictor Reshape (its shape is a Pack of consts)."""
        pack = self.g.inputs(reshape_node)[1][0]
        return int(self.g.const(self.g.inputs(pack)[-1][0]))

    # ------------------------------------------------------------------ weights
    def _t(self, name):
        if name not in self._w:
            self._w[name] = torch.from_numpy(
                np.ascontiguousarray(self.g.const(name).astype(self.dtype)))
        return self._w[name]

    # --------------------------------------------------------------- preprocess
    def preprocess(self, image_u8):
        """`Cast` + ResizeBilinear (legacy) + mul + sub.  [H,W,3] u8 -> [h,w,3] f32.

        TF kernel semantics (resize_bilinear_op.cc, LegacyScaler): scale =
        in/(float)out; in = x*scale; lo = floor(in); hi = min(ceil(in), in_size-1);
        lerp = in - floor(in); out = top + (bottom - top)*y_lerp with
        top = tl + (tr - tl)*x_lerp -- no fused multiply-add.
        The resize is always fp32 (the Cast node makes fp32) and so are mul/sub;
        only afterwards is the tensor widened when dtype is float64.
        """
        H, W, C = image_u8.shape
        f32 = np.float32
        x = image_u8.astype(f32)

        def axis(in_size, out_size):
            scale = f32(in_size) / f32(out_size)
            pos = np.arange(out_size, dtype=f32) * scale
            lo_f = np.floor(pos)
            lo = lo_f.astype(np.int64)
            hi = np.minimum(np.ceil(pos).astype(np.int64), in_size - 1)
            return lo, hi, (pos - lo_f).astype(f32)

        ylo, yhi, ylerp = axis(H, self.in_h)
        xlo, xhi, xlerp = axis(W, self.in_w)
        xl = xlerp[None, :, None]
        yl = ylerp[:, None, None]
        tl = x[ylo][:, xlo]
        tr = x[ylo][:, xhi]
        bl = x[yhi][:, xlo]
        br = x[yhi][:, xhi]
        top = tl + (tr - tl) * xl
        bot = bl + (br - bl) * xl
        out = top + (bot - top) * yl
        out = self.pre_mul * out - self.pre_sub
        assert out.dtype == f32
        return out

    # ----------------------------------------------------------------- backbone
    def _eval_tensor(self, name, memo):
        if name in memo:
            return memo[name]
        g = self.g
        n = g.node(name)
        ins = g.inputs(name)
        op = n.op
        if op in ('Conv2D', 'DepthwiseConv2dNative'):
            x = self._eval_tensor(ins[0][0], memo)
            w = self._t(ins[1][0])                      # HWIO  (dw: H,W,C,mult)
            assert n.attr['data_format'].s == b'NHWC' and n.attr['padding'].s == b'SAME'
            assert list(n.attr['dilations'].list.i) in ([], [1, 1, 1, 1])
            s = list(n.attr['strides'].list.i)
            kh, kw = int(w.shape[0]), int(w.shape[1])
            _, pt, pb = same_pad(x.shape[2], kh, s[1])
            _, pl, pr = same_pad(x.shape[3], kw, s[2])
            xp = F.pad(x, (pl, pr, pt, pb))
            if op == 'Conv2D':
                wt = w.permute(3, 2, 0, 1).contiguous()          # O,I,H,W
                v = F.conv2d(xp, wt, stride=(s[1], s[2]))
            else:
                c, m = int(w.shape[2]), int(w.shape[3])
                assert m == 1
                wt = w.permute(2, 3, 0, 1).contiguous()          # C,1,H,W
                v = F.conv2d(xp, wt, stride=(s[1], s[2]), groups=c)
        elif op in ('FusedBatchNormV3', 'FusedBatchNorm'):
            assert not n.attr['is_training'].b
            x = self._eval_tensor(ins[0][0], memo)
            gamma, beta, mean, var = (self._t(i[0]) for i in ins[1:5])
            eps = np.float32(n.attr['epsilon'].f)
            eps_t = torch.tensor(float(eps), dtype=self.tdtype)
            scale = gamma * torch.rsqrt(var + eps_t)
            offset = beta - mean * scale
            v = x * scale.view(1, -1, 1, 1) + offset.view(1, -1, 1, 1)
        elif op == 'Relu6':
            v = torch.clamp(self._eval_tensor(ins[0][0], memo), 0.0, 6.0)
        elif op == 'BiasAdd':
            v = self._eval_tensor(ins[0][0], memo) + self._t(ins[1][0]).view(1, -1, 1, 1)
        elif op in ('AddV2', 'Add'):
            v = self._eval_tensor(ins[0][0], memo) + self._eval_tensor(ins[1][0], memo)
        elif op == 'Identity':
            v = self._eval_tensor(ins[0][0], memo)
        else:
            raise NotImplementedError('oracle backbone: op %s (%s)' % (op, name))
        memo[name] = v
        return v

    def raw_heads(self, pre_hwc, return_memo=False):
        """pre-processed [h,w,3] -> (box_encodings [N,4], class_logits [N,C+1])."""
        x = torch.from_numpy(np.ascontiguousarray(pre_hwc.astype(self.dtype)))
        x = x.permute(2, 0, 1).unsqueeze(0)
        memo = {'Preprocessor/sub': x}
        with torch.no_grad():
            boxes, logits = [], []
            for r in self.box_heads:
                t = self._eval_tensor(self.g.inputs(r)[0][0], memo)      # 1,A*4,H,W
                boxes.append(t.permute(0, 2, 3, 1).reshape(-1, 4))
            for r in self.cls_heads:
                t = self._eval_tensor(self.g.inputs(r)[0][0], memo)
                logits.append(t.permute(0, 2, 3, 1).reshape(-1, self.num_classes_p1))
        enc = torch.cat(boxes, 0).numpy()
        lg = torch.cat(logits, 0).numpy()
        assert enc.shape == (self.num_anchors, 4)
        if return_memo:
            return enc, lg, memo
        return enc, lg

    def feature(self, memo, name):
        """NHWC numpy copy of an intermediate tensor (for layer-wise tests)."""
        return memo[name][0].permute(1, 2, 0).contiguous().numpy()

    # ------------------------------------------------------------ decode+scores
    def decode(self, enc):
        """`Postprocessor/Decode/*` (faster_rcnn_box_coder.decode), op for op."""
        dt = self.dtype.type
        a = self.anchors.astype(self.dtype)
        enc = enc.astype(self.dtype)
        ymin_a, xmin_a, ymax_a, xmax_a = a[:, 0], a[:, 1], a[:, 2], a[:, 3]
        wa = xmax_a - xmin_a                               # get_center.../sub
        ha = ymax_a - ymin_a                               # .../sub_1
        ycenter_a = ymin_a + ha / dt(2)                    # .../add
        xcenter_a = xmin_a + wa / dt(2)                    # .../add_1
        ty = enc[:, 0] / dt(self.scale_y)                  # Decode/truediv
        tx = enc[:, 1] / dt(self.scale_x)                  # truediv_1
        th = enc[:, 2] / dt(self.scale_h)                  # truediv_2
        tw = enc[:, 3] / dt(self.scale_w)                  # truediv_3
        w = np.exp(tw) * wa                                # Exp, mul
        h = np.exp(th) * ha                                # Exp_1, mul_1
        ycenter = ty * ha + ycenter_a                      # mul_2, add
        xcenter = tx * wa + xcenter_a                      # mul_3, add_1
        ymin = ycenter - h / dt(2)
        xmin = xcenter - w / dt(2)
        ymax = ycenter + h / dt(2)
        xmax = xcenter + w / dt(2)
        out = np.stack([ymin, xmin, ymax, xmax], axis=1)
        assert out.dtype == self.dtype
        return out

    def scores(self, logits):
        """`scale_logits` (RealDiv) + `convert_scores` (Sigmoid) + `Slice` [:,1:]."""
        dt = self.dtype.type
        z = logits.astype(self.dtype) / dt(self.logit_scale)
        s = dt(1) / (dt(1) + np.exp(-z))
        return s[:, 1:].astype(self.dtype)

    # ---------------------------------------------------------------------- NMS
    @staticmethod
    def _iou(boxes, i, js):
        """TF non_max_suppression_op.cc IOU(), vectorised over js; same op order."""
        dt = boxes.dtype.type
        bi = boxes[i]
        bj = boxes[js]
        ymin_i, ymax_i = min(bi[0], bi[2]), max(bi[0], bi[2])
        xmin_i, xmax_i = min(bi[1], bi[3]), max(bi[1], bi[3])
        ymin_j, ymax_j = np.minimum(bj[:, 0], bj[:, 2]), np.maximum(bj[:, 0], bj[:, 2])
        xmin_j, xmax_j = np.minimum(bj[:, 1], bj[:, 3]), np.maximum(bj[:, 1], bj[:, 3])
        area_i = (ymax_i - ymin_i) * (xmax_i - xmin_i)
        area_j = (ymax_j - ymin_j) * (xmax_j - xmin_j)
        iy0 = np.maximum(ymin_i, ymin_j)
        ix0 = np.maximum(xmin_i, xmin_j)
        iy1 = np.minimum(ymax_i, ymax_j)
        ix1 = np.minimum(xmax_i, xmax_j)
        inter = np.maximum(iy1 - iy0, dt(0)) * np.maximum(ix1 - ix0, dt(0))
        with np.errstate(divide='ignore', invalid='ignore'):
            iou = inter / (area_i + area_j - inter)
        iou = np.where((area_i <= 0) | (area_j <= 0), dt(0), iou)
        return iou

    def nms_single_class(self, boxes, sc):
        """NonMaxSuppressionV5 with soft_nms_sigma = 0 (hard NMS).

        Candidates = score > score_threshold, visited by (score desc, index asc);
        a candidate is dropped when IoU with any already selected box is
        > iou_threshold (strict); stop at max_per_class selections.
        """
        max_out = min(self.max_per_class, boxes.shape[0])
        cand = np.nonzero(sc > self.dtype.type(self.score_thr))[0]
        order = cand[np.lexsort((cand, -sc[cand].astype(np.float64)))]
        selected = []
        for i in order:
            if len(selected) >= max_out:
                break
            if selected:
                iou = self._iou(boxes, i, np.asarray(selected))
                if np.any(iou > self.dtype.type(self.iou_thr)):
                    continue
            selected.append(int(i))
        return np.asarray(selected, dtype=np.int64)

    def postprocess(self, enc, logits):
        """-> (boxes [100,4], scores [100], classes [100] (already +1), num)."""
        dt = self.dtype.type
        boxes = self.decode(enc)
        sc = self.scores(logits)
        n = boxes.shape[0]
        max_sel = min(self.max_per_class, n)
        all_boxes, all_scores, all_classes = [], [], []
        for c in range(self.num_classes):
            sel = self.nms_single_class(boxes, sc[:, c])
            k = len(sel)
            idx = np.concatenate([sel, np.zeros(max_sel - k, np.int64)])
            s = np.concatenate([sc[sel, c], np.zeros(max_sel - k, self.dtype)])
            s = np.where(np.arange(max_sel) < k, s, dt(-1))
            all_boxes.append(boxes[idx])
            all_scores.append(s)
            all_classes.append(np.full(max_sel, c, self.dtype))
        b = np.concatenate(all_boxes)
        s = np.concatenate(all_scores)
        cl = np.concatenate(all_classes)
        # SortByField: TopKV2 over the whole list = stable sort by score desc
        o = np.argsort(-s.astype(np.float64), kind='stable')
        b, s, cl = b[o], s[o], cl[o]
        # ClipToWindow [0,0,1,1] (window = true_image_shape/300), filter area > 0
        ymin = np.maximum(np.minimum(b[:, 0], dt(1)), dt(0))
        xmin = np.maximum(np.minimum(b[:, 1], dt(1)), dt(0))
        ymax = np.maximum(np.minimum(b[:, 2], dt(1)), dt(0))
        xmax = np.maximum(np.minimum(b[:, 3], dt(1)), dt(0))
        b = np.stack([ymin, xmin, ymax, xmax], 1)
        area = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
        keep = np.nonzero(area > 0)[0]
        b, s, cl = b[keep], s[keep], cl[keep]
        # scores of zero-area boxes -> -1 (no-op after the filter above), count valid
        area = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
        s = np.where(area != 0, s, dt(-1))
        num_valid = int(np.sum(s >= 0))
        o = np.argsort(-s.astype(np.float64), kind='stable')
        b, s, cl = b[o], s[o], cl[o]
        # ChangeCoordinateFrame with window [0,0,1,1]: (b - 0) * (1/(1-0)) == b
        m = min(self.max_total, b.shape[0])
        num_valid = min(num_valid, m)
        b, s, cl = b[:num_valid], s[:num_valid], cl[:num_valid]
        pad = self.max_total - num_valid
        b = np.concatenate([b, np.zeros((pad, 4), self.dtype)])
        s = np.concatenate([s, np.zeros(pad, self.dtype)])
        cl = np.concatenate([cl, np.zeros(pad, self.dtype)]) + dt(self.class_offset)
        return b, s, cl, num_valid

    # ------------------------------------------------------------------ end2end
    def run(self, image_u8):
        """What `sess.run` returns for one frame: boxes[100,4], classes, scores."""
        pre = self.preprocess(image_u8)
        enc, lg = self.raw_heads(pre)
        b, s, cl, n = self.postprocess(enc, lg)
        return b, cl, s, n


def to_detections(boxes, classes, scores, image_shape):
    """watsor/detection/tensorflow_cpu.py:79-90 -- the python write loop.

    `int(boxes[d][0] * max_height)`: under the reference's pinned numpy 1.23 the
    product np.float32 * int is evaluated in float64 (legacy promotion), i.e. it is
    the exact product (24-bit x 11-bit fits a double); int() truncates toward zero.
    Returns rows (label, confidence, x_min, y_min, x_max, y_max).
    """
    max_w = image_shape[1] - 1
    max_h = image_shape[0] - 1
    rows = []
    for d in range(min(len(scores), 100)):
        y0 = int(float(np.float32(boxes[d][0])) * max_h)
        x0 = int(float(np.float32(boxes[d][1])) * max_w)
        y1 = int(float(np.float32(boxes[d][2])) * max_h)
        x1 = int(float(np.float32(boxes[d][3])) * max_w)
        rows.append((int(classes[d]), float(np.float32(scores[d])), x0, y0, x1, y1))
    return rows
