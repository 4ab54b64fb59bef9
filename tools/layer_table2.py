#!/usr/bin/env python
"""Per-layer table of a bench line for profiles/ (round 2: any model of bench.py --model).

    python tools/layer_table2.py profiles/r02_bench_line.json v2 > profiles/r02_layer_table.md

Algorithmic bytes = input + output activations of the batch + the layer's weights, fp32 (the 640x480 u8 frames for the
stem); FLOPs = 2 x MACs.  Times are the bench line's `per_layer_ms` (CUDA events, every kernel launched
10x back to back, warm L2).  A row with time 0 ran inside the kernel of a neighbouring row (fused)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from watsor_b200.model import OP_ADD, OP_COPY, OP_DW, OP_HEAD, OP_NAMES, OP_STEM  # noqa: E402

line = json.load(open(sys.argv[1]))
kind = sys.argv[2] if len(sys.argv) > 2 else 'v2'
model, desc = bench.load_model(kind)
n = line['config']['cameras_per_gpu']
times = line['config']['per_layer_ms']
peaks = json.load(open('MEASURED_PEAKS.json')) if os.path.isfile('MEASURED_PEAKS.json') else {}
hbm, tf = peaks.get('hbm_gbs', 6650.0), peaks.get('bf16_tflops_sustained', 1400.0)
W, H = (1920, 1080) if kind == 'inception' else (640, 480)
print('# per-layer table: %s, batch of %d frames, %s\n' % (desc, n, line['config']['precision']))
print('Source: `%s` (value %.0f frames/s, %.3f ms per step with %d batches in flight).  Peaks: HBM %.0f GB/s, dense bf16 '
      '%.0f TFLOP/s (MEASURED_PEAKS.json); 3xTF32 ceiling = bf16 / 6 = %.0f TFLOP/s.  Serial sum of the rows: %.0f us.\n'
      % (os.path.basename(sys.argv[1]), line['value'], line['ms_per_step'], line['config']['batches_in_flight'], hbm, tf,
         tf / 6, 1e3 * sum(t for _, t in times)))
print('| # | layer | op | in -> out (HxWxC) | k/s | MMAC/frame | alg. MB/batch | us | GB/s | % HBM | TFLOP/s | % 3xTF32 ceiling |')
print('|---:|---|---|---|---|---:|---:|---:|---:|---:|---:|---:|')
for i, l in enumerate(model.layers):
    t = times[i][1] * 1e3
    w_b = 4.0 * l.kh * l.kw * (l.in_c if l.op != OP_DW else 1) * l.out_c if l.w_tensor >= 0 else 0.0
    in_b = W * H * 3 if l.op == OP_STEM else l.in_h * l.in_w * l.in_c * 4 * (2 if l.op == OP_ADD else 1)
    out_c = l.in_c if l.op == OP_COPY else l.out_c
    byts = n * (in_b + l.out_h * l.out_w * out_c * 4) + w_b
    flops = 2.0 * l.macs * n
    if t <= 0.5:
        rate = '| (fused) | | | | |'
    else:
        gbs, tfl = byts / t / 1e3, flops / t / 1e6
        rate = '| %.1f | %.0f | %.1f | %.1f | %.1f |' % (t, gbs, 100 * gbs / hbm, tfl, 100 * tfl / (tf / 6))
    print('| %d | `%s` | %s | %dx%dx%d -> %dx%dx%d | %dx%d/%d | %.2f | %.2f %s'
          % (i, l.name[-36:], OP_NAMES[l.op], l.in_h, l.in_w, l.in_c, l.out_h, l.out_w, l.out_c, l.kh, l.kw, l.stride,
             l.macs / 1e6, byts / 1e6, rate))
post = times[len(model.layers)][1] * 1e3
pb = n * (model.num_anchors * (5 + model.num_classes) * 4 + 7200.0)
print('| | post (decode + NMS + merge/filter) | post | %d anchors x %d classes | | | %.2f | %.1f | %.0f | %.2f | | |'
      % (model.num_anchors, model.num_classes, pb / 1e6, post, pb / post / 1e3, 100 * pb / post / 1e3 / hbm))
