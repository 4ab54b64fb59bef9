"""clock64 timeline of CTA 0 of k_gemm_tc on SSD-MobileNet-v2 shapes (diagnostic build only):

    make -C watsor_b200/csrc -B EXTRA=-DWB_TRACE && gpurun -- python tools/trace_v2.py

Stamps: entry (kernel entry), setup (after the prologue barrier: mbarrier init + TMEM alloc), per k-block
tma_issue / conv_full / conv_done / mma_ready / mma_issued, epi_start (accumulator complete), epi_done, exit.
"""
import ctypes
import sys

import numpy as np

sys.path.insert(0, '.')
from tests.workload import v2_coco_model  # noqa: E402
from watsor_b200.engine import Engine  # noqa: E402

SLOTS, ITERS = 12, 64
GEMM = {0: 'tma_issue', 1: 'conv_full', 2: 'conv_done', 3: 'mma_ready', 4: 'mma_issued'}


def dump(e, n_it):
    buf = (ctypes.c_longlong * (SLOTS * ITERS))()
    assert e.lib.wb_trace_read_gemm(buf) == 0
    t = np.array(buf, dtype=np.int64).reshape(SLOTS, ITERS)
    t0 = t[8, 0]
    print('   entry 0   setup %d   epi_start %d   staged %d   epi_done %d   cluster_synced %d   exit %d   (cycles @1.965 GHz; %.2f us total)'
          % (t[7, 0] - t0, t[5, 0] - t0, t[10, 0] - t0, t[6, 0] - t0, t[11, 0] - t0, t[9, 0] - t0, (t[9, 0] - t0) / 1965.0))
    print('   %-4s' % 'it' + ''.join('%11s' % GEMM[k] for k in sorted(GEMM)))
    for it in range(n_it):
        print('   %-4d' % it + ''.join('%11d' % (t[k, it] - t0) for k in sorted(GEMM)))


def main():
    m = v2_coco_model()
    e = Engine(m.to_blob(), max_batch=8, precision=2)
    pre = np.random.default_rng(0).uniform(-1, 1, (8, 300, 300, 3)).astype(np.float32)
    names = {l.name: i for i, l in enumerate(m.layers)}
    for name in ('expanded_conv_7/expand', 'expanded_conv_7/project', 'expanded_conv_2/expand', 'expanded_conv_13/expand',
                 'BoxPredictor_0', 'Conv_1', 'layer_19_2_Conv2d_2_3x3_s2_512'):
        li = names[name]
        l = m.layers[li]
        if l.op == 6:
            for rep in range(2):
                e.backbone(pre, stop_layer=li - 1)
            # heads have no activation output: run the whole net, the last k_gemm_tc launch is BoxPredictor_5
            continue
        for rep in range(3):
            e.backbone(pre, stop_layer=li)
        kb = -(-(l.kh * l.kw * l.in_c) // 32)
        print('== %s  (M=%d, K=%d, N=%d, k-blocks %d)' % (name, 8 * l.out_h * l.out_w, l.kh * l.kw * l.in_c, l.out_c, kb))
        dump(e, min(kb, 12))


main()
