"""clock64 timeline of CTA 0 of k_irb_x3 (diagnostic build: make -C watsor_b200/csrc -B EXTRA=-DWB_TRACE)."""
import ctypes
import sys

import numpy as np

sys.path.insert(0, '.')
from tests.workload import v2_coco_model  # noqa: E402
from watsor_b200.engine import Engine  # noqa: E402

SLOTS, ITERS = 12, 64
NAMES = {2: 'exp_issued', 10: 'mid_start', 11: 'mid_done', 4: 'dw_start', 5: 'dw_done', 6: 'prj_issued'}


def main():
    m = v2_coco_model()
    e = Engine(m.to_blob(), max_batch=8, precision=2)
    pre = np.random.default_rng(0).uniform(-1, 1, (8, 300, 300, 3)).astype(np.float32)
    names = {l.name: i for i, l in enumerate(m.layers)}
    import os
    os.environ['WB_IRB'] = '1'
    for name in ('expanded_conv_2/add', 'expanded_conv_4/add'):
        li = names[name]
        for rep in range(3):
            e.backbone(pre, stop_layer=li)
        buf = (ctypes.c_longlong * (SLOTS * ITERS))()
        assert e.lib.wb_trace_read_fused(buf) == 0
        t = np.array(buf, dtype=np.int64).reshape(SLOTS, ITERS)
        t0 = t[9, 0]
        print('== block ending at %s' % name)
        print('   tile:   in_issue %s   in_seen %s   epi_start %s   epi_done %s'
              % (list(t[0, :4] - t0), list(t[1, :4] - t0), list(t[7, :4] - t0), list(t[8, :4] - t0)))
        order = [2, 10, 11, 4, 5, 6]
        print('   %-4s' % 'it' + ''.join('%11s' % NAMES[k] for k in order))
        for it in range(14):
            print('   %-4d' % it + ''.join('%11d' % (t[k, it] - t0) for k in order))


main()
