#!/bin/bash
# source-level ncu captures (--set full --import-source on) of single launches: the stem (direct and staged tile build)
# and k_nms.  Reports are small (one launch each) and come back in gpurun_out/.
mkdir -p gpurun_out
B="python bench.py --steps 6 --warmup 3 --inflight 1 --no-cpu-baseline --no-roofline --no-real-weights --no-worker --min-seconds 0"
WB_STAGE=0 WB_NO_STAGE=1 timeout 300 ncu --set full --import-source on --clock-control none -k regex:'k_stem' -s 6 -c 1 -o gpurun_out/src_stem_direct -f $B > gpurun_out/src_stem_direct.log 2>&1
WB_STAGE=1 timeout 300 ncu --set full --import-source on --clock-control none -k regex:'k_stem' -s 6 -c 1 -o gpurun_out/src_stem_staged -f $B > gpurun_out/src_stem_staged.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:'k_nms' -s 6 -c 1 -o gpurun_out/src_nms -f $B > gpurun_out/src_nms.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:'k_dw_strip' -s 12 -c 1 -o gpurun_out/src_dw -f $B > gpurun_out/src_dw.log 2>&1
ls -la gpurun_out/src_*.ncu-rep
