#!/bin/bash
# full ncu capture of the fused depthwise->pointwise kernel (both fused pairs of one step), for profiles/
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_dwpw -s 12 -c 2 \
    -o gpurun_out/prof_fused -f python bench.py --steps 4 --warmup 6 --inflight 1 --no-cpu-baseline --no-roofline > gpurun_out/ncu_fused.log 2>&1
tail -2 gpurun_out/ncu_fused.log | cut -c1-200
ls -la gpurun_out/prof_fused.ncu-rep
