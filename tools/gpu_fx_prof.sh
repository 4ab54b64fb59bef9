#!/bin/bash
# source-level profile of the effects kernels on the bench's effects workload (8 masked cameras, 8 labels per frame)
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'k_fx' -s 40 -c 2 -o gpurun_out/src_fx -f \
    python bench.py --steps 5 --warmup 3 --inflight 1 --no-cpu-baseline --no-worker --no-real-weights --no-roofline --min-seconds 0 > gpurun_out/src_fx.log 2>&1
ls -la gpurun_out/src_fx.ncu-rep; tail -2 gpurun_out/src_fx.log | cut -c1-200
