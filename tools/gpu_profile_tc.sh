#!/bin/bash
# full ncu capture of one step's tensor-core GEMM launches (all layers), for profiles/
mkdir -p gpurun_out
PREC=${1:-tf32x3}
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tc -s 162 -c 27 \
    -o gpurun_out/prof_layers_${PREC} -f python bench.py --steps 4 --warmup 6 --inflight 1 --precision ${PREC} --no-cpu-baseline --no-roofline > gpurun_out/ncu_layers_${PREC}.log 2>&1
tail -2 gpurun_out/ncu_layers_${PREC}.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
