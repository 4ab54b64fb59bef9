#!/bin/bash
# ncu evidence for profiles/: per-launch durations of one bench command, and a full capture of the top kernel
mkdir -p gpurun_out
PREC=${1:-fp32}
KREGEX=${2:-k_gemm}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 200 --csv \
    --log-file gpurun_out/launches_${PREC}.csv python bench.py --steps 6 --warmup 3 --precision ${PREC} --no-cpu-baseline --no-roofline > gpurun_out/ncu_bench_${PREC}.log 2>&1
tail -2 gpurun_out/ncu_bench_${PREC}.log | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on -k regex:${KREGEX} -s 60 -c 4 \
    -o gpurun_out/prof_${PREC} -f python bench.py --steps 4 --warmup 3 --precision ${PREC} --no-cpu-baseline --no-roofline > gpurun_out/ncu_full_${PREC}.log 2>&1
tail -2 gpurun_out/ncu_full_${PREC}.log | cut -c1-300
ls -la gpurun_out/
