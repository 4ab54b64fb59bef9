#!/bin/bash
# round 2, first call: the parity suite with the configs[2] tests, then the new default bench line (v2 / 90 classes /
# 8 masks) with its per-layer table, and an ncu launch list of one step.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/smi.txt 2>&1
echo "== pytest -m gpu"
timeout 1800 python -m pytest tests -m gpu --maxfail=8 -q -s 2>&1 | tail -150 > gpurun_out/r2a_pytest.log; tail -25 gpurun_out/r2a_pytest.log
echo "== bench (default = configs[2])"
timeout 900 python bench.py --steps 200 --warmup 10 2> gpurun_out/r2a_bench.err | tail -1 > gpurun_out/r2a_bench.json; cut -c1-600 gpurun_out/r2a_bench.json; tail -3 gpurun_out/r2a_bench.err
echo "== ncu launch list (one batch in flight)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 400 --csv \
    --log-file gpurun_out/r2a_launches.csv python bench.py --steps 6 --warmup 3 --inflight 1 --no-cpu-baseline --no-roofline --no-real-weights --no-worker --min-seconds 0 > gpurun_out/r2a_ncu.log 2>&1
tail -2 gpurun_out/r2a_ncu.log | cut -c1-300
wc -l gpurun_out/r2a_launches.csv
