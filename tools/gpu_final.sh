#!/bin/bash
# Round-end evidence in one gpurun call: parity tests, smoke, both bench arms, the ncu launch list and a full
# capture of the GEMM family + the fused kernel.  Logs -> gpurun_out/
bash tools/gpu_check.sh
echo "== bench (default)"
timeout 900 python bench.py 2>/dev/null | tail -1 > gpurun_out/bench_full.json; cut -c1-400 gpurun_out/bench_full.json
echo "== bench --impl reference"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_ref.json; cut -c1-400 gpurun_out/bench_ref.json
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 240 --csv \
    --log-file gpurun_out/launches_final.csv python bench.py --steps 6 --warmup 3 --inflight 1 --no-cpu-baseline --no-roofline > gpurun_out/ncu_bench_final.log 2>&1
tail -1 gpurun_out/ncu_bench_final.log | cut -c1-200
echo "== ncu full (one step of GEMM / fused launches)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_gemm_tc|k_dwpw' -s 150 -c 25 \
    -o gpurun_out/prof_final -f python bench.py --steps 4 --warmup 6 --inflight 1 --no-cpu-baseline --no-roofline > gpurun_out/ncu_full_final.log 2>&1
tail -1 gpurun_out/ncu_full_final.log | cut -c1-200
ls -la gpurun_out/prof_final.ncu-rep gpurun_out/launches_final.csv
