#!/bin/bash
# Round-2 evidence in one gpurun call: default bench line, the CPU reference arm, the Inception line, the ncu launch
# list, a --set full capture of one step's tcgen05 launches, smoke.  Logs -> gpurun_out/ (copied to profiles/ by hand).
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu --maxfail=8 -q -s --timeout 400 --timeout-method=thread 2>&1 | tail -150 > gpurun_out/r02_pytest.log
grep -E "passed|failed|FAILED" gpurun_out/r02_pytest.log | tail -8
if grep -qE "^FAILED|^ERROR|[0-9]+ failed|[0-9]+ error" gpurun_out/r02_pytest.log; then
  # the evidence below must describe a parity-green build: fall back to the direct (unstaged) stem and say so
  export WB_NO_STAGE=1
  echo "TESTS FAILED -> WB_NO_STAGE=1 for the evidence" | tee gpurun_out/r02_fallback.txt
  timeout 900 python -m pytest tests -m gpu --maxfail=8 -q --timeout 400 --timeout-method=thread 2>&1 | tail -30 > gpurun_out/r02_pytest_nostage.log
  grep -E "passed|failed|FAILED" gpurun_out/r02_pytest_nostage.log | tail -8
fi
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee gpurun_out/r02_smoke.log
echo "== bench (default)"
timeout 900 python bench.py 2> gpurun_out/r02_bench.err | tail -1 > gpurun_out/r02_bench_line.json; cut -c1-300 gpurun_out/r02_bench_line.json
echo "== bench --steps 20 --warmup 5 (the driver's length)"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-worker --no-effects --no-real-weights 2>/dev/null | tail -1 > gpurun_out/r02_bench_k20.json; cut -c1-200 gpurun_out/r02_bench_k20.json
echo "== bench --impl reference"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/r02_bench_reference_line.json; cut -c1-300 gpurun_out/r02_bench_reference_line.json
echo "== bench --model inception --cameras 2 (configs[4] per GPU)"
timeout 600 python bench.py --model inception --cameras 2 --steps 200 --warmup 10 --no-worker 2>/dev/null | tail -1 > gpurun_out/r02_bench_inception.json; cut -c1-300 gpurun_out/r02_bench_inception.json
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 300 --csv \
    --log-file gpurun_out/launches_r02.csv python bench.py --steps 6 --warmup 3 --inflight 1 --no-cpu-baseline --no-roofline --no-real-weights --no-worker --no-effects --min-seconds 0 > gpurun_out/r02_ncu_list.log 2>&1
tail -1 gpurun_out/r02_ncu_list.log | cut -c1-120
echo "== ncu --set full (one step of tcgen05 launches; the report stays on the box, its raw page comes back as CSV)"
timeout 900 ncu --set full --clock-control none -k regex:'k_gemm_tc|k_dwpw|k_irb' -s 300 -c 48 \
    -o /tmp/prof_r02 -f python bench.py --steps 6 --warmup 3 --inflight 1 --no-cpu-baseline --no-roofline --no-real-weights --no-worker --no-effects --min-seconds 0 > gpurun_out/r02_ncu_full.log 2>&1
tail -1 gpurun_out/r02_ncu_full.log | cut -c1-120
ncu -i /tmp/prof_r02.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_raw.csv 2>/dev/null
echo "== stem kernel of the final build: time + DRAM bytes (the staging A/B of profiles/r02_stem.md was run earlier)"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 300 ncu --metrics $M --clock-control none -k regex:'k_stem' -s 6 -c 4 --csv --log-file gpurun_out/r02_stem_640_final.csv \
    python bench.py --steps 6 --warmup 3 --inflight 1 --no-cpu-baseline --no-roofline --no-real-weights --no-worker --no-effects --min-seconds 0 > /dev/null 2>&1
grep -h "gpu__time" gpurun_out/r02_stem_640_final.csv | awk -F'","' '{print $NF}' | head -4
echo "== effects kernels: time + DRAM bytes"
timeout 300 ncu --metrics $M --clock-control none -k regex:'k_fx' -c 12 --csv --log-file gpurun_out/r02_fx.csv \
    python -m pytest tests/test_gpu_effects.py -m gpu -q -k "batch_of_cameras" > /dev/null 2>&1
grep -h "gpu__time\|dram__" gpurun_out/r02_fx.csv | awk -F'","' '{print $1, $(NF-2), $NF}' | cut -c1-160 | head -12
ls -la /tmp/prof_r02.ncu-rep gpurun_out/r02_ncu_full_raw.csv gpurun_out/launches_r02.csv
du -sh gpurun_out
