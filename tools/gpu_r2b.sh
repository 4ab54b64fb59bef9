#!/bin/bash
# round 2: parity suite, default bench line, ncu launch list, then (diagnostic rebuild on the box) the clock64 trace of
# the GEMM kernel on SSD-MobileNet-v2 shapes.
mkdir -p gpurun_out
T=${1:-r2b}
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu --maxfail=8 -q -s --timeout 400 --timeout-method=thread 2>&1 | tail -150 > gpurun_out/${T}_pytest.log; grep -E "passed|failed|FAILED|configs\[2\] rows|v2 3-class" gpurun_out/${T}_pytest.log | tail -20
echo "== bench (default = configs[2])"
timeout 900 python bench.py --steps 200 --warmup 10 2> gpurun_out/${T}_bench.err | tail -1 > gpurun_out/${T}_bench.json; cut -c1-300 gpurun_out/${T}_bench.json; tail -3 gpurun_out/${T}_bench.err
echo "== bench with the fused inverted-residual-block kernel (WB_IRB=1)"
WB_IRB=1 timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-real-weights --no-worker 2> /dev/null | tail -1 > gpurun_out/${T}_bench_irb.json; cut -c1-200 gpurun_out/${T}_bench_irb.json
echo "== ncu launch list (one batch in flight)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 300 --csv \
    --log-file gpurun_out/${T}_launches.csv python bench.py --steps 6 --warmup 3 --inflight 1 --no-cpu-baseline --no-roofline --no-real-weights --no-worker --min-seconds 0 > gpurun_out/${T}_ncu.log 2>&1
tail -1 gpurun_out/${T}_ncu.log | cut -c1-200
echo "== trace (WB_TRACE rebuild on the box)"
make -C watsor_b200/csrc -B EXTRA=-DWB_TRACE -j16 > /dev/null 2>&1 && timeout 300 python tools/trace_v2.py > gpurun_out/${T}_trace.txt 2>&1; timeout 300 python tools/trace_irb.py > gpurun_out/${T}_trace_irb.txt 2>&1; tail -5 gpurun_out/${T}_trace_irb.txt
