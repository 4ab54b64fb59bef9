#!/bin/bash
# parity suite + default bench line (+ optional extra bench args in $2..)
mkdir -p gpurun_out
T=${1:-r2q}
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu --maxfail=8 -q -s --timeout 400 --timeout-method=thread 2>&1 | tail -150 > gpurun_out/${T}_pytest.log; grep -E "passed|failed|FAILED" gpurun_out/${T}_pytest.log | tail -8
echo "== bench (default = configs[2])"
timeout 900 python bench.py --steps 200 --warmup 10 2> gpurun_out/${T}_bench.err | tail -1 > gpurun_out/${T}_bench.json; cut -c1-300 gpurun_out/${T}_bench.json; tail -3 gpurun_out/${T}_bench.err
