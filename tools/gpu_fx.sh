#!/bin/bash
# effects parity suite + the bench line's effects record (short bench: no CPU baseline, no worker, no second model)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_effects.py tests/test_gpu_stages.py -m gpu -q --timeout 300 --timeout-method=thread 2>&1 | tail -5 | tee gpurun_out/r02_fx_pytest.log
timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-worker --no-real-weights --no-roofline 2>/dev/null | tail -1 > gpurun_out/r02_bench_effects.json
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_effects.json')); print(d['value'], d['effects'])"
