#!/bin/bash
# Experimental TMEM-staged A operand (WB_TMEM_A=1): GEMM numerics, then per-layer times with and without it.
mkdir -p gpurun_out
WB_TMEM_A=1 timeout 400 python -m pytest tests/test_gpu_tc.py tests/test_gpu_detect.py -q --timeout=180 -x 2>&1 | tail -6 | tee gpurun_out/ta_pytest.log
for ta in 1; do
  WB_TMEM_A=$ta timeout 300 python bench.py --steps 300 --warmup 20 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/ta_bench.json
  python - <<'PY'
import json
d = json.load(open('gpurun_out/ta_bench.json'))
print('TA value %.0f e2e %.0f' % (d['value'], d['e2e']['value']))
print(' '.join('%.4f' % t for n, t in d['config']['per_layer_ms'] if t > 0.006))
PY
done
