#!/bin/bash
# A/B of opt-in switches: short bench runs (device value + e2e), plus the tensor-core tests under the switch
mkdir -p gpurun_out
T=${1:-r2ab}
B="python bench.py --steps 400 --warmup 10 --no-cpu-baseline --no-real-weights --no-worker --no-effects --no-roofline"
for cfg in "default" "WB_GEMM_2CTA=1" "WB_TMEM_A=1" "WB_GEMM_2CTA=1 WB_WIDE_N=1"; do
  name=$(echo "$cfg" | tr ' =' '__')
  if [ "$cfg" = "default" ]; then env timeout 300 $B 2>/dev/null | tail -1 > gpurun_out/${T}_${name}.json
  else env $cfg timeout 300 $B 2>/dev/null | tail -1 > gpurun_out/${T}_${name}.json; fi
  python - <<PY
import json
d=json.load(open('gpurun_out/${T}_${name}.json'))
print('%-32s value %.0f  long %.0f  e2e %.0f' % ('$cfg', d['value'], d['value_long']['value'], d['e2e']['value']))
PY
done
echo "== tests under WB_GEMM_2CTA=1"
WB_GEMM_2CTA=1 timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_v2.py tests/test_gpu_stages.py -m gpu -q --timeout 400 --timeout-method=thread 2>&1 | tail -5
