#!/bin/bash
# Quick iteration run: tensor-core tests, a short bench, the per-layer table of its JSON line.  Logs -> gpurun_out/
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -q --timeout=300 2>&1 | tail -5 | tee gpurun_out/quick_pytest.log
timeout 600 python bench.py --steps 1000 --warmup 20 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/quick_bench.json
python - <<'PY'
import json
d = json.load(open('gpurun_out/quick_bench.json'))
print('value %.0f e2e %.0f ms/step %.4f launches/step %d' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['gpu_launches'] // d['steps']))
print(' '.join('%s:%.4f' % (n.split('/')[-1][-14:], t) for n, t in d['config']['per_layer_ms'] if t > 0.006))
print('sum %.4f' % sum(t for _, t in d['config']['per_layer_ms']))
PY
