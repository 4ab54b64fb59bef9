#!/bin/bash
# multi-GPU evidence: N = $1 ranks (torchrun, NCCL over NVLink), optional 2-rank scatter test
N=${1:-2}
T=${2:-r2m}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${T}_topo_${N}.txt 2>&1
if [ "$N" = "2" ]; then
  echo "== pytest scatter (2 ranks, NCCL)"
  timeout 600 python -m pytest tests/test_gpu_scatter.py -m gpu -q -s --timeout 500 --timeout-method=thread 2>&1 | tail -15 > gpurun_out/${T}_pytest_scatter.log; tail -4 gpurun_out/${T}_pytest_scatter.log
fi
echo "== bench N=$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 200 --warmup 10 --no-cpu-baseline --no-real-weights --no-worker --no-roofline 2> gpurun_out/${T}_bench_${N}.err | tail -1 > gpurun_out/${T}_bench_${N}.json
cut -c1-250 gpurun_out/${T}_bench_${N}.json; tail -3 gpurun_out/${T}_bench_${N}.err
if [ "$N" = "2" ]; then
  echo "== bench N=$N, scatter through torch.distributed for comparison"
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
      bench.py --gpus $N --steps 200 --warmup 10 --no-cpu-baseline --no-real-weights --no-worker --no-roofline --scatter-impl torch 2> gpurun_out/${T}_bench_${N}_torch.err | tail -1 > gpurun_out/${T}_bench_${N}_torch.json
  cut -c1-250 gpurun_out/${T}_bench_${N}_torch.json
fi
