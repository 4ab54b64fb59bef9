#!/bin/bash
# Round-2 starting point: pipeline timeline of the 19x19x512 GEMM with and without the TMEM-staged A operand.
#   make -C watsor_b200/csrc -B EXTRA=-DWB_TRACE && gpurun -- bash tools/gpu_trace_ta.sh; make -C watsor_b200/csrc -B
mkdir -p gpurun_out
timeout 200 python tools/trace_pipeline.py > gpurun_out/trace_smem_a.txt 2>&1
WB_TMEM_A=1 timeout 200 python tools/trace_pipeline.py > gpurun_out/trace_tmem_a.txt 2>&1
tail -20 gpurun_out/trace_smem_a.txt
tail -20 gpurun_out/trace_tmem_a.txt
