#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_kernel_tests.sh
RF_TMA_TF32_ROUND=1 timeout 300 python tools/gemm_precision.py > gpurun_out/gemm_precision.log 2>&1
RF_TMA_TF32_ROUND=0 timeout 300 python tools/gemm_precision.py >> gpurun_out/gemm_precision.log 2>&1
cat gpurun_out/gemm_precision.log
if ! grep -q "exit 1" gpurun_out/summary.txt; then
  timeout 900 python tools/microbench.py > gpurun_out/microbench.log 2>&1; tail -40 gpurun_out/microbench.log
fi
