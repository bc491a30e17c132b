#!/bin/bash
mkdir -p gpurun_out
echo "=== TMA store ON"; RF_GEMM_TMA_STORE=1 timeout 300 python tools/gemm_timeline.py 2>&1 | tee gpurun_out/gemm_timeline_tma.log
echo "=== TMA store OFF"; RF_GEMM_TMA_STORE=0 timeout 300 python tools/gemm_timeline.py 2>&1 | tee gpurun_out/gemm_timeline_direct.log
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "gemm or conv3" --timeout 300 -p no:cacheprovider 2>&1 | tail -5
