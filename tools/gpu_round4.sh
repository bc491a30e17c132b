#!/bin/bash
mkdir -p gpurun_out
RF_GEMM_TMA_STORE=1 timeout 300 python tools/gemm_timeline.py 2>&1 | tee gpurun_out/gemm_timeline_tma.log
echo "=== TMA store OFF"; RF_GEMM_TMA_STORE=0 timeout 300 python tools/gemm_timeline.py 2>&1 | head -4
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "gemm or conv3" --timeout 300 -p no:cacheprovider 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu --timeout 600 -p no:cacheprovider 2>&1 | tail -3
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v1.log 2>&1; tail -2 gpurun_out/bench_v1.log
