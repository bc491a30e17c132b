#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "gemm or conv3" --timeout 120 -p no:cacheprovider 2>&1 | tail -4
timeout 600 python -m pytest tests/test_gpu_model.py -q -m gpu --timeout 300 -p no:cacheprovider 2>&1 | tail -4
timeout 300 python tools/microbench.py 2>&1 | grep -E "^gemm" | tee gpurun_out/microbench_gemm_persistent.log
RF_GEMM_PERSISTENT=0 timeout 300 python tools/microbench.py 2>&1 | grep -E "^gemm" | tee gpurun_out/microbench_gemm_tilewise.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v7.log 2>&1; grep "^{" gpurun_out/bench_v7.log | cut -c1-220; tail -2 gpurun_out/bench_v7.log | cut -c1-300
