#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "gemm or conv3" --timeout 300 -p no:cacheprovider 2>&1 | tail -5
timeout 600 python -m pytest tests/test_gpu_model.py -q -m gpu --timeout 300 -p no:cacheprovider 2>&1 | tail -5
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v16.log 2>&1; grep "^{" gpurun_out/bench_v16.log | cut -c1-220
RF_GEMM_DEEP=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v16_nodeep.log 2>&1; grep "^{" gpurun_out/bench_v16_nodeep.log | cut -c1-220
