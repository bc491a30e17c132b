#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "gemm or conv3" --timeout 300 -p no:cacheprovider 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu --timeout 600 -p no:cacheprovider 2>&1 | tail -8
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v5.log 2>&1; grep "^{" gpurun_out/bench_v5.log | cut -c1-220; tail -3 gpurun_out/bench_v5.log | cut -c1-300
head -30 gpurun_out/gemm_shapes.txt
