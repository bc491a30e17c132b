#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "attention" --timeout 300 -p no:cacheprovider 2>&1 | tail -8
timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu --timeout 300 -p no:cacheprovider 2>&1 | tail -6
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v31.log 2>&1; grep "^{" gpurun_out/bench_v31.log | cut -c1-200; tail -2 gpurun_out/bench_v31.log | cut -c1-200
