#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_kernel_tests.sh 2>&1 | grep -E "==|passed|failed|error" | head -30
timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu --timeout 300 -p no:cacheprovider 2>&1 | tail -6
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v27.log 2>&1; grep "^{" gpurun_out/bench_v27.log | cut -c1-200; tail -2 gpurun_out/bench_v27.log | cut -c1-300
RF_PDL=0 timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v27_nopdl.log 2>&1; grep "^{" gpurun_out/bench_v27_nopdl.log | cut -c1-200
