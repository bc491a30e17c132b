#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
bash tools/gpu_kernel_tests.sh > /dev/null 2>&1; cat gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu --timeout 600 -p no:cacheprovider 2>&1 | tail -3
timeout 600 python tools/microbench.py 2>&1 | grep -E "fov_crop|direct kernel|attention|layernorm" | tee gpurun_out/microbench_v2.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v2.log 2>&1; tail -1 gpurun_out/bench_v2.log | cut -c1-300
