#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
bash tools/gpu_kernel_tests.sh > /dev/null 2>&1; cat gpurun_out/summary.txt | grep -E "exit|passed|failed"
timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu --timeout 600 -p no:cacheprovider 2>&1 | tail -4
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v6.log 2>&1; grep "^{" gpurun_out/bench_v6.log | cut -c1-220; tail -3 gpurun_out/bench_v6.log | cut -c1-300
timeout 300 python tools/microbench.py 2>&1 | grep -E "attention|frame_ffn" 
