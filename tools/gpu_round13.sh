#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "attention" --timeout 300 -p no:cacheprovider 2>&1 | tail -15
PYTHONPATH=. timeout 300 python tools/microbench.py 2>&1 | grep -E "^attention" | tee gpurun_out/microbench_attn_v5.log
timeout 600 python -m pytest tests/test_gpu_model.py -q -m gpu --timeout 300 -p no:cacheprovider 2>&1 | tail -8
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v15.log 2>&1; grep "^{" gpurun_out/bench_v15.log | cut -c1-220
