#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "gemm or conv3" --timeout 300 -p no:cacheprovider 2>&1 | tail -4
timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu --timeout 300 -p no:cacheprovider 2>&1 | tail -4
PYTHONPATH=. timeout 200 python tools/gemm_epi_bench.py 2>&1 | grep -E "colsum|dx dgrad|saved" | tee gpurun_out/gemm_epi_v5.log
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v30.log 2>&1; grep "^{" gpurun_out/bench_v30.log | cut -c1-200; tail -2 gpurun_out/bench_v30.log | cut -c1-200
