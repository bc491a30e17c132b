#!/bin/bash
# N4 (dataset-side scaling on the device): bit-exactness tests, then the staging micro-benchmark.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "area_resize" --timeout 300 -p no:cacheprovider 2>&1 | tail -n 8
timeout 300 python bench.py --mode stage_micro > gpurun_out/bench_stage_micro.json 2> gpurun_out/bench_stage_micro.err; echo "stage_micro exit $?"; cut -c1-1500 gpurun_out/bench_stage_micro.json; tail -n 3 gpurun_out/bench_stage_micro.err
