#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_v18.log 2>&1; grep "^{" gpurun_out/bench_v18.log | cut -c1-200
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_v18.log 2>&1; tail -1 gpurun_out/bench_ref_v18.log | cut -c1-300
timeout 1500 bash tools/gpu_ncu.sh
