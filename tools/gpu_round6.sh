#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "layernorm or reductions or conv3 or attention" --timeout 300 -p no:cacheprovider 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu --timeout 600 -p no:cacheprovider 2>&1 | tail -15
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/bench_v4_eager.log 2>&1; grep "^{" gpurun_out/bench_v4_eager.log | cut -c1-220
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v4_graph.log 2>&1; grep "^{" gpurun_out/bench_v4_graph.log | cut -c1-220; tail -5 gpurun_out/bench_v4_graph.log | cut -c1-300
