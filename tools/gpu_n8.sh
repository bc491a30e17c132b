#!/bin/bash
mkdir -p gpurun_out
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_n8_v29.log 2>&1
echo "exit $?"; grep "^{" gpurun_out/bench_n8_v29.log | cut -c1-250; tail -3 gpurun_out/bench_n8_v29.log | cut -c1-300
