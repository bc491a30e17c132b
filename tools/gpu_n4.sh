#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/bench_n4_v29.log 2>&1
echo "exit $?"; grep "^{" gpurun_out/bench_n4_v29.log | cut -c1-250; tail -3 gpurun_out/bench_n4_v29.log | cut -c1-300
