#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2_v33.log 2>&1
echo "exit $?"; grep "^{" gpurun_out/bench_n2_v33.log | cut -c1-250
