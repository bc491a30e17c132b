#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2_v28.log 2>&1
echo "exit $?"; grep "^{" gpurun_out/bench_n2_v28.log | cut -c1-250; tail -3 gpurun_out/bench_n2_v18.log | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/bench_ref_n2.log 2>&1
echo "ref exit $?"; tail -1 gpurun_out/bench_ref_n2.log | cut -c1-200
