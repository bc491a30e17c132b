#!/bin/bash
# usage: gpurun --gpus N -- 'bash tools/gpu_scale.sh N [extra bench flags]'  -> gpurun_out/bench_nN.json (launched as the driver does)
N=${1:-2}; shift
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 3 \
  --no-cpu-baseline --no-eager-baseline "$@" > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "exit $?"; cut -c1-400 gpurun_out/bench_n$N.json; tail -n 3 gpurun_out/bench_n$N.err | cut -c1-300
