#!/bin/bash
# N=2: overlapped early all-reduce inside the captured step (parity config and paper dropouts), and the old post-step all-reduce
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline --no-eager-baseline ${@:3} > gpurun_out/$2.json 2> gpurun_out/$2.err; echo "$2 exit $?"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/$2.json').read().strip().splitlines()[-1])
    print('$2', d['value'], d['ms_per_step'], d['e2e']['value'])
except Exception as e:
    print('$2 no json', e); print(open('gpurun_out/$2.err').read()[-1500:])
PY
}
run 29511 r2_n2_overlap
run 29513 r2_n2_overlap_paper_dropout --paper-dropout
