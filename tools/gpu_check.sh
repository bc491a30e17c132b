#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu --timeout 300 -p no:cacheprovider -k "graph or train" 2>&1 | tail -4
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v32.log 2>&1; python - <<'PY'
import json
for l in open('gpurun_out/bench_v32.log'):
    if l.startswith('{'): d=json.loads(l)
print(d['value'], d['ms_per_step'], d['e2e'])
PY
tail -2 gpurun_out/bench_v32.log | cut -c1-200
