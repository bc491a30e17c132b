#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "fov_crop" --timeout 300 -p no:cacheprovider 2>&1 | tail -5
PYTHONPATH=. timeout 300 python tools/microbench.py 2>&1 | grep -E "^fov" | tee gpurun_out/microbench_crop_v3.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v20.log 2>&1; grep "^{" gpurun_out/bench_v20.log | cut -c1-200
python - <<'PY'
import json
for l in open('gpurun_out/bench_v20.log'):
    if l.startswith('{'): d=json.loads(l)
print(d['kernels']['fov_crop'])
PY
