#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -p no:cacheprovider -k "gemm or conv3 or ffn or layer" 2>&1 | tail -4
for v in 1 0; do
RF_GEMM_DEEP_RING=$v timeout 600 python bench.py --no-eager-baseline --no-cpu-baseline > gpurun_out/r2j_bench_ring$v.json 2> gpurun_out/r2j_bench_ring$v.err; echo "ring=$v exit $?"
done
python - <<'PY'
import json
for v in (1,0):
    d=json.load(open(f'gpurun_out/r2j_bench_ring{v}.json')); print("deep ring" if v else "3-stage ring", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["avg_launch_us"])
PY
