#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -p no:cacheprovider -k "fov" > gpurun_out/r2h_pytest_crop.log 2>&1
echo "crop pytest exit $?"; tail -12 gpurun_out/r2h_pytest_crop.log | cut -c1-300
RF_CROP_TILED=1 timeout 300 python bench.py --mode crop_micro > gpurun_out/r2h_bench_crop_strip.json 2> gpurun_out/r2h_bench_crop.err; echo "crop exit $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2h_bench_crop_strip.json'))
for r in d["crop"]: print(r["case"], r["layout"], r["ms"], r["gbs"], r["frac_of_hbm_peak"])
PY
