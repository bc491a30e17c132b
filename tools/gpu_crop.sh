#!/bin/bash
# FoV crop: parity tests of the three kernels, then the crop micro-benchmark with each of them.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k fov_crop --timeout 300 -p no:cacheprovider > gpurun_out/kernels_fov_crop.log 2>&1
echo "== fov_crop tests: exit $?"; tail -n 15 gpurun_out/kernels_fov_crop.log
timeout 300 python bench.py --mode crop_micro > gpurun_out/bench_crop_micro_default.json 2> gpurun_out/bench_crop_micro_default.err; echo "default exit $?"
RF_CROP_PAIRS=1 timeout 300 python bench.py --mode crop_micro > gpurun_out/bench_crop_micro_staged.json 2> /dev/null; echo "pairs exit $?"
python - <<'PY'
import json
for k in ("default", "staged"):
    try:
        d = json.loads(open(f"gpurun_out/bench_crop_micro_{k}.json").read().strip().splitlines()[-1])
        for r in d["crop"]:
            print(k, r["case"], r["layout"], r["out_dtype"], r["ms"], r["gbs"], r["frac_of_hbm_peak"])
    except Exception as e:
        print(k, "failed", e)
PY
