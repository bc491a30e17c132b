#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_raw_errors.txt
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2i_pytest.log 2>&1
echo "pytest exit $?"; tail -8 gpurun_out/r2i_pytest.log | cut -c1-300
RF_ATTN_WIDE=1 timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -p no:cacheprovider -k attention 2>&1 | tail -3
timeout 600 python bench.py --no-eager-baseline --no-cpu-baseline > gpurun_out/r2i_bench_train.json 2> gpurun_out/r2i_bench_train.err; echo "train exit $?"
timeout 600 python bench.py --no-eager-baseline --no-cpu-baseline --paper-dropout --steps 100 > gpurun_out/r2i_bench_train_pd.json 2> gpurun_out/r2i_bench_train_pd.err; echo "train pd exit $?"
python - <<'PY'
import json
for f in ("r2i_bench_train.json","r2i_bench_train_pd.json"):
    d=json.load(open('gpurun_out/'+f)); print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"])
PY
