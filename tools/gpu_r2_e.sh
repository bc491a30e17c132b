#!/bin/bash
# round 2, call E: full GPU suite after the test / kernel changes; u8-frame bench; TC attention microbench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_raw_errors.txt
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2e_pytest.log 2>&1
echo "pytest exit $?"; tail -25 gpurun_out/r2e_pytest.log | cut -c1-400
python tools/attn_bench.py > gpurun_out/r2e_attn_bench.txt 2>&1; cat gpurun_out/r2e_attn_bench.txt
timeout 600 python bench.py --no-eager-baseline --no-cpu-baseline --u8-frames > gpurun_out/r2e_bench_train_u8.json 2> gpurun_out/r2e_bench_train_u8.err; echo "train u8 exit $?"; tail -c 700 gpurun_out/r2e_bench_train_u8.json; tail -3 gpurun_out/r2e_bench_train_u8.err
grep -E "^train|worst|dp2" gpurun_out/parity_raw_errors.txt | cut -c1-700
