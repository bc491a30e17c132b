#!/bin/bash
# round 2, call D: tensor-core attention forward bring-up (kernel tests first, then model tests + bench)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_raw_errors.txt
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -p no:cacheprovider -k "attention" -rP > gpurun_out/r2d_pytest_attn.log 2>&1
rc=$?; echo "attention pytest exit $rc"; grep -E "tcgen05 attention|passed|failed|Error|error" gpurun_out/r2d_pytest_attn.log | head -40
if [ $rc -ne 0 ]; then tail -60 gpurun_out/r2d_pytest_attn.log; fi
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x --deselect tests/test_gpu_kernels.py > gpurun_out/r2d_pytest.log 2>&1
echo "pytest exit $?"; tail -30 gpurun_out/r2d_pytest.log
timeout 600 python bench.py --no-eager-baseline --no-cpu-baseline > gpurun_out/r2d_bench_train.json 2> gpurun_out/r2d_bench_train.err; echo "train exit $?"; tail -c 1500 gpurun_out/r2d_bench_train.json; tail -5 gpurun_out/r2d_bench_train.err
RF_ATTN_TC=0 timeout 600 python bench.py --no-eager-baseline --no-cpu-baseline > gpurun_out/r2d_bench_train_notc.json 2> gpurun_out/r2d_bench_train_notc.err; echo "train (no tc) exit $?"; tail -c 600 gpurun_out/r2d_bench_train_notc.json
cat gpurun_out/parity_raw_errors.txt | cut -c1-400
