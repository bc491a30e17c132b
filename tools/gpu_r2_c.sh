#!/bin/bash
# round 2, call C: kernel tests (new tiled crop), model parity tests, crop micro-benchmark, train bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_raw_errors.txt
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -p no:cacheprovider -k "fov" > gpurun_out/r2c_pytest_crop.log 2>&1
echo "crop pytest exit $?"; tail -15 gpurun_out/r2c_pytest_crop.log
timeout 300 python bench.py --mode crop_micro > gpurun_out/r2c_bench_crop.json 2> gpurun_out/r2c_bench_crop.err; echo "crop exit $?"; cat gpurun_out/r2c_bench_crop.json; tail -5 gpurun_out/r2c_bench_crop.err
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2c_pytest.log 2>&1
echo "pytest exit $?"; tail -30 gpurun_out/r2c_pytest.log
timeout 600 python bench.py --no-eager-baseline --no-cpu-baseline > gpurun_out/r2c_bench_train.json 2> gpurun_out/r2c_bench_train.err; echo "train exit $?"; tail -c 1500 gpurun_out/r2c_bench_train.json; tail -5 gpurun_out/r2c_bench_train.err
cat gpurun_out/parity_raw_errors.txt
