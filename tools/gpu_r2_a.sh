#!/bin/bash
# round 2, call A: full GPU test suite (new parity tests), then every bench mode once
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_raw_errors.txt
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2b_pytest.log 2>&1
echo "pytest exit $?"; tail -15 gpurun_out/r2b_pytest.log
timeout 600 python bench.py > gpurun_out/r2b_bench_train.json 2> gpurun_out/r2b_bench_train.err; echo "train exit $?"; tail -c 3000 gpurun_out/r2b_bench_train.json; tail -5 gpurun_out/r2b_bench_train.err
timeout 300 python bench.py --mode fwd > gpurun_out/r2b_bench_fwd.json 2> gpurun_out/r2b_bench_fwd.err; echo "fwd exit $?"; cat gpurun_out/r2b_bench_fwd.json; tail -5 gpurun_out/r2b_bench_fwd.err
timeout 600 python bench.py --mode dreyeve_sweep > gpurun_out/r2b_bench_dreyeve.json 2> gpurun_out/r2b_bench_dreyeve.err; echo "dreyeve exit $?"; cat gpurun_out/r2b_bench_dreyeve.json; tail -5 gpurun_out/r2b_bench_dreyeve.err
timeout 300 python bench.py --mode crop_micro > gpurun_out/r2b_bench_crop.json 2> gpurun_out/r2b_bench_crop.err; echo "crop exit $?"; cat gpurun_out/r2b_bench_crop.json; tail -5 gpurun_out/r2b_bench_crop.err
cat gpurun_out/parity_raw_errors.txt
