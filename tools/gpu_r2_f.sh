#!/bin/bash
# round 2, call F: graph + dropout tests, full suite, paper-dropout bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_raw_errors.txt
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q --tb=short -p no:cacheprovider -k "graph or dropout or caller" > gpurun_out/r2f_pytest_graph.log 2>&1
echo "graph pytest exit $?"; tail -40 gpurun_out/r2f_pytest_graph.log | cut -c1-600
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2f_pytest.log 2>&1
echo "pytest exit $?"; tail -12 gpurun_out/r2f_pytest.log | cut -c1-400
timeout 600 python bench.py --no-eager-baseline --no-cpu-baseline --paper-dropout --steps 100 > gpurun_out/r2f_bench_train_paper_dropout.json 2> gpurun_out/r2f_bench_train_paper_dropout.err; echo "train paper-dropout exit $?"; tail -c 1200 gpurun_out/r2f_bench_train_paper_dropout.json; tail -5 gpurun_out/r2f_bench_train_paper_dropout.err
timeout 600 python bench.py --no-eager-baseline --no-cpu-baseline > gpurun_out/r2f_bench_train.json 2> gpurun_out/r2f_bench_train.err; echo "train exit $?"; tail -c 500 gpurun_out/r2f_bench_train.json; tail -5 gpurun_out/r2f_bench_train.err
