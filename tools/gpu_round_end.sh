#!/bin/bash
# The round-end checks in the order the driver runs them: GPU tests, build()+smoke(), the reference arm, the default bench.
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 3 gpurun_out/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -n 2 gpurun_out/smoke.log
timeout 600 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference arm exit $?"; cut -c1-300 gpurun_out/bench_reference.json
timeout 900 python bench.py > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; echo "bench exit $?"; cut -c1-400 gpurun_out/bench_train.json
