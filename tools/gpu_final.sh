#!/bin/bash
# round-end checks exactly as the driver runs them
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider 2>&1 | tail -5
timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/bench_default.log 2>&1; echo "bench exit $?"; grep "^{" gpurun_out/bench_default.log | cut -c1-300
timeout 600 python bench.py --impl reference > gpurun_out/bench_reference_default.log 2>&1; echo "ref exit $?"; tail -1 gpurun_out/bench_reference_default.log | cut -c1-300
