#!/bin/bash
# LayerNorm: parity tests, then the micro-benchmark lines with the multi-row forward kernel on / off, then the training step A/B.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "layernorm" --timeout 300 -p no:cacheprovider 2>&1 | tail -n 3
python tools/microbench.py 2>/dev/null | grep layernorm
RF_LN_ROWS=0 python tools/microbench.py 2>/dev/null | grep layernorm | sed 's/^/rows_off: /'
run() { timeout 600 python bench.py "${@:2}" > gpurun_out/$1.json 2> gpurun_out/$1.err; echo "$1 exit $?"; cut -c1-160 gpurun_out/$1.json; }
run ln_train --no-eager-baseline --no-cpu-baseline
RF_LN_ROWS=0 run ln_train_rows_off --no-eager-baseline --no-cpu-baseline
