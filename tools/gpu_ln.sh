#!/bin/bash
# LayerNorm / optimiser kernels: parity tests, micro-benchmark lines, and the training step with the multi-row kernels off / on / off / on.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "layernorm or reductions or median_metrics" --timeout 300 -p no:cacheprovider 2>&1 | tail -n 3
timeout 600 python -m pytest tests/test_gpu_model.py -q -m gpu -k "cuda_graph or two_shard or training_step" --timeout 600 -p no:cacheprovider 2>&1 | tail -n 3
python tools/microbench.py 2>/dev/null | grep layernorm
RF_LN_ROWS=0 python tools/microbench.py 2>/dev/null | grep layernorm | sed 's/^/rows_off: /'
run() { timeout 600 python bench.py "${@:2}" > gpurun_out/$1.json 2> gpurun_out/$1.err; echo "$1 $(python -c "import json;d=json.loads(open(\"gpurun_out/$1.json\").read().strip().splitlines()[-1]);print(d[\"ms_per_step\"], d[\"e2e\"][\"value\"])")"; }
RF_LN_ROWS=0 run ln_off1 --no-eager-baseline --no-cpu-baseline
run ln_on1 --no-eager-baseline --no-cpu-baseline
RF_LN_ROWS=0 run ln_off2 --no-eager-baseline --no-cpu-baseline
run ln_on2 --no-eager-baseline --no-cpu-baseline
