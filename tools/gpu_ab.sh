#!/bin/bash
# A/B of this session's changes: the re-measured tests, then bench lines (train, fwd, fwd --bf16) without the slow baseline legs.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -q -m gpu -k "bf16 or two_shard or train_step_raw" --timeout 600 -p no:cacheprovider > gpurun_out/ab_tests.log 2>&1
echo "== tests: exit $?"; tail -n 4 gpurun_out/ab_tests.log
run() { timeout 600 python bench.py "${@:2}" > gpurun_out/$1.json 2> gpurun_out/$1.err; echo "$1 exit $?"; cut -c1-200 gpurun_out/$1.json; }
run ab_train --no-eager-baseline --no-cpu-baseline
RF_CROP_DIRECT=1 run ab_train_direct_crop --no-eager-baseline --no-cpu-baseline
run ab_fwd --mode fwd --no-eager-baseline --no-cpu-baseline
run ab_fwd_bf16 --mode fwd --bf16 --no-eager-baseline --no-cpu-baseline
run ab_dreyeve_bf16 --mode dreyeve_sweep --bf16 --no-eager-baseline --no-cpu-baseline
