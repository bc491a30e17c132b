#!/bin/bash
# Every bench.py mode at N=1, one JSON line each under gpurun_out/ (copied into profiles/ by hand once read).
mkdir -p gpurun_out
run() { timeout 900 python bench.py "${@:2}" > gpurun_out/$1.json 2> gpurun_out/$1.err; echo "$1 exit $?"; cut -c1-260 gpurun_out/$1.json; }
run bench_train                                   # default: training step, parity configuration, with the eager-GPU and CPU baselines
run bench_train_paper_dropout --paper-dropout --no-eager-baseline --no-cpu-baseline
run bench_train_u8_frames --u8-frames --no-eager-baseline --no-cpu-baseline
run bench_fwd --mode fwd --no-eager-baseline --no-cpu-baseline
run bench_fwd_bf16 --mode fwd --bf16 --no-eager-baseline --no-cpu-baseline
run bench_eval_step --mode eval_step
run bench_dreyeve_sweep --mode dreyeve_sweep --no-eager-baseline --no-cpu-baseline
run bench_crop_micro --mode crop_micro
run bench_stage_micro --mode stage_micro
run bench_reference --impl reference
timeout 600 python tools/microbench.py > gpurun_out/microbench.txt 2>&1; echo "microbench exit $?"
timeout 300 python tools/attn_bench.py > gpurun_out/attn_bench.txt 2>&1; timeout 300 python tools/attn_bench.py --generic >> gpurun_out/attn_bench.txt 2>&1; echo "attn_bench exit $?"
