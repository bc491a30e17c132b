#!/bin/bash
# ncu evidence: (1) per-launch device times of one training step, (2) full capture of the dominant kernels.
mkdir -p gpurun_out
CMD="python bench.py --profile --steps 1 --no-graph"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 1000 -c 1300 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"gemm_tf32|attention_fwd|attention_bwd|fov_crop" -s 40 -c 14 -o gpurun_out/prof_top $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $?"
tail -n 2 gpurun_out/plain.log; tail -n 2 gpurun_out/ncu_launches.log; tail -n 3 gpurun_out/ncu_full.log
