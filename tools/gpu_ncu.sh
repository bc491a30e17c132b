#!/bin/bash
# ncu evidence: (1) per-launch device times of one training step, (2) full capture of the dominant kernels, exported to text on
# the box (the .ncu-rep files are kept only while gpurun_out stays far below the 64 MiB that travels back).
mkdir -p gpurun_out
CMD="python bench.py --profile --steps 1 --no-graph"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --nvtx --nvtx-include "timed_step/" --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none -k regex:"gemm_tf32_persistent" -s 340 -c 20 -f -o /tmp/prof_gemm $CMD > gpurun_out/ncu_full_gemm.log 2>&1
echo "gemm capture exit $?"
python tools/ncu_export.py /tmp/prof_gemm.ncu-rep > gpurun_out/ncu_full_gemm.txt
$CMD > gpurun_out/plain3.log 2>&1 && \
ncu --set full --clock-control none -k regex:"attention_small|fov_crop|layernorm_fwd|layernorm_bwd|adamw|colsum|gemm_tf32_kernel" -s 100 -c 14 -f -o /tmp/prof_misc $CMD > gpurun_out/ncu_full_misc.log 2>&1
echo "misc capture exit $?"
python tools/ncu_export.py /tmp/prof_misc.ncu-rep > gpurun_out/ncu_full_misc.txt
ls -la /tmp/*.ncu-rep
for f in /tmp/prof_gemm.ncu-rep /tmp/prof_misc.ncu-rep; do [ $(stat -c %s $f) -lt 12000000 ] && cp $f gpurun_out/; done
du -sh gpurun_out
