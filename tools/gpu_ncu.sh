#!/bin/bash
# ncu evidence for one training step (1 GPU).  Every ncu pass is preceded by the same command run plain (it must exit 0 first),
# and is restricted to the NVTX start/end range "timed_step" that `bench.py --profile` opens around the one timed step.
#   (1) per-launch device times of the step  -> gpurun_out/launches.csv  (tools/ncu_summary.py turns it into the summaries)
#   (2) --set full of 24 persistent-GEMM launches -> gpurun_out/ncu_full_gemm.txt
#   (3) --set full of two launches of every other kernel family -> gpurun_out/ncu_full_<family>.txt
#   (4) source-level captures of the GEMM probe and of the opt-in tcgen05 attention kernel (.ncu-rep kept: small)
# The reports are exported to text on the box (tools/ncu_export.py).  usage: bash tools/gpu_ncu.sh [families...]
mkdir -p gpurun_out
CMD="python bench.py --profile --steps 1 --no-graph"
NV='--nvtx --nvtx-include timed_step --kernel-name-base demangled'
FAMS=${@:-"attention_small_fwd attention_small_bwd wide::attention_fwd wide::attention_bwd narrow::attention_fwd narrow::attention_bwd fov_crop_kernel layernorm_fwd layernorm_bwd adamw colsum gemm_tf32_kernel"}
$CMD > gpurun_out/plain.log 2>&1 && \
ncu $NV --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
python tools/ncu_summary.py gpurun_out/launches.csv > gpurun_out/launch_summary.txt
python tools/ncu_summary.py gpurun_out/launches.csv --by-grid > gpurun_out/launch_summary_by_grid.txt
head -n 3 gpurun_out/launch_summary.txt
if [ -z "$SKIP_GEMM" ]; then
ncu $NV --set full --import-source on --clock-control none -k regex:"gemm_tf32_persistent" -s 60 -c 24 -f -o /tmp/prof_gemm $CMD > gpurun_out/ncu_full_gemm.log 2>&1
echo "gemm capture exit $?"
python tools/ncu_export.py /tmp/prof_gemm.ncu-rep > gpurun_out/ncu_full_gemm.txt
fi
for fam in $FAMS; do
  tag=$(echo $fam | tr -d ':')
  ncu $NV --set full --import-source on --clock-control none -k regex:"$fam" -c 2 -f -o /tmp/prof_$tag $CMD > gpurun_out/ncu_full_$tag.log 2>&1
  echo "$fam capture exit $?"
  python tools/ncu_export.py /tmp/prof_$tag.ncu-rep > gpurun_out/ncu_full_$tag.txt
done
python tools/gemm_probe.py > gpurun_out/gemm_probe.txt 2>&1 && \
ncu --set full --import-source on --clock-control none -k regex:"gemm_tf32_persistent" -f -o gpurun_out/prof_gemm_probe python tools/gemm_probe.py --ncu > gpurun_out/ncu_gemm_probe.log 2>&1
echo "gemm probe capture exit $?"; cat gpurun_out/gemm_probe.txt
if [ -z "$SKIP_TC" ]; then
python tools/attn_bench.py > gpurun_out/attn_bench.log 2>&1 && \
ncu --set full --import-source on --clock-control none -k regex:"attention_tc_fwd" -s 3 -c 2 -f -o gpurun_out/prof_attn_tc python tools/attn_bench.py > gpurun_out/ncu_full_attn_tc.log 2>&1
echo "attention_tc capture exit $?"
python tools/ncu_export.py gpurun_out/prof_attn_tc.ncu-rep > gpurun_out/ncu_full_attention_tc.txt
fi
du -sh gpurun_out
