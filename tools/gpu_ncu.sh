#!/bin/bash
# ncu evidence for one training step (1 GPU).  Every ncu pass is preceded by the same command run plain (it must exit 0 first),
# and is restricted to the NVTX range "timed_step" that `bench.py --profile` pushes around the one timed step.
#   (1) per-launch device times of the step  -> gpurun_out/launches.csv  (tools/ncu_summary.py turns it into the summary)
#   (2) --set full of 24 persistent-GEMM launches -> gpurun_out/ncu_full_gemm.txt
#   (3) --set full of two launches of every other kernel family -> gpurun_out/ncu_full_<family>.txt
# The reports are exported to text on the box (tools/ncu_export.py); the .ncu-rep files stay there.
mkdir -p gpurun_out
CMD="python bench.py --profile --steps 1 --no-graph"
NV='--nvtx --nvtx-include timed_step/'
$CMD > gpurun_out/plain.log 2>&1 && \
ncu $NV --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
python tools/ncu_summary.py gpurun_out/launches.csv > gpurun_out/launch_summary.txt
python tools/ncu_summary.py gpurun_out/launches.csv --by-grid > gpurun_out/launch_summary_by_grid.txt
ncu $NV --set full --import-source on --clock-control none -k regex:"gemm_tf32_persistent" -s 60 -c 24 -f -o /tmp/prof_gemm $CMD > gpurun_out/ncu_full_gemm.log 2>&1
echo "gemm capture exit $?"
python tools/ncu_export.py /tmp/prof_gemm.ncu-rep > gpurun_out/ncu_full_gemm.txt
for fam in attention_small_fwd attention_small_bwd "wide::attention_fwd" "wide::attention_bwd" "narrow::attention_fwd" "narrow::attention_bwd" \
           fov_crop_kernel layernorm adamw colsum gemm_tf32_kernel; do
  tag=$(echo $fam | tr -d ':')
  ncu $NV --set full --import-source on --clock-control none -k regex:"$fam" -c 2 -f -o /tmp/prof_$tag $CMD > gpurun_out/ncu_full_$tag.log 2>&1
  echo "$fam capture exit $?"
  python tools/ncu_export.py /tmp/prof_$tag.ncu-rep > gpurun_out/ncu_full_$tag.txt
done
# the opt-in tcgen05 attention kernel, on the frame-encoder problem (micro-benchmark process, kernel filtered by name)
python tools/attn_bench.py > gpurun_out/attn_bench.log 2>&1 && \
ncu --set full --import-source on --clock-control none -k regex:"attention_tc_fwd" -s 3 -c 2 -f -o /tmp/prof_attn_tc python tools/attn_bench.py > gpurun_out/ncu_full_attn_tc.log 2>&1
echo "attention_tc capture exit $?"
python tools/ncu_export.py /tmp/prof_attn_tc.ncu-rep > gpurun_out/ncu_full_attention_tc.txt
ls -la /tmp/*.ncu-rep
for f in /tmp/prof_gemm.ncu-rep /tmp/prof_attn_tc.ncu-rep /tmp/prof_attention_small_fwd.ncu-rep; do [ $(stat -c %s $f) -lt 12000000 ] && cp $f gpurun_out/; done
du -sh gpurun_out
