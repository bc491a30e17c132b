#!/bin/bash
# Launch list of ONE training step of the current build (NVTX range "timed_step" of `bench.py --profile`), its summaries, and
# --set full captures of two launches of the kernel families given as arguments.  usage: bash tools/gpu_ncu_launches.sh [families...]
mkdir -p gpurun_out
CMD="python bench.py --profile --steps 1 --no-graph"
NV='--nvtx --nvtx-include timed_step --kernel-name-base demangled'
$CMD > gpurun_out/plain.log 2>&1 && \
ncu $NV --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
python tools/ncu_summary.py gpurun_out/launches.csv > gpurun_out/launch_summary.txt
python tools/ncu_summary.py gpurun_out/launches.csv --by-grid > gpurun_out/launch_summary_by_grid.txt
head -n 12 gpurun_out/launch_summary.txt
for fam in "$@"; do
  tag=$(echo $fam | tr -d ':')
  ncu $NV --set full --import-source on --clock-control none -k regex:"$fam" -c 2 -f -o /tmp/prof_$tag $CMD > gpurun_out/ncu_full_$tag.log 2>&1
  echo "$fam capture exit $?"
  python tools/ncu_export.py /tmp/prof_$tag.ncu-rep > gpurun_out/ncu_full_$tag.txt
done
