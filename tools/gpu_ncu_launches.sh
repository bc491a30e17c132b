#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --profile --steps 1 --no-graph"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --nvtx --nvtx-include "timed_step/" --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
tail -n 2 gpurun_out/plain.log | cut -c1-300; tail -n 2 gpurun_out/ncu_launches.log | cut -c1-300
