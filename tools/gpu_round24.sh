#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --profile --steps 1 --no-graph"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 880 -c 1200 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"; tail -1 gpurun_out/plain.log | cut -c1-200; grep -c '^"' gpurun_out/launches.csv
