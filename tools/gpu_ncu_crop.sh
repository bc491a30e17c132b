#!/bin/bash
mkdir -p gpurun_out
python tools/crop_probe.py > gpurun_out/crop_probe.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"fov_crop" -s 2 -c 2 -f -o gpurun_out/prof_crop python tools/crop_probe.py > gpurun_out/ncu_crop.log 2>&1
echo "exit $?"; tail -2 gpurun_out/ncu_crop.log; ls -la gpurun_out/prof_crop.ncu-rep
