#!/bin/bash
# ncu --set full of the FoV-crop kernels on the micro-benchmark's problems (tools/crop_probe.py), exported to text with per-line shares.
mkdir -p gpurun_out
python tools/crop_probe.py > gpurun_out/crop_probe.log 2>&1 || { echo "probe failed"; tail gpurun_out/crop_probe.log; exit 1; }
ncu --set full --import-source on --clock-control none -k regex:"fov_crop" -s 3 -c 3 -f -o gpurun_out/prof_crop python tools/crop_probe.py > gpurun_out/ncu_crop.log 2>&1
echo "capture exit $?"
python tools/ncu_export.py gpurun_out/prof_crop.ncu-rep > gpurun_out/ncu_full_fov_crop.txt
python tools/ncu_lines.py gpurun_out/prof_crop.ncu-rep fov_crop 1.0 >> gpurun_out/ncu_full_fov_crop.txt
head -50 gpurun_out/ncu_full_fov_crop.txt
python tools/ncu_stalls.py gpurun_out/prof_crop.ncu-rep fov_crop 14 > gpurun_out/ncu_stalls_fov_crop.txt; cat gpurun_out/ncu_stalls_fov_crop.txt
