#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export RF_CROP_TILED=1
python bench.py --mode crop_micro > gpurun_out/ncu_crop_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fov_crop_tiled -s 2 -c 2 -o gpurun_out/r2_crop_strip python bench.py --mode crop_micro > gpurun_out/ncu_crop.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_crop.log
