#!/bin/bash
# Round-end evidence in one call: all GPU tests, every bench mode, the crop ncu capture.
bash tools/gpu_tests.sh
bash tools/gpu_bench_modes.sh
bash tools/gpu_crop_ncu.sh > gpurun_out/crop_ncu_stdout.txt 2>&1; tail -n 3 gpurun_out/crop_ncu_stdout.txt
