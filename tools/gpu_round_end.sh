#!/bin/bash
# Round-end evidence in one call: all GPU tests, every bench mode, the crop ncu capture.
bash tools/gpu_tests.sh
bash tools/gpu_bench_modes.sh

