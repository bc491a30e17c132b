#!/bin/bash
# Round-end evidence in one call: all GPU tests (one process per kernel family), then every bench mode.
bash tools/gpu_tests.sh
bash tools/gpu_bench_modes.sh

