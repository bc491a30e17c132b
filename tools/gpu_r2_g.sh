#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/debug_capture.py 2>&1 | tail -8
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q --tb=short -p no:cacheprovider -k "graph" 2>&1 | tail -15 | cut -c1-300
