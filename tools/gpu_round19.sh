#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "reductions or attention" --timeout 300 -p no:cacheprovider 2>&1 | tail -15
timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu --timeout 300 -p no:cacheprovider 2>&1 | tail -25
