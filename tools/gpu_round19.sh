#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu --timeout 300 -p no:cacheprovider -k "dropout" 2>&1 | tail -25
