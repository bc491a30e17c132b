#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_model.py -q -m gpu --timeout 600 -p no:cacheprovider > gpurun_out/model_tests.log 2>&1
echo "== model tests: exit $?" | tee -a gpurun_out/summary.txt
tail -n 40 gpurun_out/model_tests.log
