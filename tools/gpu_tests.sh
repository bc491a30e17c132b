#!/bin/bash
# GPU parity tests, one pytest process per kernel family (a CUDA fault in one family must not poison the others), then the model tests.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu_info.txt 2>&1
for fam in gemm fov_crop conv3 attention layernorm distil "motion or decoder_input or stream_tokens or median or reductions or dropout"; do
  tag=$(echo "$fam" | cut -d' ' -f1)
  timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "$fam" --timeout 600 -p no:cacheprovider > gpurun_out/kernels_$tag.log 2>&1
  echo "== $tag: exit $?"; tail -n 2 gpurun_out/kernels_$tag.log
done
timeout 1800 python -m pytest tests/test_gpu_model.py -q -m gpu --timeout 900 -p no:cacheprovider > gpurun_out/model_tests.log 2>&1
echo "== model tests: exit $?"; tail -n 5 gpurun_out/model_tests.log
