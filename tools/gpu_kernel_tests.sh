#!/bin/bash
# Runs the per-kernel GPU parity tests, one pytest process per kernel family (a CUDA fault in one family
# must not poison the context of the others).  Logs go to gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu_info.txt 2>&1
for fam in gemm fov_crop conv3 attention layernorm distil "motion or decoder_input or stream_tokens or median or reductions"; do
  tag=$(echo "$fam" | cut -d' ' -f1)
  timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "$fam" --timeout 300 -p no:cacheprovider > gpurun_out/kernels_$tag.log 2>&1
  echo "== $tag: exit $?" | tee -a gpurun_out/summary.txt
  tail -n 3 gpurun_out/kernels_$tag.log | tee -a gpurun_out/summary.txt
done
