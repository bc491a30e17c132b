#!/bin/bash
mkdir -p gpurun_out
for lib in "" routeformer_b200/_lib/alt_old_attention.so routeformer_b200/_lib/alt_scalar_attention.so; do
  echo "== lib: ${lib:-current (FFMA2)}"
  if [ -z "$lib" ]; then unset RF_LIB_PATH; else export RF_LIB_PATH=$PWD/$lib; fi
  PYTHONPATH=. timeout 300 python tools/microbench.py > gpurun_out/mb_tmp.log 2>&1; grep -E "^attention" gpurun_out/mb_tmp.log || tail -5 gpurun_out/mb_tmp.log
done | tee gpurun_out/attn_ab.log
