#!/bin/bash
mkdir -p gpurun_out
python tools/attn_probe.py > gpurun_out/attn_probe.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"attention_small" -s 2 -c 2 -f -o gpurun_out/prof_attn python tools/attn_probe.py > gpurun_out/ncu_attn.log 2>&1
echo "exit $?"; tail -3 gpurun_out/ncu_attn.log
