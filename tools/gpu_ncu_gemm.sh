#!/bin/bash
# final persistent-GEMM evidence: epilogue cost table, then a full ncu capture of the same launches (exported to text)
mkdir -p gpurun_out
PYTHONPATH=. timeout 200 python tools/gemm_epi_bench.py 2>&1 | tee gpurun_out/gemm_epi_v4.log
PYTHONPATH=. ncu --set full --clock-control none -k regex:"gemm_tf32_persistent" -s 120 -c 15 -f -o /tmp/prof_gemm2 python tools/gemm_epi_bench.py > gpurun_out/ncu_gemm2.log 2>&1
echo "ncu exit $?"; python tools/ncu_export.py /tmp/prof_gemm2.ncu-rep > gpurun_out/ncu_full_gemm_v10.txt; ls -la /tmp/prof_gemm2.ncu-rep; grep -c "^void" gpurun_out/ncu_full_gemm_v10.txt
