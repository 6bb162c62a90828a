#!/bin/bash
# round 2, GPU call 6: step-gap diagnosis (guard sync / launch gaps / graph) + ncu of the packed GEMMs
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_stages_gpu.py -m gpu -q -k "gelu" 2>&1 | tail -3
for s in Be He Ne; do timeout 300 python tools/step_gaps.py $s 30 2>&1 | tail -1; done | tee gpurun_out/r02f_step_gaps.jsonl
GEMM_PACKED=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_2cta -s 4 -c 1 -f -o gpurun_out/r02f_gemm_pk_plain python tools/gemm_bench.py 16384 14 256 768 1 > gpurun_out/r02f_ncu_pk_plain.log 2>&1
GEMM_PACKED=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_2cta -s 8 -c 1 -f -o gpurun_out/r02f_gemm_pk_gelu python tools/gemm_bench.py 16384 14 256 1024 1 > gpurun_out/r02f_ncu_pk_gelu.log 2>&1
for f in r02f_gemm_pk_plain r02f_gemm_pk_gelu; do python profiles/ncu_summary.py gpurun_out/$f.ncu-rep > gpurun_out/$f.summary.txt 2>&1; done
ls -la gpurun_out/*.ncu-rep
