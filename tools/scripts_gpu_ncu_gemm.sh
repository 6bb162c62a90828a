#!/bin/bash
# ncu --set full with source correlation of the fp16-split GEMM: plain (QKV shape) and fused payload-GELU (FC shape)
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_2cta -s 4 -c 1 -f -o gpurun_out/gemm_h_plain python tools/gemm_bench.py 16384 14 256 768 1 > gpurun_out/ncu_h_plain.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_2cta -s 8 -c 1 -f -o gpurun_out/gemm_h_gelu python tools/gemm_bench.py 16384 14 256 1024 1 > gpurun_out/ncu_h_gelu.log 2>&1
ls -la gpurun_out
