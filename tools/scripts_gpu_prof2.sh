#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_ts -s 48 -c 4 -f -o gpurun_out/prof_gemm_ts_r01 python bench.py --steps 2 --warmup 3 --profile-mode > gpurun_out/ncu_gemm.log 2>&1
ls -la gpurun_out/*.ncu-rep
