#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 300 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 40 -c 4 -f -o gpurun_out/prof_gemm_r01 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_gemm.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_payload -s 8 -c 2 -f -o gpurun_out/prof_att_r01 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_att.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"layernorm_payload|gelu_payload|det_combine" -s 24 -c 3 -f -o gpurun_out/prof_misc_r01 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_misc.log 2>&1
ls -la gpurun_out/
