#!/bin/bash
# round-1 refresh: GPU tests, bench line, ncu launch list of the same command, full captures of the top kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_n1.log 2>&1
tail -1 gpurun_out/bench_n1.log | cut -c1-3000
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1
tail -1 gpurun_out/bench_ref.log | cut -c1-600
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01b.csv python bench.py --steps 2 --warmup 3 --profile-mode > gpurun_out/bench_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 30 -c 15 -f -o gpurun_out/prof_gemm_r01b python bench.py --steps 2 --warmup 3 --profile-mode > gpurun_out/ncu_gemm.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attention_payload" -s 8 -c 2 -f -o gpurun_out/prof_att_r01b python bench.py --steps 2 --warmup 3 --profile-mode > gpurun_out/ncu_att.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"layernorm_payload|gelu_payload|det_combine|orbital_envelope|embed_kernel|gemm_tn_ffma" -s 30 -c 8 -f -o gpurun_out/prof_misc_r01b python bench.py --steps 2 --warmup 3 --profile-mode > gpurun_out/ncu_misc.log 2>&1
ls -la gpurun_out/
