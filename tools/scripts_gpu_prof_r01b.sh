#!/bin/bash
# ncu launch list of the bench command + full captures of the top kernels, summarised on the box (the reports stay there:
# gpurun_out/ is limited to 64 MiB)
mkdir -p gpurun_out /tmp/rep
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01b.csv python bench.py --steps 2 --warmup 3 --profile-mode > gpurun_out/bench_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 30 -c 6 -f -o /tmp/rep/gemm python bench.py --steps 2 --warmup 3 --profile-mode > gpurun_out/ncu_gemm.log 2>&1
python profiles/ncu_summary.py /tmp/rep/gemm.ncu-rep > gpurun_out/ncu_gemm_r01b_summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attention_payload" -s 8 -c 1 -f -o /tmp/rep/att python bench.py --steps 2 --warmup 3 --profile-mode > gpurun_out/ncu_att.log 2>&1
python profiles/ncu_summary.py /tmp/rep/att.ncu-rep > gpurun_out/ncu_attention_r01b_summary.txt
timeout 600 ncu --set full --clock-control none -k regex:"layernorm_payload|det_combine|orbital_envelope|embed_kernel|gemm_tn_ffma|jastrow" -s 20 -c 7 -f -o /tmp/rep/misc python bench.py --steps 2 --warmup 3 --profile-mode > gpurun_out/ncu_misc.log 2>&1
python profiles/ncu_summary.py /tmp/rep/misc.ncu-rep > gpurun_out/ncu_misc_r01b_summary.txt
ls -la gpurun_out/ /tmp/rep
