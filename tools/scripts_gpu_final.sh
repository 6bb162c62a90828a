#!/bin/bash
# round-1 evidence: GPU tests, both bench arms, per-system table, ncu launch list + full captures (summarised on the box)
mkdir -p gpurun_out /tmp/rep
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log | cut -c1-200
timeout 600 python bench.py > gpurun_out/bench_n1.log 2>&1
tail -1 gpurun_out/bench_n1.log | cut -c1-3000
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1
tail -1 gpurun_out/bench_ref.log | cut -c1-300
timeout 900 python tools/bench_configs.py > gpurun_out/configs_n1.jsonl 2> gpurun_out/configs_err.log
cat gpurun_out/configs_n1.jsonl | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01c.csv python bench.py --steps 2 --warmup 3 --profile-mode > gpurun_out/bench_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 32 -c 6 -f -o /tmp/rep/gemm python bench.py --steps 2 --warmup 3 --profile-mode > gpurun_out/ncu_gemm.log 2>&1
python profiles/ncu_summary.py /tmp/rep/gemm.ncu-rep > gpurun_out/ncu_gemm_r01c_summary.txt
timeout 600 ncu --set full --clock-control none -k regex:"layernorm_payload|attention_payload" -s 20 -c 3 -f -o /tmp/rep/misc python bench.py --steps 2 --warmup 3 --profile-mode > gpurun_out/ncu_misc.log 2>&1
python profiles/ncu_summary.py /tmp/rep/misc.ncu-rep > gpurun_out/ncu_misc_r01c_summary.txt
ls -la gpurun_out/
