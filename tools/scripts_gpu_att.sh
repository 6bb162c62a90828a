#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_att2.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 113 -c 37 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 3 --profile-mode > gpurun_out/bench_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 48 -c 4 -f -o gpurun_out/prof_gemm_r01 python bench.py --steps 2 --warmup 3 --profile-mode > gpurun_out/ncu_gemm.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_payload -s 12 -c 1 -f -o gpurun_out/prof_att_r01 python bench.py --steps 2 --warmup 3 --profile-mode > gpurun_out/ncu_att.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"layernorm_payload|gelu_payload|det_combine" -s 36 -c 4 -f -o gpurun_out/prof_misc_r01 python bench.py --steps 2 --warmup 3 --profile-mode > gpurun_out/ncu_misc.log 2>&1
grep -E "^\[.*\] (sample )?\|E_L|^FAILED|passed|failed" gpurun_out/pytest_gpu.log
python - <<PY
import json
l=[x for x in open("gpurun_out/bench_att2.log") if x.startswith("{")]
d=json.loads(l[-1]); print(d["value"], d["ms_per_step"], d["roofline"]["achieved"], d["mh_walker_steps_per_s"], {k:v["ms"] for k,v in d["kernel_breakdown"].items()})
PY
