#!/bin/bash
# round 2, GPU call 19: evidence of the current build: GPU suite, smoke, default bench line, ncu launch list of the bench
# command, ncu --set full of one layer's GEMMs + attention + LayerNorm from the same command
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --maxfail=5 > gpurun_out/r02aq_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02aq_pytest_gpu.log
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1 | tee gpurun_out/r02aq_smoke.log
timeout 900 python bench.py > gpurun_out/r02aq_bench.json 2> gpurun_out/r02aq_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02aq_bench_reference.json 2>> gpurun_out/r02aq_bench.err; echo "reference arm rc=$?"
python - <<'PY'
import json
b=json.load(open("gpurun_out/r02aq_bench.json"))
print({k:b[k] for k in ("value","ms_per_step","gpu_launches","mh_walker_steps_per_s")}, b["e2e"]["value"], b["roofline"]["achieved"], b["roofline"]["frac"], b["clocks"], b["sustained"]["value"], b.get("train_step"), b.get("cpu_baseline",{}).get("value"))
for n,s in b["systems"].items():
    print(n, {k:(round(v,1) if isinstance(v,float) else v) for k,v in s.items() if k!="kernel_ms"}, s.get("kernel_ms"))
print(open("gpurun_out/r02aq_bench_reference.json").read()[:600])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02aq_launches_step.csv python bench.py --profile-mode --steps 2 --warmup 1 > gpurun_out/r02aq_ncu_launches.log 2>&1
timeout 400 ncu --set full --clock-control none -k regex:'tc_gemm|attention_payload|layernorm_payload' -s 7 -c 7 -f -o gpurun_out/r02aq_layer python bench.py --profile-mode --steps 1 --warmup 1 > gpurun_out/r02aq_ncu_layer.log 2>&1
python profiles/ncu_summary.py gpurun_out/r02aq_layer.ncu-rep > gpurun_out/r02aq_layer.summary.txt 2>&1
grep -E "^---|gpu__time|dram__bytes|tensor_cycles" gpurun_out/r02aq_layer.summary.txt
rm -f gpurun_out/r02aq_layer.ncu-rep
# layer 0 works on the compact payload: LayerNorm-1, QKV GEMM, attention
timeout 300 ncu --set full --clock-control none -k regex:'tc_gemm|attention_payload|attention_first_layer|layernorm_payload' -s 0 -c 3 -f -o gpurun_out/r02aq_layer0 python bench.py --profile-mode --steps 1 --warmup 1 > gpurun_out/r02aq_ncu_layer0.log 2>&1
python profiles/ncu_summary.py gpurun_out/r02aq_layer0.ncu-rep > gpurun_out/r02aq_layer0.summary.txt 2>&1
grep -E "^---|gpu__time|dram__bytes|tensor_cycles" gpurun_out/r02aq_layer0.summary.txt
rm -f gpurun_out/r02aq_layer0.ncu-rep
