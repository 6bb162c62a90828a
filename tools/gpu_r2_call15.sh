#!/bin/bash
# round 2, GPU call 15: full GPU suite, smoke, full default bench line (all five systems, train step, cpu baseline)
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_stages_gpu.py -m gpu -q -x -k "tcgen05 or packed" 2>&1 | tail -2
if [ "${PIPESTATUS[0]}" != "0" ]; then echo "stage tests failed or hung: stop"; exit 1; fi
timeout 600 python -m pytest tests -m gpu -q --maxfail=5 > gpurun_out/r02o_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02o_pytest_gpu.log
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r02o_bench.json 2> gpurun_out/r02o_bench.err; echo "bench rc=$?"
tail -c 300 gpurun_out/r02o_bench.err
python - <<'PY'
import json
b=json.load(open("gpurun_out/r02o_bench.json"))
print({k:b[k] for k in ("value","ms_per_step","gpu_launches","mh_walker_steps_per_s")}, b["e2e"]["value"], b["roofline"]["achieved"], b["roofline"]["frac"], b["clocks"], b["sustained"], b.get("train_step"), b.get("cpu_baseline"))
for n,s in b["systems"].items():
    print(n, {k:(round(v,1) if isinstance(v,float) else v) for k,v in s.items() if k!="kernel_ms"}, s.get("kernel_ms"))
PY
