#!/bin/bash
# round 2, GPU call 28: first-layer attention kernel: tests, timing alone, instruction count, Ne / N2 bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x -k "first_layer" > gpurun_out/r02ap_pytest_l0.log 2>&1; rc=$?; echo "pytest l0 rc=$rc"; grep -E "^E  .*assert|passed|failed" gpurun_out/r02ap_pytest_l0.log | head -8
python tools/afl_only.py 14 1702 5
python tools/afl_only.py 10 3276 5
python tools/afl_only.py 4 4096 5
timeout 200 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_write.sum --clock-control none -k regex:attention_first_layer -s 1 -c 1 python tools/afl_only.py 14 1702 2 2>&1 | grep -E "inst_executed|gpu__time|issue_active|dram__bytes"
if [ $rc -ne 0 ]; then exit 1; fi
timeout 600 python bench.py --systems Ne,N2 --no-cpu-baseline --no-train-step --steps 30 --warmup 5 > gpurun_out/r02ap_bench.json 2> gpurun_out/r02ap_bench.err; echo "bench rc=$?"
python - <<PY
import json
b=json.load(open("gpurun_out/r02ap_bench.json"))
print("Be", b["value"], b["ms_per_step"], b["e2e"]["value"], b["roofline"]["achieved"], b["clocks"]["sm_mhz"])
for n,s in b["systems"].items():
    print(n, s.get("evals_per_s"), s.get("ms_per_step"), s.get("kernel_ms"))
PY
