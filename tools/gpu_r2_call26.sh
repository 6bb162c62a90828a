#!/bin/bash
# round 2, GPU call 26: first-layer attention kernel: stage test, pipeline test, then all systems with PSIF_L0_SPARSE=0/1
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x -k "first_layer" > gpurun_out/r02am_pytest_l0.log 2>&1; rc=$?; echo "pytest l0 rc=$rc"; grep -E "^E  .*assert|passed|failed" gpurun_out/r02am_pytest_l0.log | head -8; tail -3 gpurun_out/r02am_pytest_l0.log
echo continuing
for sp in 0 1; do
  PSIF_L0_SPARSE=$sp timeout 600 python bench.py --systems LiH,Ne,N2 --no-cpu-baseline --no-train-step --steps 30 --warmup 5 > gpurun_out/r02am_bench_sp$sp.json 2> gpurun_out/r02am_bench_sp$sp.err; echo "bench sp=$sp rc=$?"
  python - <<PY
import json
b=json.load(open("gpurun_out/r02am_bench_sp$sp.json"))
print("sp=$sp Be", b["value"], b["ms_per_step"], b["e2e"]["value"], b["roofline"]["achieved"], b["clocks"]["sm_mhz"])
for n,s in b["systems"].items():
    print(n, s.get("evals_per_s"), s.get("ms_per_step"), s.get("kernel_ms"))
PY
done
PSIF_L0_N4=0 timeout 600 python bench.py --systems none --no-cpu-baseline --no-train-step --steps 30 --warmup 5 > gpurun_out/r02am_bench_n4.json 2> gpurun_out/r02am_bench_n4.err
python - <<PY
import json
b=json.load(open("gpurun_out/r02am_bench_n4.json"))
print("n4 Be", b["value"], b["ms_per_step"], b["clocks"]["sm_mhz"])
PY
