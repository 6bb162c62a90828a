#!/bin/bash
# round 2, GPU call 25: first-layer compact payload (4-electron systems): equality test, GPU suite, Be / LiH A/B
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x -k "first_layer" > gpurun_out/r02ak_pytest_l0.log 2>&1; rc=$?; echo "pytest l0 rc=$rc"; tail -5 gpurun_out/r02ak_pytest_l0.log
if [ $rc -ne 0 ]; then exit 1; fi
for sp in 0 1; do
  PSIF_L0_SPARSE=$sp timeout 600 python bench.py --systems LiH --no-cpu-baseline --no-train-step --steps 50 --warmup 5 > gpurun_out/r02ak_bench_sp$sp.json 2> gpurun_out/r02ak_bench_sp$sp.err; echo "bench sp=$sp rc=$?"
  python - <<PY
import json
b=json.load(open("gpurun_out/r02ak_bench_sp$sp.json"))
print("sp=$sp Be", b["value"], b["ms_per_step"], b["e2e"]["value"], b["roofline"]["achieved"], b["clocks"]["sm_mhz"])
for n,s in b["systems"].items():
    print(n, s.get("evals_per_s"), s.get("ms_per_step"), s.get("kernel_ms"))
PY
done
timeout 600 python -m pytest tests -m gpu -q --maxfail=5 > gpurun_out/r02ak_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02ak_pytest_gpu.log
