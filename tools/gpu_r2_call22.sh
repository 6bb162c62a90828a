#!/bin/bash
# round 2, GPU call 22: the register-resident clamp fix-up kernel: clamp tests, then Ne / N2 / LiH with the fix-up off and on
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x -k "clamp or determinant or slogdet or logdet" > gpurun_out/r02af_pytest_clamp.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02af_pytest_clamp.log
for fx in 0 1; do
  PSIF_CLAMP_FIXUP=$fx timeout 600 python bench.py --systems LiH,Ne,N2 --no-cpu-baseline --no-train-step --steps 20 --warmup 3 > gpurun_out/r02af_bench_fix$fx.json 2> gpurun_out/r02af_bench_fix$fx.err; echo "bench fix=$fx rc=$?"
  python - <<PY
import json
b=json.load(open("gpurun_out/r02af_bench_fix$fx.json"))
print("fix=$fx Be", b["value"], b["ms_per_step"])
for n,s in b["systems"].items():
    print(n, s.get("evals_per_s"), s.get("ms_per_step"), s.get("kernel_ms",{}).get("det"))
PY
done
