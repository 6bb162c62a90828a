#!/bin/bash
# round 2, GPU call 5: bulk-copy attention ring + slim GELU epilogue: all GPU tests, GEMM A/B, Be / Ne / N2 timings
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=12 > gpurun_out/r02e_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -12 gpurun_out/r02e_pytest_gpu.log | cut -c1-300
for p in 0 1; do for shape in "16384 14 256 768" "16384 14 256 1024" "16384 14 1024 256" "4682 44 256 1024"; do
  GEMM_PACKED=$p timeout 120 python tools/gemm_bench.py $shape 20 2>&1 | tail -1
done; done | tee gpurun_out/r02e_gemm_bench.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train-step --systems none > gpurun_out/r02e_bench_be.json 2> gpurun_out/r02e_bench_be.err
tail -c 300 gpurun_out/r02e_bench_be.err
timeout 600 python bench.py --steps 5 --warmup 3 --system Ne --systems N2 --no-cpu-baseline --no-train-step > gpurun_out/r02e_bench_ne.json 2> gpurun_out/r02e_bench_ne.err
tail -c 300 gpurun_out/r02e_bench_ne.err
python - <<'PY'
import json
for f in ("gpurun_out/r02e_bench_be.json","gpurun_out/r02e_bench_ne.json"):
    try: b=json.load(open(f))
    except Exception as e: print(f, e); continue
    for n,s in b["systems"].items():
        print(n, {k:(round(v,1) if isinstance(v,float) else v) for k,v in s.items() if k!="kernel_ms"}, s.get("kernel_ms"))
PY
