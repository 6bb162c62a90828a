#!/bin/bash
# round 2, GPU call 7: packed-operand SS GEMM (double-buffered accumulators): stage tests, GEMM A/B, Be bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_stages_gpu.py -m gpu -q -x -k "tcgen05 or packed" 2>&1 | tail -15
for ss in 0 1; do for shape in "16384 14 256 768" "16384 14 256 256" "16384 14 256 1024" "16384 14 1024 256" "4682 44 256 1024"; do
  PSIF_TC_SS=$ss GEMM_PACKED=1 timeout 120 python tools/gemm_bench.py $shape 20 2>&1 | tail -1
done; done | tee gpurun_out/r02g_gemm_bench.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train-step --systems none > gpurun_out/r02g_bench_be.json 2> gpurun_out/r02g_bench_be.err
tail -c 300 gpurun_out/r02g_bench_be.err
python - <<'PY'
import json
b=json.load(open("gpurun_out/r02g_bench_be.json"))
print(b["value"], b["ms_per_step"], b["roofline"]["achieved"], b["roofline"]["frac"], {k:v["ms"] for k,v in b["kernel_breakdown"].items()})
PY
