#!/bin/bash
# round 2, GPU call 12: KP2 (both K halves of FC2 in one launch): tests, timing A/B, Be line
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_stages_gpu.py -m gpu -q -x -k "tcgen05 or packed" 2>&1 | tail -6
if [ "${PIPESTATUS[0]}" != "0" ]; then echo "stage tests failed or hung: stop"; exit 1; fi
for kp2 in 0 1; do for shape in "16384 14 1024 256" "6400 32 1024 256"; do
  echo -n "kp2=$kp2 "; PSIF_TC_KP2=$kp2 GEMM_PACKED=1 timeout 60 python tools/gemm_bench.py $shape 20 2>&1 | tail -1
done; done | tee gpurun_out/r02l_gemm_kp2.txt
timeout 300 python -m pytest tests/test_stages_gpu.py tests/test_parity_gpu.py -m gpu -q --maxfail=5 2>&1 | tail -4
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train-step --systems none > gpurun_out/r02l_bench_be.json 2> gpurun_out/r02l_bench_be.err
tail -c 300 gpurun_out/r02l_bench_be.err
python - <<'PY'
import json
b=json.load(open("gpurun_out/r02l_bench_be.json"))
print(b["value"], b["ms_per_step"], b["roofline"]["achieved"], b["roofline"]["frac"], {k:v["ms"] for k,v in b["kernel_breakdown"].items()})
PY
