#!/bin/bash
# round 2, GPU call 14: epilogue hands the accumulator pair back at once, bias off the critical path
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_stages_gpu.py -m gpu -q -x -k "tcgen05 or packed" 2>&1 | tail -3
if [ "${PIPESTATUS[0]}" != "0" ]; then echo "stage tests failed or hung: stop"; exit 1; fi
for shape in "16384 14 256 768 0" "16384 14 1024 256 0"; do timeout 60 python tools/ss_stats.py $shape 2>&1 | tail -1; done | tee gpurun_out/r02n_ss_stats.txt
for shape in "16384 14 256 768" "16384 14 256 256" "16384 14 256 1024" "16384 14 1024 256" "6400 32 256 1024" "4682 44 256 1024"; do
  GEMM_PACKED=1 timeout 60 python tools/gemm_bench.py $shape 20 2>&1 | tail -1
done | tee gpurun_out/r02n_gemm_bench.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train-step --systems none > gpurun_out/r02n_bench_be.json 2> gpurun_out/r02n_bench_be.err
tail -c 300 gpurun_out/r02n_bench_be.err
python - <<'PY'
import json
b=json.load(open("gpurun_out/r02n_bench_be.json"))
print(b["value"], b["ms_per_step"], b["roofline"]["achieved"], b["roofline"]["frac"], {k:v["ms"] for k,v in b["kernel_breakdown"].items()})
PY
