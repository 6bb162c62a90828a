#!/bin/bash
# round 2, GPU call 9: SS GEMM with the three-phase payload-GELU epilogue: tests, GEMM timings, Be / Ne / N2 lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/r02i_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/r02i_pytest_gpu.log | cut -c1-300
for shape in "16384 14 256 768" "16384 14 256 256" "16384 14 256 1024" "16384 14 1024 256" "6400 32 256 1024" "4682 44 256 1024"; do
  GEMM_PACKED=1 timeout 120 python tools/gemm_bench.py $shape 20 2>&1 | tail -1
done | tee gpurun_out/r02i_gemm_bench.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train-step --systems Ne,N2 > gpurun_out/r02i_bench_be.json 2> gpurun_out/r02i_bench_be.err
tail -c 300 gpurun_out/r02i_bench_be.err
python - <<'PY'
import json
b=json.load(open("gpurun_out/r02i_bench_be.json"))
print(b["value"], b["ms_per_step"], b["roofline"]["achieved"], b["roofline"]["frac"], b["sustained"])
for n,s in b["systems"].items():
    print(n, {k:(round(v,1) if isinstance(v,float) else v) for k,v in s.items() if k!="kernel_ms"}, s.get("kernel_ms"))
PY
GEMM_PACKED=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_ss -s 8 -c 1 -f -o gpurun_out/r02i_gemm_ss_gelu python tools/gemm_bench.py 16384 14 256 1024 1 > gpurun_out/r02i_ncu_ss_gelu.log 2>&1
python profiles/ncu_summary.py gpurun_out/r02i_gemm_ss_gelu.ncu-rep > gpurun_out/r02i_gemm_ss_gelu.summary.txt 2>&1
