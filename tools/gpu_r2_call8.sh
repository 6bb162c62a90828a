#!/bin/bash
# round 2, GPU call 8: what bounds the packed-operand SS GEMM: pipeline parts taken away one at a time + ncu
mkdir -p gpurun_out
for dbg in 0 1 8 2 4 6 10 14; do for shape in "16384 14 256 768" "16384 14 1024 256"; do
  echo -n "dbg=$dbg  "; PSIF_TC_EXPERIMENT=$dbg GEMM_PACKED=1 timeout 120 python tools/gemm_bench.py $shape 20 2>&1 | tail -1
done; done | tee gpurun_out/r02h_gemm_parts.txt
GEMM_PACKED=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_ss -s 4 -c 1 -f -o gpurun_out/r02h_gemm_ss_plain python tools/gemm_bench.py 16384 14 256 768 1 > gpurun_out/r02h_ncu_ss_plain.log 2>&1
GEMM_PACKED=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_ss -s 8 -c 1 -f -o gpurun_out/r02h_gemm_ss_gelu python tools/gemm_bench.py 16384 14 256 1024 1 > gpurun_out/r02h_ncu_ss_gelu.log 2>&1
for f in r02h_gemm_ss_plain r02h_gemm_ss_gelu; do python profiles/ncu_summary.py gpurun_out/$f.ncu-rep > gpurun_out/$f.summary.txt 2>&1; done
ls -la gpurun_out/*.ncu-rep
