#!/bin/bash
# round 2, GPU call 13: L2 prefetch of the next tile's X rows + wait statistics of the packed-operand GEMM
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_stages_gpu.py -m gpu -q -x -k "tcgen05 or packed" 2>&1 | tail -3
if [ "${PIPESTATUS[0]}" != "0" ]; then echo "stage tests failed or hung: stop"; exit 1; fi
for dbg in 16 0; do
  for shape in "16384 14 256 768 0" "16384 14 1024 256 0"; do
    echo -n "dbg=$dbg "; PSIF_TC_EXPERIMENT=$dbg timeout 60 python tools/ss_stats.py $shape 2>&1 | tail -1
  done
  for shape in "16384 14 256 768" "16384 14 256 256" "16384 14 256 1024" "16384 14 1024 256"; do
    echo -n "dbg=$dbg "; PSIF_TC_EXPERIMENT=$dbg GEMM_PACKED=1 timeout 60 python tools/gemm_bench.py $shape 20 2>&1 | tail -1
  done
done | tee gpurun_out/r02m_gemm_prefetch.txt
