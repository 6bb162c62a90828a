#!/bin/bash
# GELU epilogue A/B on one box: dbg 96 = single columns + single items (previous behaviour), 64, 32, 0 = new
for rep in 1 2; do for dbg in 96 64 32 0; do for shape in "16384 14 256 1024" "6400 32 256 1024"; do
  echo -n "dbg=$dbg "; PSIF_TC_EXPERIMENT=$dbg GEMM_PACKED=1 timeout 60 python tools/gemm_bench.py $shape 30 2>&1 | tail -1 | cut -c1-140
done; done; done
