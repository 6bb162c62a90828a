#!/bin/bash
# GEMM iteration loop: stage tests, A/B timing of the four Be GEMM shapes, bench line, clock64 timeline of cluster 0
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_stages_gpu.py -m gpu -x -q -k "linear_tcgen05" -p no:cacheprovider 2>&1 | tail -3
for shape in "16384 14 256 768" "16384 14 256 256" "16384 14 256 1024" "16384 14 1024 256"; do
  timeout 120 python tools/gemm_bench.py $shape 20 2>&1 | tail -1
done
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_h.log 2>&1
tail -1 gpurun_out/bench_h.log | cut -c1-2500
timeout 100 python tools/trace_gemm2.py 256 768 0 14 2>&1 | tail -22
