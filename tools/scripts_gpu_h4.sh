#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_stages_gpu.py -m gpu -x -q -k "linear_tcgen05" -p no:cacheprovider 2>&1 | tail -2
for a in 1 0; do PSIF_TC_TPT_ALIGN=$a timeout 120 python tools/gemm_bench.py 16384 14 256 1024 20 2>&1 | tail -1 | sed "s/^/[align4=$a] /"; done
for a in 1 0; do PSIF_TC_TPT_ALIGN=$a timeout 300 python bench.py --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200 | sed "s/^/[align4=$a] /"; done
