#!/bin/bash
# fp16-split GEMM: correctness of the stage tests, A/B timing against the tf32-split kernel, then the full suite + bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_stages_gpu.py -m gpu -x -q -s -k "linear_tcgen05" -p no:cacheprovider > gpurun_out/pytest_h.log 2>&1
echo "exit $?" >> gpurun_out/pytest_h.log
grep -E "^\[tcgen05|passed|failed|Error|error|exit" gpurun_out/pytest_h.log | cut -c1-200 | head -40
for v in h 2cta; do
  for shape in "16384 14 256 768" "16384 14 256 256" "16384 14 256 1024" "16384 14 1024 256"; do
    PSIF_TC_VARIANT=$v timeout 120 python tools/gemm_bench.py $shape 20 2>&1 | tail -1 | sed "s/^/[$v] /"
  done
done
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_h.log 2>&1
tail -1 gpurun_out/bench_h.log | cut -c1-2500
for s in Be LiH; do timeout 300 python tools/eloc_error_stats.py $s 256 2>&1 | tail -4; done
