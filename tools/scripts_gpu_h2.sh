#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_h.log 2>&1
tail -1 gpurun_out/bench_h.log | cut -c1-2500
for s in Be LiH Ne; do timeout 300 python tools/eloc_error_stats.py $s 256 2>&1 | tail -3; done
