#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench.log 2>&1
echo "bench exit $?" >> gpurun_out/bench.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/bench.log
