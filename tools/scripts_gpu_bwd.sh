#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_backward_gpu.py -m gpu -q -s -p no:cacheprovider > gpurun_out/pytest_bwd.log 2>&1
echo "exit $?" >> gpurun_out/pytest_bwd.log
grep -E "^\[|^FAILED|passed|failed|^E  " gpurun_out/pytest_bwd.log | cut -c1-250 | head -40
