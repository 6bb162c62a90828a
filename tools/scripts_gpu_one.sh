#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -s -k "full_size" -p no:cacheprovider > gpurun_out/pytest_one.log 2>&1
grep -E "^\[|^FAILED|passed|failed|^E  " gpurun_out/pytest_one.log | cut -c1-300
