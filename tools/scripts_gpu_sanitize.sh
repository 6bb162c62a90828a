#!/bin/bash
# compute-sanitizer memcheck over the smoke test and the tensor-core GEMM stage tests
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python __graft_entry__.py smoke > gpurun_out/sanitize_smoke.log 2>&1
echo "exit $?" >> gpurun_out/sanitize_smoke.log
grep -E "ERROR SUMMARY|smoke ok|Invalid|exit" gpurun_out/sanitize_smoke.log | head
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_stages_gpu.py -m gpu -x -q -k "linear_tcgen05" -p no:cacheprovider > gpurun_out/sanitize_gemm.log 2>&1
echo "exit $?" >> gpurun_out/sanitize_gemm.log
grep -E "ERROR SUMMARY|passed|failed|Invalid|exit" gpurun_out/sanitize_gemm.log | head
