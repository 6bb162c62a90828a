#!/bin/bash
# round 2, GPU call 24: compute-sanitizer memcheck over the kernels added since the r02t pass (clamp fix-up, Jastrow lanes,
# value-mode embed / envelope, Metropolis accept) and racecheck over the clamp fix-up's shared-memory phases
mkdir -p gpurun_out
timeout 700 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests -m gpu -x -q -p no:cacheprovider -k "clamp or jastrow or potential or metropolis or embed or envelope or determinant" > gpurun_out/r02ai_sanitize_memcheck.log 2>&1
echo "exit $?" >> gpurun_out/r02ai_sanitize_memcheck.log
grep -E "ERROR SUMMARY|passed|failed|Invalid|exit" gpurun_out/r02ai_sanitize_memcheck.log | head
timeout 500 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests -m gpu -x -q -p no:cacheprovider -k "clamp_active" > gpurun_out/r02ai_sanitize_racecheck.log 2>&1
echo "exit $?" >> gpurun_out/r02ai_sanitize_racecheck.log
grep -E "RACECHECK SUMMARY|passed|failed|hazard|exit" gpurun_out/r02ai_sanitize_racecheck.log | head
