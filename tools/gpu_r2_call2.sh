#!/bin/bash
# round 2, GPU call 2: pipe-rate microbenchmark (FFMA / legacy HMMA), re-run of the tests that failed in call 1,
# ncu --set full of the small kernels of the Be value path (Metropolis step) and of the energy-mode tail kernels.
mkdir -p gpurun_out
./tools/_pipe_rates > gpurun_out/r02b_pipe_rates.txt 2>&1; cat gpurun_out/r02b_pipe_rates.txt
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_logdet_op_gpu.py tests/test_dropin_gpu.py tests/test_train_rules_gpu.py -m gpu -q --maxfail=12 -s > gpurun_out/r02b_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r02b_pytest_gpu.log
timeout 300 ncu --set full --clock-control none --import-source on \
  -k regex:'mh_propose|mh_accept|det_combine|jastrow_potential|embed_kernel|orbital_envelope|layernorm|attention_payload_n4' -s 20 -c 14 \
  -o gpurun_out/r02b_small_be python tools/mh_only.py Be 3 > gpurun_out/r02b_ncu_small_be.log 2>&1
echo "ncu be rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on \
  -k regex:'det_combine_kernel<7, true>|det_combine_kernel<7, 1>|jastrow_potential|embed_kernel|orbital_envelope' -s 12 -c 4 \
  -o gpurun_out/r02b_small_n2 python tools/mh_only.py N2 1 > gpurun_out/r02b_ncu_small_n2.log 2>&1
echo "ncu n2 rc=$?"
ls -la gpurun_out | head -30
