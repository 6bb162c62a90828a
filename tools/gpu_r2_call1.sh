#!/bin/bash
# round 2, GPU call 1: GPU tests, default bench line (all five systems), reference arm, smoke, ncu of the small
# HBM/latency-bound kernels on Ne and N2 (metrics only: one replay pass).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/r02a_gpu.txt 2>&1
(python -c "import os; print(os.cpu_count())"; lscpu | head -20) > gpurun_out/r02a_cpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --maxfail=12 -s > gpurun_out/r02a_pytest_gpu.log 2>&1
echo "pytest rc=$?"
tail -5 gpurun_out/r02a_pytest_gpu.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/r02a_smoke.log 2>&1; tail -1 gpurun_out/r02a_smoke.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02a_bench_n1.json 2> gpurun_out/r02a_bench_n1.err
echo "bench rc=$?"; tail -c 600 gpurun_out/r02a_bench_n1.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02a_bench_ref.json 2> gpurun_out/r02a_bench_ref.err
for s in Ne N2; do
  timeout 240 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k regex:'mh_propose|mh_accept|det_combine|jastrow_potential|embed_kernel|orbital_envelope' -c 16 --csv \
    --log-file gpurun_out/r02a_ncu_small_${s}.csv python tools/mh_only.py ${s} 2 > gpurun_out/r02a_ncu_small_${s}.log 2>&1
  echo "ncu ${s} rc=$?"
done
ls -la gpurun_out
