#!/bin/bash
mkdir -p gpurun_out
for pf in 0 4 8 12 24 48; do
PSIF_TC_CLUSTER=1 PSIF_TC_PREFETCH=$pf timeout 600 python bench.py --steps 5 --warmup 3 --profile-mode > gpurun_out/bench_pf$pf.log 2>&1
echo "pf $pf: $(grep profile_mode gpurun_out/bench_pf$pf.log)"
done
PSIF_TC_CLUSTER=2 PSIF_TC_PREFETCH=12 timeout 600 python bench.py --steps 5 --warmup 3 --profile-mode > gpurun_out/bench_pfc2.log 2>&1; echo "cl2 pf12: $(grep profile_mode gpurun_out/bench_pfc2.log)"
