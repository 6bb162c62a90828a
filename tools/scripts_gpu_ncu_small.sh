#!/bin/bash
# ncu --set full of the small kernels around the network (MH propose / accept, determinant combine, Jastrow +
# potential, embedding, envelope) on the Be and N2 workloads: their DRAM bytes and time give the achieved HBM GB/s
mkdir -p gpurun_out
for sysname in Be N2; do
  timeout 240 ncu --set full --clock-control none --import-source on \
    -k regex:'mh_propose|mh_accept|det_combine|jastrow_potential|embed|envelope|orbital' -c 24 -f \
    -o gpurun_out/small_${sysname} python tools/mh_only.py ${sysname} 2 > gpurun_out/ncu_small_${sysname}.log 2>&1
  python profiles/ncu_summary.py gpurun_out/small_${sysname}.ncu-rep > gpurun_out/ncu_small_${sysname}_summary.txt 2>&1
done
ls -la gpurun_out
