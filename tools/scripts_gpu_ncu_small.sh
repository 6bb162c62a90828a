#!/bin/bash
# ncu of the small kernels around the network (MH propose / accept, determinant combine, Jastrow + potential,
# embedding, envelope): DRAM bytes and time per launch give the achieved HBM GB/s.
#
# NOT YET RUN TO COMPLETION.  The first version of this script (--set full --import-source on, 24 launches, Be and N2
# in one call) was killed at a 400 s gpurun limit in round 1 and its reports were larger than the 64 MiB that
# gpurun_out/ brings back, so no numbers exist from it.  --set full replays every kernel ~40 times and saves/restores
# the multi-GB energy workspace around each replay.  This version asks for three metrics (one replay pass), one
# system per call, and keeps only the CSV.
# usage: bash tools/scripts_gpu_ncu_small.sh [Be]
sysname=${1:-Be}
mkdir -p gpurun_out
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  -k regex:'mh_propose|mh_accept|det_combine|jastrow_potential|embed|envelope' -c 24 --csv \
  --log-file gpurun_out/ncu_small_${sysname}.csv python tools/mh_only.py ${sysname} 2 > gpurun_out/ncu_small_${sysname}.log 2>&1
tail -3 gpurun_out/ncu_small_${sysname}.log
ls -la gpurun_out
