#!/bin/bash
# round 2, GPU call 27: the first-layer attention kernel alone: timing at N = 10 / 14, ncu --set full with source lines
mkdir -p gpurun_out
python tools/afl_only.py 14 1702 3
python tools/afl_only.py 10 3276 3
python tools/afl_only.py 4 4096 3
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_first_layer -s 1 -c 1 -f -o gpurun_out/afl_n14 python tools/afl_only.py 14 1702 2 > gpurun_out/r02an_ncu_afl.log 2>&1
python profiles/ncu_summary.py gpurun_out/afl_n14.ncu-rep > gpurun_out/r02an_afl_n14_summary.txt 2>&1
python tools/ncu_hot_lines.py gpurun_out/afl_n14.ncu-rep 45 > gpurun_out/r02an_afl_n14_hot.txt 2>&1
grep -E "gpu__time|dram__bytes|issue|warp_cycles_per|achieved_occupancy|l1tex__data_bank_conflicts|smsp__inst_executed.sum " gpurun_out/r02an_afl_n14_summary.txt | head -20
head -60 gpurun_out/r02an_afl_n14_hot.txt
