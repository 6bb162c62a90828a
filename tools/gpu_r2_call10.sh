#!/bin/bash
# round 2, GPU call 10: ncu (full set + source) of the warp-per-unit attention kernel on Ne and N2
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_payload_warp -s 1 -c 1 -f -o gpurun_out/r02j_att_ne python tools/energy_only.py Ne 1 > gpurun_out/r02j_ncu_ne.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_payload_warp -s 1 -c 1 -f -o gpurun_out/r02j_att_n2 python tools/energy_only.py N2 1 > gpurun_out/r02j_ncu_n2.log 2>&1
for f in r02j_att_ne r02j_att_n2; do
  python profiles/ncu_summary.py gpurun_out/$f.ncu-rep > gpurun_out/$f.summary.txt 2>&1
  python tools/ncu_hot_lines.py gpurun_out/$f.ncu-rep 45 > gpurun_out/$f.hot.txt 2>&1
done
head -50 gpurun_out/r02j_att_ne.hot.txt
