#!/bin/bash
# ncu --set full with source correlation of the CTA-per-(walker, head) attention kernel on the Ne workload
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_payload_v2 -s 5 -c 1 -f -o gpurun_out/att_v2_ne python tools/energy_only.py Ne 3 > gpurun_out/ncu_att_ne.log 2>&1
ls -la gpurun_out
