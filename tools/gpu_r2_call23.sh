#!/bin/bash
# round 2, GPU call 23: clamp fix-up: how many bench walkers are flagged, and walkers-per-CTA sweep on N2
mkdir -p gpurun_out
timeout 300 python tools/clamp_bench_count.py N2 Ne Be 2>&1 | tail -4
for wpc in 1 2 4 8; do
  PSIF_CLAMP_WPC=$wpc timeout 300 python bench.py --system N2 --systems none --no-cpu-baseline --no-train-step --steps 4 --warmup 3 > gpurun_out/r02ag_n2_wpc$wpc.json 2> gpurun_out/r02ag_n2_wpc$wpc.err; echo "wpc=$wpc rc=$?"
  python -c "
import json
b=json.load(open('gpurun_out/r02ag_n2_wpc$wpc.json'))
print('wpc=$wpc', b['value'], b['ms_per_step'])"
done
PSIF_CLAMP_FIXUP=0 timeout 300 python bench.py --system N2 --systems none --no-cpu-baseline --no-train-step --steps 4 --warmup 3 > gpurun_out/r02ag_n2_off.json 2> gpurun_out/r02ag_n2_off.err
python -c "
import json
b=json.load(open('gpurun_out/r02ag_n2_off.json'))
print('off', b['value'], b['ms_per_step'])"
