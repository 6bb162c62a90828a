#!/bin/bash
# round 2, GPU call 18: value path on the packed-operand kernel (plain GELU writes the pair), warp-per-token embed
mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_stages_gpu.py -m gpu -q -x -k "tcgen05 or packed or embed" 2>&1 | tail -4
if [ "${PIPESTATUS[0]}" != "0" ]; then echo "stage tests failed or hung: stop"; exit 1; fi
timeout 600 python -m pytest tests -m gpu -q --maxfail=5 2>&1 | tail -4
for pv in 0 1; do for s in Be Ne; do echo -n "pack_value=$pv "; PSIF_PACK_VALUE=$pv timeout 200 python tools/value_breakdown.py $s 2>&1 | tail -1; done; done | tee gpurun_out/r02r_value_breakdown.jsonl
