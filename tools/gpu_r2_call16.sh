#!/bin/bash
# round 2, GPU call 16: two-warps-per-unit attention kernel: tests (short timeouts), Ne / N2 lines, ncu
mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_stages_gpu.py -m gpu -q -x -k "attention or producers" 2>&1 | tail -4
if [ "${PIPESTATUS[0]}" != "0" ]; then echo "attention stage tests failed or hung: stop"; exit 1; fi
timeout 400 python -m pytest tests/test_parity_gpu.py -m gpu -q --maxfail=3 2>&1 | tail -3
for pair in 0 1; do
PSIF_ATT_PAIR=$pair timeout 400 python bench.py --steps 5 --warmup 3 --system Ne --systems N2 --no-cpu-baseline --no-train-step > gpurun_out/r02p_bench_ne_pair$pair.json 2> gpurun_out/r02p_bench_ne.err
tail -c 200 gpurun_out/r02p_bench_ne.err
python - <<PY
import json
b=json.load(open("gpurun_out/r02p_bench_ne_pair$pair.json"))
for n,s in b["systems"].items():
    print("pair=$pair", n, {k:(round(v,1) if isinstance(v,float) else v) for k,v in s.items() if k in ("evals_per_s","ms_per_step")}, s.get("kernel_ms"))
PY
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_payload_pair -s 1 -c 1 -f -o gpurun_out/r02p_att_ne python tools/energy_only.py Ne 1 > gpurun_out/r02p_ncu_ne.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_payload_pair -s 1 -c 1 -f -o gpurun_out/r02p_att_n2 python tools/energy_only.py N2 1 > gpurun_out/r02p_ncu_n2.log 2>&1
for f in r02p_att_ne r02p_att_n2; do
python profiles/ncu_summary.py gpurun_out/$f.ncu-rep > gpurun_out/$f.summary.txt 2>&1
python tools/ncu_hot_lines.py gpurun_out/$f.ncu-rep 30 > gpurun_out/$f.hot.txt 2>&1
grep -E "gpu__time|issue_active|inst_executed.sum|pipe_fma|dram__bytes|warps_active|registers" gpurun_out/$f.summary.txt
done
head -24 gpurun_out/r02p_att_ne.hot.txt
