#!/bin/bash
# round 2, GPU call 11: attention warp kernel with elected-lane bulk copies + reordered output sweeps; value-path breakdown
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_stages_gpu.py tests/test_parity_gpu.py -m gpu -q --maxfail=10 -k "attention or local_energy or full_size or producers" 2>&1 | tail -4
timeout 600 python bench.py --steps 5 --warmup 3 --system Ne --systems N2 --no-cpu-baseline --no-train-step > gpurun_out/r02k_bench_ne.json 2> gpurun_out/r02k_bench_ne.err
tail -c 300 gpurun_out/r02k_bench_ne.err
python - <<'PY'
import json
b=json.load(open("gpurun_out/r02k_bench_ne.json"))
for n,s in b["systems"].items():
    print(n, {k:(round(v,1) if isinstance(v,float) else v) for k,v in s.items() if k!="kernel_ms"}, s.get("kernel_ms"))
PY
for s in Be Ne He; do timeout 200 python tools/value_breakdown.py $s 2>&1 | tail -1; done | tee gpurun_out/r02k_value_breakdown.jsonl
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_payload_warp -s 1 -c 1 -f -o gpurun_out/r02k_att_ne python tools/energy_only.py Ne 1 > gpurun_out/r02k_ncu_ne.log 2>&1
python profiles/ncu_summary.py gpurun_out/r02k_att_ne.ncu-rep > gpurun_out/r02k_att_ne.summary.txt 2>&1
python tools/ncu_hot_lines.py gpurun_out/r02k_att_ne.ncu-rep 30 > gpurun_out/r02k_att_ne.hot.txt 2>&1
grep -E "gpu__time|issue_active|inst_executed.sum|pipe_fma" gpurun_out/r02k_att_ne.summary.txt; head -20 gpurun_out/r02k_att_ne.hot.txt
