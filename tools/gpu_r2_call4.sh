#!/bin/bash
# round 2, GPU call 4: warp-per-unit attention kernel for 5..14 electrons: stage + parity tests, Ne / N2 timings, ncu.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_stages_gpu.py tests/test_parity_gpu.py -m gpu -q --maxfail=12 -s > gpurun_out/r02d_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/r02d_pytest_gpu.log | cut -c1-300
timeout 600 python bench.py --steps 5 --warmup 3 --system Ne --systems N2 --no-cpu-baseline --no-train-step > gpurun_out/r02d_bench_ne.json 2> gpurun_out/r02d_bench_ne.err
tail -c 400 gpurun_out/r02d_bench_ne.err
python - <<'PY'
import json
b=json.load(open('gpurun_out/r02d_bench_ne.json'))
for n,s in b["systems"].items():
    print(n, {k:(round(v,1) if isinstance(v,float) else v) for k,v in s.items() if k!="kernel_ms"}, s.get("kernel_ms"))
PY
timeout 300 ncu --set full --clock-control none -k regex:attention_payload_warp -s 1 -c 2 -o gpurun_out/r02d_att_ne python tools/energy_only.py Ne 1 > gpurun_out/r02d_ncu_ne.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:attention_payload_warp -s 1 -c 1 -o gpurun_out/r02d_att_n2 python tools/energy_only.py N2 1 > gpurun_out/r02d_ncu_n2.log 2>&1
for f in r02d_att_ne r02d_att_n2; do
  python profiles/ncu_summary.py gpurun_out/$f.ncu-rep > gpurun_out/$f.summary.txt 2>&1
  ls -la gpurun_out/$f.ncu-rep; rm -f gpurun_out/$f.ncu-rep
done
head -40 gpurun_out/r02d_att_ne.summary.txt
du -sh gpurun_out
