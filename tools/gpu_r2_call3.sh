#!/bin/bash
# round 2, GPU call 3: packed-operand GEMM (producer split): tests, A/B GEMM timings, Be bench.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=12 -s > gpurun_out/r02c_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/r02c_pytest_gpu.log
for shape in "16384 14 256 768" "16384 14 256 256" "16384 14 256 1024" "16384 14 1024 256" "4682 44 256 768" "4682 44 256 1024"; do
  for pk in 0 1; do
    echo "== shape $shape packed $pk"; GEMM_PACKED=$pk timeout 120 python tools/gemm_bench.py $shape 20
  done
done > gpurun_out/r02c_gemm_bench.txt 2>&1
cat gpurun_out/r02c_gemm_bench.txt
timeout 300 python bench.py --steps 20 --warmup 5 --systems none --no-cpu-baseline > gpurun_out/r02c_bench_be.json 2> gpurun_out/r02c_bench_be.err
python - <<'PY'
import json
b=json.load(open('gpurun_out/r02c_bench_be.json'))
print("Be value", b["value"], "ms", b["ms_per_step"], "roofline", b["roofline"]["achieved"], b["roofline"]["frac"], "mh", b["mh_walker_steps_per_s"])
print(b["kernel_breakdown"])
PY
PSIF_PACK_PRODUCERS=0 timeout 300 python bench.py --steps 20 --warmup 5 --systems none --no-cpu-baseline --no-train-step > gpurun_out/r02c_bench_be_nopack.json 2> gpurun_out/r02c_bench_be_nopack.err
python - <<'PY'
import json
b=json.load(open('gpurun_out/r02c_bench_be_nopack.json'))
print("NOPACK Be value", b["value"], "ms", b["ms_per_step"], "roofline", b["roofline"]["achieved"], b["roofline"]["frac"])
print(b["kernel_breakdown"])
PY
du -sh gpurun_out
