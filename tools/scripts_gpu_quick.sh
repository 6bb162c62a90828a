#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_q.log 2>&1
grep -E "^\[tcgen05|^\[.*\] (sample )?\|E_L|^FAILED|passed|failed" gpurun_out/pytest_gpu.log
python - <<PY
import json
l=[x for x in open("gpurun_out/bench_q.log") if x.startswith("{")]
d=json.loads(l[-1]); print(d["value"], d["ms_per_step"], d["roofline"]["achieved"], d["mh_walker_steps_per_s"], {k:v["ms"] for k,v in d["kernel_breakdown"].items()})
PY
