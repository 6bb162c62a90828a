#!/bin/bash
mkdir -p gpurun_out
for cl in 1 2; do
PSIF_TC_CLUSTER=$cl timeout 300 python -m pytest tests/test_stages_gpu.py -m gpu -q -s -k "tcgen05" -p no:cacheprovider > gpurun_out/pytest_cl$cl.log 2>&1
echo "cl $cl pytest exit $?" >> gpurun_out/pytest_cl$cl.log
grep -E "^\[tcgen05|passed|failed|rror" gpurun_out/pytest_cl$cl.log | head -8
PSIF_TC_CLUSTER=$cl timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_cl$cl.log 2>&1
python - <<PY
import json
l=[x for x in open("gpurun_out/bench_cl$cl.log") if x.startswith("{")]
if l:
    d=json.loads(l[-1]); print("cl$cl", d["value"], d["ms_per_step"], d["roofline"]["achieved"], d["mh_walker_steps_per_s"], {k:v["ms"] for k,v in d["kernel_breakdown"].items()})
else:
    print(open("gpurun_out/bench_cl$cl.log").read()[-1500:])
PY
done
PSIF_TC_VARIANT=ss timeout 600 python bench.py --no-cpu-baseline --steps 5 > gpurun_out/bench_ss.log 2>&1; grep -o '"gemm": {"ms": [0-9.]*' gpurun_out/bench_ss.log
timeout 900 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "^FAILED|passed|failed" gpurun_out/pytest_gpu.log
