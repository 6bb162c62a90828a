#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_stages_gpu.py -m gpu -q -s -k "tcgen05" -p no:cacheprovider > gpurun_out/pytest_ts.log 2>&1
echo "ts pytest exit $?" >> gpurun_out/pytest_ts.log
grep -E "^\[tcgen05|passed|failed|Error|error" gpurun_out/pytest_ts.log | head
timeout 900 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_ts.log 2>&1
PSIF_TC_VARIANT=ss timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_ss.log 2>&1
grep -E "^\[.*\] (sample )?\|E_L|^FAILED|passed|failed" gpurun_out/pytest_gpu.log
for v in ts ss; do python - <<PY
import json
l=[x for x in open("gpurun_out/bench_$v.log") if x.startswith("{")]
d=json.loads(l[-1]); print("$v", d["value"], d["ms_per_step"], d["roofline"]["achieved"], d["mh_walker_steps_per_s"], {k:v["ms"] for k,v in d["kernel_breakdown"].items()})
PY
done
