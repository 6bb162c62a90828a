#!/bin/bash
mkdir -p gpurun_out
for bn in 256 128; do
PSIF_TC_BN=$bn timeout 300 python -m pytest tests/test_stages_gpu.py -m gpu -q -s -k "tcgen05" -p no:cacheprovider > gpurun_out/pytest_tc_$bn.log 2>&1
echo "tc pytest exit $?" >> gpurun_out/pytest_tc_$bn.log
PSIF_TC_BN=$bn timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_tc_$bn.log 2>&1
done
timeout 900 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -h "^\[tcgen05" gpurun_out/pytest_tc_*.log; tail -3 gpurun_out/pytest_tc_256.log;  tail -3 gpurun_out/pytest_tc_128.log; grep -E "^\[.*\] (sample )?\|E_L|^FAILED|passed|failed" gpurun_out/pytest_gpu.log
for bn in 256 128; do python - <<PY
import json
l=[x for x in open("gpurun_out/bench_tc_$bn.log") if x.startswith("{")]
d=json.loads(l[-1]); print($bn, d["value"], d["ms_per_step"], d["roofline"]["achieved"], {k:v["ms"] for k,v in d["kernel_breakdown"].items()})
PY
done
