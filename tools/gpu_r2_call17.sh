#!/bin/bash
# round 2, GPU call 17 (2 GPUs): quick tests, then the bench under torchrun on 2 ranks (weak Be line + sharded Ne / N2)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_stages_gpu.py tests/test_parity_gpu.py -m gpu -q -x 2>&1 | tail -3
if [ "${PIPESTATUS[0]}" != "0" ]; then echo "tests failed or hung: stop"; exit 1; fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --systems Be,Ne,N2 --no-train-step > gpurun_out/r02q_bench_n2.json 2> gpurun_out/r02q_bench_n2.err
echo "bench rc=$?"; tail -c 400 gpurun_out/r02q_bench_n2.err
python - <<'PY'
import json
b=json.loads([l for l in open("gpurun_out/r02q_bench_n2.json") if l.startswith("{")][-1])
print({k:b[k] for k in ("value","ms_per_step","n_gpus","scaling")}, b["e2e"]["value"], b["roofline"]["frac"], b["clocks"], b.get("cpu_baseline"))
for n,s in b["systems"].items():
    print(n, {k:(round(v,1) if isinstance(v,float) else v) for k,v in s.items() if k!="kernel_ms"}, s.get("kernel_ms"))
PY
