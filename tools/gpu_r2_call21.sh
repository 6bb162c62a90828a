#!/bin/bash
# round 2, GPU call 21 (8 GPUs): the bench under torchrun on 8 ranks: Be weak + Ne / N2 sharded (strong)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 20 --warmup 5 --systems Be,Ne,N2 --no-train-step > gpurun_out/r02aa_bench_n8.json 2> gpurun_out/r02aa_bench_n8.err
echo "bench rc=$?"; tail -c 300 gpurun_out/r02aa_bench_n8.err
python - <<'PY'
import json
b=json.loads([l for l in open("gpurun_out/r02aa_bench_n8.json") if l.startswith("{")][-1])
print({k:b[k] for k in ("value","ms_per_step","n_gpus","scaling")}, b["e2e"]["value"], b["roofline"]["frac"], b["clocks"])
for n,s in b["systems"].items():
    print(n, {k:(round(v,1) if isinstance(v,float) else v) for k,v in s.items() if k!="kernel_ms"}, s.get("kernel_ms"))
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-300
