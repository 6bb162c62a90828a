#!/bin/bash
# weak-scaling bench on N GPUs of one box (the driver's own launch line)
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.log 2>&1
echo "exit $?" >> gpurun_out/bench_n$N.log
tail -3 gpurun_out/bench_n$N.log | cut -c1-1200
