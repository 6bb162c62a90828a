#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -p no:cacheprovider -k "attention or local_energy or logpsi" 2>&1 | tail -3
timeout 600 python tools/bench_configs.py He Ne N2 2>&1 | cut -c1-420
