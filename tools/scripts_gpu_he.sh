#!/bin/bash
timeout 600 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -3
timeout 300 python tools/bench_configs.py He 2>&1 | cut -c1-420
PSIF_ATT_N4=0 timeout 300 python tools/bench_configs.py He 2>&1 | cut -c1-420
