#!/bin/bash
timeout 100 python tools/trace_gemm2.py 256 1024 2 14 2>&1 | tail -18 | cut -c1-300
