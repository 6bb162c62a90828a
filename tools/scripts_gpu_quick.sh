timeout 600 python -m pytest tests/test_stages_gpu.py -m gpu -q -x -p no:cacheprovider -k "tcgen05 or gelu" 2>&1 | grep -E "passed|failed|^E  |Error" | cut -c1-220
python tools/trace_gemm2.py 256 1024 2 14 | grep -E "^CTA0 tiles|per-k"
for c in 14 32 8 44; do timeout 300 python tools/gemm_bench.py $((229376 / c)) $c 256 1024 10; done
timeout 300 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], {k:v['ms'] for k,v in d['kernel_breakdown'].items()})"
