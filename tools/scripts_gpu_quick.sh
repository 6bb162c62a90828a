timeout 600 python -m pytest tests/test_stages_gpu.py -m gpu -q -x -p no:cacheprovider -k "tcgen05 or gelu" -s 2>&1 | grep -E "passed|failed|^E  |Error|rel err" | cut -c1-220
for kp in 512 0 256; do echo "kpass $kp"; PSIF_TC_KPASS=$kp timeout 300 python tools/gemm_bench.py 16384 14 1024 256 10; done 2>&1 | cut -c1-100
timeout 300 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], {k:v['ms'] for k,v in d['kernel_breakdown'].items()}, d['clocks'])"
