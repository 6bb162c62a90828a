"""A/B timing of the tensor-core Linear variants on one box (CUDA events, inputs larger than L2).
usage: python tools/gemm_bench.py [tokens C K N reps]"""
import sys
import torch
sys.path.insert(0, ".")
from psiformer_torch_b200 import _lib

tokens, C, K, N, reps = (int(a) for a in (sys.argv[1:6] + ["16384", "14", "256", "1024", "20"][len(sys.argv) - 1:]))
rows = tokens * C
L = _lib.load()
X = torch.randn(rows, K, device="cuda")
W = torch.randn(N, K, device="cuda") / K ** 0.5
b = torch.randn(N, device="cuda")
Y = torch.empty(rows, N, device="cuda")
scr = torch.empty(3 * N * K + 4, device="cuda")
PACKED = int(__import__("os").environ.get("GEMM_PACKED", "0"))   # 1: A operand as the packed fp16 pair (contents irrelevant for timing)
MODE = int(__import__("os").environ.get("GEMM_MODE", "0"))   # 0 fp16 split, 1 tf32 split
st = torch.cuda.current_stream().cuda_stream


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return a.elapsed_time(e) / reps * 1e3


def lin(act):
    _lib.check(L.psif_stage_linear_tc(X.data_ptr(), W.data_ptr(), b.data_ptr(), None, rows, C, K, N, act, MODE, PACKED, Y.data_ptr(), scr.data_ptr(), None, st))


def gelu():
    _lib.check(L.psif_stage_gelu(Y.data_ptr(), tokens, C, N, Y.data_ptr(), st))


split = timed(lambda: _lib.check(L.psif_stage_linear_tc(X.data_ptr(), W.data_ptr(), b.data_ptr(), None, 512, C, K, N, 0, MODE, PACKED, Y.data_ptr(), scr.data_ptr(), None, st)))
t0, tg = timed(lambda: lin(0)), timed(gelu)
try:
    t2 = timed(lambda: lin(2))
except Exception:
    t2 = float("nan")
fl = 2.0 * rows * K * N
print(f"rows {rows} K {K} N {N} C {C}: plain {t0:.1f} us ({fl / t0 / 1e6:.1f} TF/s)  gelu kernel {tg:.1f} us  fused {t2:.1f} us "
      f"(saves {t0 + tg - t2:.1f} us; tiny-launch overhead {split:.1f} us)")
