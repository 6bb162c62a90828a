"""Where cluster 0 of the packed-operand GEMM (tc_gemm_ss_kernel) waits (tools only): cycles its MMA issuer, TMA producer
and one epilogue warp spend blocked, from the kernel's own clock64 counters.
Needs a library built with the counters: python -m psiformer_torch_b200.build --stats (rebuild without it afterwards).
usage: python tools/ss_stats.py [tokens C K N act]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from psiformer_torch_b200 import _lib as L  # noqa: E402

tokens, C, K, N, act = (int(a) for a in (sys.argv[1:6] + ["16384", "14", "256", "768", "0"][len(sys.argv) - 1:]))
rows = tokens * C
lib = L.load()
X = torch.randn(rows, K, device="cuda")
W = torch.randn(N, K, device="cuda") / K ** 0.5
b = torch.randn(N, device="cuda")
Y = torch.zeros(rows, N, device="cuda")
scr = torch.empty(3 * N * K + 4, device="cuda")
st = torch.cuda.current_stream().cuda_stream
tr = torch.zeros(2 * 18 * 512, dtype=torch.int64, device="cuda")
res = Y.data_ptr() if os.environ.get("GEMM_RES", "0") == "1" else None
for it in range(3):
    tr.zero_()
    L.check(lib.psif_stage_linear_tc(X.data_ptr(), W.data_ptr(), b.data_ptr(), res, rows, C, K, N, act, 0, 1, Y.data_ptr(), scr.data_ptr(),
                                     tr.data_ptr(), st))
torch.cuda.synchronize()
s = tr[:16].cpu().tolist()
kb = max(s[4], 1)
print(f"rows {rows} K {K} N {N} act {act}: MMA loop {s[0]} cycles = {s[0] / kb:.0f} per K block (12 MMAs = 852); blocked on FULL {100 * s[1] / max(s[0], 1):.1f}% "
      f"({s[3]} of {s[4]} K blocks had to block), on ACC_EMPTY {100 * s[2] / max(s[0], 1):.1f}%; producer blocked on EMPTY {100 * s[5] / max(s[0], 1):.1f}%; "
      f"epilogue warp: {s[9]} tiles, per tile: waiting for ACC_FULL {s[6] / max(s[9], 1):.0f}, staging-tile waits {s[7] / max(s[9], 1):.0f}, busy {s[8] / max(s[9], 1):.0f} cycles "
      f"(TMEM->registers {s[10] / max(s[9], 1):.0f}, store rounds: wait+bias+staging {s[12] / max(s[9], 1):.0f}, proxy fences {s[11] / max(s[9], 1):.0f}; payload GELU per tile: staging {s[13] / max(s[9], 1):.0f}, token table {s[14] / max(s[9], 1):.0f}, row items {s[15] / max(s[9], 1):.0f})")
