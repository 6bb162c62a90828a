"""One launch of the first-layer attention kernel at workload size (tools only; for ncu).
usage: python tools/afl_only.py [N=14] [walkers=1702] [reps=3]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from psiformer_torch_b200 import _lib as L  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 14
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1702
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
d, H = 256, 4
C = 3 * N + 2
lib = L.load()
q5 = torch.randn(B, N, 5, 3 * d, device="cuda") * 0.7
out = torch.empty(B, N, C, d, device="cuda")
st = torch.cuda.current_stream().cuda_stream
ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
ev[0].record()
for r in range(reps):
    L.check(lib.psif_stage_attention_first_layer(q5.data_ptr(), B, N, d, H, 1, out.data_ptr(), st))
    ev[r + 1].record()
torch.cuda.synchronize()
ts = [ev[r].elapsed_time(ev[r + 1]) for r in range(reps)]
gb = (q5.numel() + out.numel()) * 4 / 1e9
print(f"N={N} walkers={B}: {min(ts) * 1e3:.1f} us, {gb / (min(ts) * 1e-3) / 1e3:.2f} TB/s ({gb:.2f} GB)")
