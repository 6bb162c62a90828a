"""Where the time of an energy step goes besides the kernels (tools only): the public API (status guard = one sync per
call), the asynchronous engine call, and a whole-pass CUDA graph, interleaved on one box; plus the in-library per-class
kernel times.  usage: python tools/step_gaps.py [Be] [reps]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from psiformer_torch_b200 import _lib  # noqa: E402
from psiformer_torch_b200.config import BENCH_SYSTEMS  # noqa: E402
from psiformer_torch_b200.hamiltonian import Hamiltonian  # noqa: E402
from psiformer_torch_b200.psiformer import PsiFormer  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "Be"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
mcfg, W, _ = BENCH_SYSTEMS[name]
dev = torch.device("cuda", 0)
torch.manual_seed(1234)
model = PsiFormer(mcfg).to(dev)
N = mcfg.n_electron_num
x = torch.randn(W, N, 3, device=dev)
eng = model.ready_engine(dev)
ham = Hamiltonian(model, n_elec=N, Z=mcfg.nuclear_charge)
accum = torch.zeros(3, dtype=torch.float64, device=dev)


def timed(fn, n):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def api():
    accum.zero_()
    return ham.local_energy(x, accum=accum)


def raw():
    return eng.local_energy(x, accum=accum, guard=False)


big = eng.graph_max_rows
def graphed():
    eng.graph_max_rows = 1 << 40
    try:
        return eng.local_energy(x, accum=accum, guard=False)
    finally:
        eng.graph_max_rows = big


for f in (api, raw, graphed):
    for _ in range(3):
        f()
for _ in range(60):          # bring the part to its sustained clocks
    raw()
res = {"api_guard": [], "engine_async": [], "graph_async": []}
for _ in range(3):
    res["api_guard"].append(timed(api, reps))
    res["engine_async"].append(timed(raw, reps))
    res["graph_async"].append(timed(graphed, reps))
_lib.profile_enable(eng._handle, True)
raw()
prof = _lib.profile_read(eng._handle)
_lib.profile_enable(eng._handle, False)
print(json.dumps({"system": name, "walkers": W, "ms_per_step": {k: [round(t, 4) for t in v] for k, v in res.items()},
                  "kernel_sum_ms": round(sum(v["ms"] for v in prof.values()), 4),
                  "breakdown_ms": {k: round(v["ms"], 4) for k, v in prof.items() if v["groups"] > 0}}))
