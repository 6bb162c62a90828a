"""Per-kernel-class device time of ONE log|psi| evaluation (the Metropolis step's forward) at the bench walker count
(tools only).  usage: python tools/value_breakdown.py [Be] [walkers]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from psiformer_torch_b200 import _lib  # noqa: E402
from psiformer_torch_b200.config import BENCH_SYSTEMS, Train_Config  # noqa: E402
from psiformer_torch_b200.mcmc import MH  # noqa: E402
from psiformer_torch_b200.psiformer import PsiFormer  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "Be"
mcfg, W, step = BENCH_SYSTEMS[name]
if len(sys.argv) > 2:
    W = int(sys.argv[2])
dev = torch.device("cuda", 0)
torch.manual_seed(1234)
model = PsiFormer(mcfg).to(dev)
N = mcfg.n_electron_num
mh = MH(model, Train_Config(batch_size=W, step_size=step, seed=1234), N, device=dev)
mh._run_steps(torch.randn(W, N, 3, device=dev), 32)
x = mh._state.clone()
eng = model.ready_engine(dev)
for _ in range(3):
    eng.logpsi(x, guard=False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n0 = _lib.launch_count()
e0.record()
for _ in range(20):
    eng.logpsi(x, guard=False)
e1.record()
torch.cuda.synchronize()
launches = (_lib.launch_count() - n0) / 20
t_fwd = e0.elapsed_time(e1) / 20
mh._run_steps(mh._state, 32)
torch.cuda.synchronize()
e0.record()
for _ in range(3):
    mh._run_steps(mh._state, 32)
e1.record()
torch.cuda.synchronize()
t_mh = e0.elapsed_time(e1) / 96
_lib.profile_enable(eng._handle, True)
eng.logpsi(x, guard=False)
prof = _lib.profile_read(eng._handle)
_lib.profile_enable(eng._handle, False)
print(json.dumps({"system": name, "walkers": W, "ms_per_logpsi_eager": round(t_fwd, 4), "launches_per_logpsi": launches,
                  "ms_per_mh_step_graph": round(t_mh, 4), "mh_walker_steps_per_s": round(W / t_mh * 1e3),
                  "breakdown_ms": {k: round(v["ms"], 4) for k, v in prof.items() if v["groups"] > 0},
                  "sum_ms": round(sum(v["ms"] for v in prof.values()), 4)}))
