"""A few Metropolis-Hastings steps (no CUDA graph) and one energy pass of one BASELINE.json system: the target for ncu
captures of the small HBM/latency-bound kernels (mh_propose / mh_accept / det_combine / jastrow_potential / embed).
usage: python tools/mh_only.py [Be] [steps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from psiformer_torch_b200.config import BENCH_SYSTEMS  # noqa: E402
from psiformer_torch_b200.psiformer import PsiFormer  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "Be"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
mcfg, W, step_size = BENCH_SYSTEMS[name]
dev = torch.device("cuda", 0)
torch.manual_seed(1234)
model = PsiFormer(mcfg).to(dev)
x = torch.randn(W, mcfg.n_electron_num, 3, device=dev)
eng = model.ready_engine(dev)
logabs = torch.empty(W, device=dev)
sign = torch.empty(W, device=dev)
n_acc = torch.zeros(1, dtype=torch.int64, device=dev)
eng.mh_steps(x, logabs, sign, steps, step_size, have_logabs=False, seed=7, n_accept=n_acc)
out = eng.local_energy(x, guard=False)
torch.cuda.synchronize()
print(name, W, "accepted", int(n_acc.item()), "of", steps * W, "E", float(out["e_loc"][torch.isfinite(out["e_loc"])].mean()))
