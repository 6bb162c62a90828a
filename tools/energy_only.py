"""A few energy passes of one BASELINE.json system on random walkers, nothing else (target for ncu -k captures).
usage: python tools/energy_only.py [Ne] [passes]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from psiformer_torch_b200.config import BENCH_SYSTEMS  # noqa: E402
from psiformer_torch_b200.psiformer import PsiFormer  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "Ne"
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 3
mcfg, W, _ = BENCH_SYSTEMS[name]
dev = torch.device("cuda", 0)
torch.manual_seed(1234)
model = PsiFormer(mcfg).to(dev)
x = torch.randn(W, mcfg.n_electron_num, 3, device=dev)
eng = model.ready_engine(dev)
for _ in range(passes):
    out = eng.local_energy(x, guard=False)
torch.cuda.synchronize()
print(name, W, float(out["e_loc"][torch.isfinite(out["e_loc"])].mean()))
