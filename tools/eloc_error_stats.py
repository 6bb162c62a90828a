"""Distribution of |E_L(GPU) - E_L(fp64 oracle)| over a sample of walkers (tools only; A/B of kernel variants)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from oracle import psiformer_oracle as O
from gpu_util import make_engine

name, n = sys.argv[1], int(sys.argv[2])
sysm = O.SYSTEMS[name]
params = O.synthetic_params(sysm, 1234)
eng = make_engine(sysm, params)
x = O.synthetic_walkers(sysm, n, 99).cuda()
out = eng.local_energy(x, want_grad=True)
ref = O.local_energy_parts(sysm, O.cast_params(params, torch.float64), x.cpu().double())
ref32 = O.local_energy_parts(sysm, params, x.cpu())
ok = out["status"].cpu() == 0
e = (out["e_loc"].double().cpu() - ref["e_loc"]).abs()[ok]
e32 = (ref32["e_loc"].double() - ref["e_loc"]).abs()[ok]
q = torch.tensor([0.5, 0.9, 0.99], dtype=torch.float64)
print(f"{name} n={int(ok.sum())}: ours  med/p90/p99/max {e.quantile(q).tolist()} {e.max():.3e}  >1e-4: {(e > 1e-4).sum().item()}")
print(f"{name} n={int(ok.sum())}: ref32 med/p90/p99/max {e32.quantile(q).tolist()} {e32.max():.3e}  >1e-4: {(e32 > 1e-4).sum().item()}")
print("ratio of medians ours/ref32:", (e.median() / e32.median()).item())
