"""Clamp-active walkers in the full pipeline (tools only): E_L of flagged N2 walkers against the fp64 oracle, with and
without the fix-up kernel.  usage: python tools/clamp_walkers.py [N2] [walkers]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from oracle import psiformer_oracle as O  # noqa: E402
from gpu_util import make_engine  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "N2"
W = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
sysm = O.SYSTEMS[name]
params = O.synthetic_params(sysm, 1234)
x = O.synthetic_walkers(sysm, W, 99).cuda()
res = {}
for fix in ("1", "0"):
    os.environ["PSIF_CLAMP_FIXUP"] = fix
    eng = make_engine(sysm, params)
    out = eng.local_energy(x, want_grad=True)
    res[fix] = {k: v.cpu() for k, v in out.items()}
st = res["1"]["status"]
flag = ((st & 2) != 0).nonzero().flatten()
print(f"{name}: {W} walkers, {flag.numel()} clamp-active ({100.0 * flag.numel() / W:.2f} %), status==0: {(st == 0).float().mean():.4f}")
idx = flag[:6]
if idx.numel():
    ref = O.local_energy_parts(sysm, O.cast_params(params, torch.float64), x[idx].cpu().double())
    for j, i in enumerate(idx.tolist()):
        print(f"walker {i}: oracle64 E_L {ref['e_loc'][j].item():+.6e}  fix-up {res['1']['e_loc'][i].item():+.6e}  "
              f"smooth {res['0']['e_loc'][i].item():+.6e}   log|psi| oracle {ref['logabs'][j].item():+.6f} ours {res['1']['logabs'][i].item():+.6f}")
