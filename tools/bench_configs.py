"""Throughput of every BASELINE.json system on ONE GPU at its per-GPU walker count (tools; prints a table).

    python tools/bench_configs.py [He Be LiH Ne N2]
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from psiformer_torch_b200 import _lib  # noqa: E402
from psiformer_torch_b200.config import BENCH_SYSTEMS, Train_Config  # noqa: E402
from psiformer_torch_b200.mcmc import MH  # noqa: E402
from psiformer_torch_b200.psiformer import PsiFormer  # noqa: E402


def flops_fwd(N, d, L, K, nu, nd):
    return 8 * N * d + L * (24 * N * d * d + 4 * N * N * d) + 2 * d * K * (nu * nu + nd * nd) + (2 / 3) * K * (nu ** 3 + nd ** 3)


def main():
    names = sys.argv[1:] or list(BENCH_SYSTEMS)
    dev = torch.device("cuda", 0)
    rows = []
    for name in names:
        mcfg, W, step = BENCH_SYSTEMS[name]
        torch.manual_seed(1234)
        model = PsiFormer(mcfg).to(dev)
        N = mcfg.n_electron_num
        tcfg = Train_Config(batch_size=W, step_size=step, seed=1234)
        mh = MH(model, tcfg, N, device=dev)
        mh._run_steps(torch.randn(W, N, 3, device=dev), 32)
        x = mh._state.clone()
        eng = model.ready_engine(dev)
        for _ in range(3):
            out = eng.local_energy(x)
        torch.cuda.synchronize()
        reps = 10 if name in ("He", "Be", "LiH") else 4
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            out = eng.local_energy(x)
        e1.record()
        torch.cuda.synchronize()
        ms_e = e0.elapsed_time(e1) / reps
        mh._run_steps(mh._state, 32)      # warm-up (captures the CUDA graph for 32 steps)
        torch.cuda.synchronize()
        e0.record()
        mh._run_steps(mh._state, 32)
        e1.record()
        torch.cuda.synchronize()
        ms_mh = e0.elapsed_time(e1) / 32
        _lib.profile_enable(True)
        eng.local_energy(x)
        prof = _lib.profile_read()
        _lib.profile_enable(False)
        fwd = flops_fwd(N, mcfg.n_embd, mcfg.n_layer, mcfg.n_determinants, mcfg.n_spin_up, mcfg.n_spin_down)
        ok = float((out["status"] == 0).float().mean())
        row = {"system": name, "walkers": W, "N": N, "K": mcfg.n_determinants, "ms_per_energy_pass": round(ms_e, 3),
               "evals_per_s": round(W / ms_e * 1e3, 1), "algorithmic_tflops": round(W * (3 * N + 2) * fwd / ms_e / 1e9, 1),
               "mh_walker_steps_per_s": round(W / ms_mh * 1e3, 1), "status_ok_frac": round(ok, 4),
               "acceptance": round(mh.acceptance_rate, 3),
               "breakdown_ms": {k: round(v["ms"], 3) for k, v in prof.items() if v["groups"] > 0}}
        rows.append(row)
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
