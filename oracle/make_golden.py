"""Generate tests/golden/*.npz by running the UNMODIFIED reference on CPU.

TEST INFRASTRUCTURE.  Run in the build container only (it needs
/root/reference, which does not exist on the GPU box):

    WANDB_MODE=disabled python oracle/make_golden.py

For every single-nucleus case the reference's own classes produce the vectors
(PsiFormer, Hamiltonian, MH, logdet_matmul; fp32 as shipped and a .double()
copy as ground truth).  Weights come from
``psiformer_oracle.synthetic_params`` (numpy PCG64, reproducible anywhere) and
are loaded with ``load_state_dict(strict=True)`` so the 12.8 MB state_dict
need not be committed; its sha256 is stored instead.

Molecular cases (LiH, N2) cannot be run by the reference (psiformer.py:133);
their vectors come from the oracle's App. A.7 extension and are stored with
``pinned=0`` ("parity unpinned").
"""
from __future__ import annotations

import copy
import hashlib
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
os.environ.setdefault("WANDB_MODE", "disabled")

from oracle import psiformer_oracle as O  # noqa: E402

REF_SRC = "/root/reference/src"

# name -> (system, n_walkers, param_seed, walker_seed, step_size, mh_fixture_steps)
CASES = {
    # reference presets (config.py:77-152)
    "debug": (O.OracleSystem(1, 2, 4, 1, 2, 1, ((3.0, (0.0, 0.0, 0.0)),)), 16, 11, 12, 1.0, 6),
    "he_small": (O.SYSTEMS["He"], 64, 21, 22, 1.0, 6),
    "large": (O.OracleSystem(4, 32, 256, 4, 4, 2, ((6.0, (0.0, 0.0, 0.0)),)), 16, 31, 32, 0.8, 4),
    # BASELINE.json configs that the reference can run
    "be": (O.SYSTEMS["Be"], 32, 41, 42, 0.8, 4),
    "ne": (O.SYSTEMS["Ne"], 16, 51, 52, 0.8, 4),
    # 14 electrons / 32 determinants on ONE nucleus: pins the n_sigma=7 code path
    "z14": (O.OracleSystem(4, 4, 256, 32, 7, 7, ((14.0, (0.0, 0.0, 0.0)),)), 8, 61, 62, 0.8, 3),
}
MOLECULES = {
    "lih": (O.SYSTEMS["LiH"], 16, 71, 72, 0.8, 4),
    "n2": (O.SYSTEMS["N2"], 8, 81, 82, 0.8, 3),
}
BURN_STEPS = 24


def params_sha256(p) -> str:
    h = hashlib.sha256()
    for k in p:
        h.update(k.encode())
        h.update(p[k].numpy().tobytes())
    return h.hexdigest()


def system_arrays(sysm: O.OracleSystem):
    return dict(
        sys_shape=np.array([sysm.n_layer, sysm.n_head, sysm.n_embd, sysm.n_det, sysm.n_up, sysm.n_dn], dtype=np.int64),
        sys_Z=np.array([z for z, _ in sysm.nuclei], dtype=np.float64),
        sys_R=np.array([list(r) for _, r in sysm.nuclei], dtype=np.float64),
    )


def build_reference_model(sysm: O.OracleSystem, params):
    sys.path.insert(0, REF_SRC)
    from psiformer_torch.config import Model_Config
    from psiformer_torch.psiformer import PsiFormer

    cfg = Model_Config(
        n_layer=sysm.n_layer, n_head=sysm.n_head, n_embd=sysm.n_embd, n_features=3,
        n_determinants=sysm.n_det, n_electron_num=sysm.n_elec, n_spin_up=sysm.n_up,
        n_spin_down=sysm.n_dn, nuclear_charge=int(sysm.nuclei[0][0]),
    )
    model = PsiFormer(cfg)
    assert list(model.state_dict().keys()) == list(params.keys()), "state_dict order differs from App. A.6"
    model.load_state_dict(params, strict=True)
    return model.eval()


class SignTap:
    """Record the (log_abs, sign) pair PsiFormer drops at psiformer.py:191."""

    def __init__(self):
        import psiformer_torch.psiformer as mod
        self.mod = mod
        self.orig = mod.logdet_matmul
        self.last = None
        self.phis = None

    def __enter__(self):
        def tapped(x1, x2, w):
            out = self.orig(x1, x2, w)
            self.last = (out[0].detach().clone(), out[1].detach().clone())
            self.phis = (x1.detach().clone(), x2.detach().clone())
            return out
        self.mod.logdet_matmul = tapped
        return self

    def __exit__(self, *a):
        self.mod.logdet_matmul = self.orig


def run_reference(model, sysm, x, dtype):
    from psiformer_torch.hamiltonian import Hamiltonian, Potential

    m = copy.deepcopy(model).to(dtype)
    xx = x.to(dtype)
    with SignTap() as tap, torch.no_grad():
        logabs = m(xx)
        sign = tap.last[1].squeeze(-1)
        smin = O.min_singular_values(tap.phis[0].double(), tap.phis[1].double())
    ham = Hamiltonian(m, n_elec=sysm.n_elec, Z=int(sysm.nuclei[0][0]))
    g = ham.grad_log_psi(xx).detach()
    lap = ham.laplacian_log_psi(xx).detach()
    v = Potential(xx, int(sysm.nuclei[0][0])).potential()
    e = ham.local_energy(xx).detach()
    return dict(logabs=logabs, sign=sign, grad=g, lap=lap, pot=v, eloc=e, smin=smin)


def run_oracle(sysm, params, x, dtype):
    p = O.cast_params(params, dtype)
    xx = x.to(dtype)
    with torch.no_grad():
        logabs, sign = O.log_psi_and_sign(sysm, p, xx)
        feats, r_ae = O.electron_features(sysm, xx)
        phi = O.orbital_matrices(sysm, p, O.backbone(sysm, p, feats), r_ae)
        smin = O.min_singular_values(phi[0].double(), phi[1].double())
    parts = O.local_energy_parts(sysm, p, xx)
    return dict(logabs=logabs, sign=sign, grad=parts["grad"], lap=parts["lap"], pot=parts["pot"],
                eloc=parts["e_loc"], smin=smin)


def burn_in(target, x, step_size, seed):
    g = torch.Generator().manual_seed(seed)
    for _ in range(BURN_STEPS):
        eps = torch.randn(x.shape, generator=g)
        u = torch.rand(x.shape[0], generator=g)
        x, _ = O.mh_step(target, x, step_size, eps, u)
    return x


def reference_mh_fixture(model, sysm, x, step_size, steps, seed):
    """Drive the reference's MH._mh_step and replay its RNG (App. A.3)."""
    from psiformer_torch.config import Train_Config
    from psiformer_torch.mcmc import MH

    cfg = Train_Config(batch_size=x.shape[0], step_size=step_size)
    mh = MH(model, cfg, sysm.n_elec, device=torch.device("cpu"))
    eps_l, u_l, acc_l, lt_l, ls_l = [], [], [], [], []
    state = x.clone()
    with torch.inference_mode():
        for s in range(steps):
            torch.manual_seed(seed + s)
            new_state = mh._mh_step(state)
            torch.manual_seed(seed + s)
            eps = torch.randn_like(state)
            u = torch.rand(state.shape[0])
            trial = state + step_size * eps
            acc = (new_state != state).any(dim=(1, 2)) | (trial == state).all(dim=(1, 2))
            lt, ls = model(trial), model(state)
            chk = torch.log(u) < torch.min(2 * (lt - ls), torch.zeros_like(lt))
            assert torch.equal(chk, acc), "RNG replay does not reproduce the reference's decisions"
            eps_l.append(eps); u_l.append(u); acc_l.append(acc); lt_l.append(lt); ls_l.append(ls)
            state = new_state.clone()
    return dict(mh_eps=torch.stack(eps_l), mh_u=torch.stack(u_l), mh_accept=torch.stack(acc_l),
                mh_logpsi_trial=torch.stack(lt_l), mh_logpsi_state=torch.stack(ls_l), mh_final=state)


def oracle_mh_fixture(target, x, step_size, steps, seed):
    eps_l, u_l, acc_l, lt_l, ls_l = [], [], [], [], []
    state = x.clone()
    for s in range(steps):
        g = torch.Generator().manual_seed(seed + s)
        eps = torch.randn(state.shape, generator=g)
        u = torch.rand(state.shape[0], generator=g)
        with torch.no_grad():
            lt, ls = target(state + step_size * eps), target(state)
        new_state, acc = O.mh_step(target, state, step_size, eps, u)
        eps_l.append(eps); u_l.append(u); acc_l.append(acc); lt_l.append(lt); ls_l.append(ls)
        state = new_state
    return dict(mh_eps=torch.stack(eps_l), mh_u=torch.stack(u_l), mh_accept=torch.stack(acc_l),
                mh_logpsi_trial=torch.stack(lt_l), mh_logpsi_state=torch.stack(ls_l), mh_final=state)


def to_np(d, prefix=""):
    return {prefix + k: (v.numpy() if isinstance(v, torch.Tensor) else v) for k, v in d.items()}


def make_case(name, spec, pinned):
    sysm, nw, pseed, wseed, step, mh_steps = spec
    t0 = time.time()
    params = O.synthetic_params(sysm, pseed)
    x0 = O.synthetic_walkers(sysm, nw, wseed)
    with torch.no_grad():
        x = burn_in(lambda t: O.log_psi(sysm, params, t), x0, step, wseed + 1000)
    out = dict(pinned=np.int64(1 if pinned else 0), param_seed=np.int64(pseed), walker_seed=np.int64(wseed),
               step_size=np.float64(step), params_sha256=np.array(params_sha256(params)), x=x.numpy())
    out.update(system_arrays(sysm))
    if pinned:
        model = build_reference_model(sysm, params)
        r32 = run_reference(model, sysm, x, torch.float32)
        r64 = run_reference(model, sysm, x, torch.float64)
        mh = reference_mh_fixture(model, sysm, x, step, mh_steps, wseed + 2000)
        # the oracle must agree with the reference it restates (checked again in tests/)
        o64 = run_oracle(sysm, params, x, torch.float64)
        for k in ("logabs", "grad", "lap", "pot", "eloc"):
            err = (o64[k] - r64[k]).abs().max().item()
            scale = max(1.0, r64[k].abs().max().item())
            assert err <= 1e-9 * scale, (name, k, err)
        assert torch.equal(o64["sign"], r64["sign"])
    else:
        r32 = run_oracle(sysm, params, x, torch.float32)
        r64 = run_oracle(sysm, params, x, torch.float64)
        with torch.no_grad():
            mh = oracle_mh_fixture(lambda t: O.log_psi(sysm, params, t), x, step, mh_steps, wseed + 2000)
    out.update(to_np(r32, "ref32_"))
    out.update(to_np(r64, "ref64_"))
    out.update(to_np(mh))
    path = os.path.join(ROOT, "tests", "golden", f"{name}.npz")
    np.savez_compressed(path, **out)
    d = (r32["eloc"].double() - r64["eloc"]).abs()
    print(f"{name:9s} pinned={int(pinned)} B={nw} N={sysm.n_elec} K={sysm.n_det}  "
          f"|E32-E64| med {d.median():.2e} max {d.max():.2e}  smin min {r64['smin'].min():.2e}  "
          f"accept {mh['mh_accept'].float().mean():.2f}  {os.path.getsize(path)/1024:.0f} KiB  {time.time()-t0:.0f}s")


def make_kats():
    """Analytic known-answer vectors resurrected from the reference's stale
    scripts: logdet_matmul_stability_test.py:13-16 inputs,
    evaluated by the reference's logdet_matmul in fp64."""
    sys.path.insert(0, REF_SRC)
    from psiformer_torch.logdet_matmul import logdet_matmul

    g = torch.Generator().manual_seed(5)
    x1 = torch.randn(24, 3, 4, 4, generator=g, dtype=torch.float64)
    x2 = torch.randn(24, 3, 2, 2, generator=g, dtype=torch.float64)
    x1[0, 0] = 0.0          # exactly singular block -> jitter/clamp path
    x1[1, :, 3] = x1[1, :, 2]  # rank deficient in every determinant
    w = torch.softmax(torch.randn(3, generator=g, dtype=torch.float64), 0).unsqueeze(-1)
    la, sg = logdet_matmul(x1, x2, w)
    near = torch.tensor([[[[1.0, 2.0], [2.0001, 4.0]]]], dtype=torch.float64)
    one = torch.ones(1, 1, 1, 1, dtype=torch.float64)
    la2, sg2 = logdet_matmul(near, one, torch.ones(1, 1, dtype=torch.float64))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "logdet_kat.npz"),
                        x1=x1.numpy(), x2=x2.numpy(), w=w.numpy(), logabs=la.numpy(), sign=sg.numpy(),
                        near=near.numpy(), near_logabs=la2.numpy(), near_sign=sg2.numpy())
    print("logdet_kat written")


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count() or 1)
    only = set(sys.argv[1:])
    make_kats()
    for n, s in CASES.items():
        if not only or n in only:
            make_case(n, s, pinned=True)
    for n, s in MOLECULES.items():
        if not only or n in only:
            make_case(n, s, pinned=False)
