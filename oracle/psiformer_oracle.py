"""CPU oracle for the psiformer_torch VMC hot path.  TEST INFRASTRUCTURE ONLY.

This file is a functional (state_dict in, tensors out) restatement of the
reference's algorithm, written to be the *checker* for the CUDA path.  Only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl
reference`` legs of ``bench.py`` may import it.  The product package
``psiformer_torch_b200`` never does.

Parity status
-------------
* single nucleus at the origin (everything the reference can run): **pinned** —
  ``oracle/make_golden.py`` imports the reference from ``/root/reference/src``,
  runs it in fp32 and fp64 on fixed weights / walkers and commits the results
  under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks this file
  against those vectors.
* ``natom > 1`` (LiH, N2): **parity unpinned** — the reference hard-codes one
  nucleus at the origin (psiformer.py:133, hamiltonian.py:6-9).  The
  extension here follows SURVEY App. A.7 and reduces bit-for-bit to the pinned
  case for ``nuclei=((Z,(0,0,0)),)``.

Every function cites the reference lines it follows (paths relative to
``/root/reference/src/psiformer_torch/``).  The arithmetic that lives in
third-party code is ``torch.linalg.svd`` / ``torch.linalg.det`` and
``torch.autograd`` (torch 2.11.0, pinned by the reference's uv.lock:445-470);
torch is importable wherever this oracle runs, so those calls are executed, not
restated.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, Dict, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

# constants of logdet_matmul.py:15-18
DET_EPS = 1e-12
MIN_SINGULAR = 1e-6
OUTPUT_FLOOR = 1e-12
DET_JITTER = 1e-4
# hamiltonian.py:16 and jastrow.py:34,60
COULOMB_EPS = 1e-5
JASTROW_EPS = 1e-12
LN_EPS = 1e-5  # nn.LayerNorm default used at psiformer.py:86-87


@dataclass(frozen=True)
class OracleSystem:
    """Shape description of one wavefunction (mirrors config.py:6-16)."""

    n_layer: int
    n_head: int
    n_embd: int
    n_det: int
    n_up: int
    n_dn: int
    # ((Z, (x, y, z)), ...) ; the reference is nuclei=((Z,(0,0,0)),)
    nuclei: Tuple[Tuple[float, Tuple[float, float, float]], ...] = field(
        default=((2.0, (0.0, 0.0, 0.0)),)
    )

    @property
    def n_elec(self) -> int:
        return self.n_up + self.n_dn

    @property
    def natom(self) -> int:
        return len(self.nuclei)

    def charges(self, dtype=torch.float64) -> Tensor:
        return torch.tensor([z for z, _ in self.nuclei], dtype=dtype)

    def positions(self, dtype=torch.float64) -> Tensor:
        return torch.tensor([list(r) for _, r in self.nuclei], dtype=dtype)


def param_shapes(sysm: OracleSystem) -> Dict[str, Tuple[int, ...]]:
    """state_dict names and shapes, in registration order (SURVEY App. A.6;
    follows the module construction order of psiformer.py:204-217)."""
    d, L, K = sysm.n_embd, sysm.n_layer, sysm.n_det
    na = sysm.natom
    shapes: Dict[str, Tuple[int, ...]] = {}
    shapes["l_0.weight"] = (d, 4 * na)
    shapes["l_0.bias"] = (d,)
    for i in range(L):
        p = f"layers.{i}."
        shapes[p + "attn.c_attn.weight"] = (3 * d, d)
        shapes[p + "attn.c_attn.bias"] = (3 * d,)
        shapes[p + "attn.c_proj.weight"] = (d, d)
        shapes[p + "attn.c_proj.bias"] = (d,)
        shapes[p + "mlp.c_fc.weight"] = (4 * d, d)
        shapes[p + "mlp.c_fc.bias"] = (4 * d,)
        shapes[p + "mlp.c_proj.weight"] = (d, 4 * d)
        shapes[p + "mlp.c_proj.bias"] = (d,)
        shapes[p + "ln_1.weight"] = (d,)
        shapes[p + "ln_1.bias"] = (d,)
        shapes[p + "ln_2.weight"] = (d,)
        shapes[p + "ln_2.bias"] = (d,)
    shapes["orbital_head.det_logits"] = (K,)
    shapes["orbital_head.envelope_up.pi"] = (na, K * sysm.n_up)
    shapes["orbital_head.envelope_up.raw_sigma"] = (na, K * sysm.n_up)
    shapes["orbital_head.envelope_down.pi"] = (na, K * sysm.n_dn)
    shapes["orbital_head.envelope_down.raw_sigma"] = (na, K * sysm.n_dn)
    shapes["orbital_head.orb_up.weight"] = (K * sysm.n_up, d)
    shapes["orbital_head.orb_up.bias"] = (K * sysm.n_up,)
    shapes["orbital_head.orb_down.weight"] = (K * sysm.n_dn, d)
    shapes["orbital_head.orb_down.bias"] = (K * sysm.n_dn,)
    shapes["jastrow.alpha_anti"] = (1,)
    shapes["jastrow.alpha_par"] = (1,)
    return shapes


# --------------------------------------------------------------------------
# wavefunction
# --------------------------------------------------------------------------
def electron_features(sysm: OracleSystem, x: Tensor) -> Tuple[Tensor, Tensor]:
    """psiformer.py:233-234 and :243-244, with the App. A.7 per-nucleus
    generalisation.  Returns (features (B,N,4*natom), r_ae (B,N,natom,1))."""
    R = sysm.positions(x.dtype).to(x.device)              # (natom,3)
    disp = x[:, :, None, :] - R[None, None, :, :]          # (B,N,natom,3)
    r_ae = torch.linalg.norm(disp, dim=-1, keepdim=True)   # no epsilon, :233
    feats = torch.cat([disp, r_ae], dim=-1)                # [dx,dy,dz,r] per atom
    return feats.reshape(x.shape[0], x.shape[1], -1), r_ae


def _attention(sysm: OracleSystem, p: Dict[str, Tensor], pre: str, h: Tensor) -> Tensor:
    """psiformer.py:32-63 (explicit q k^T / sqrt(hd), softmax, @ v; no mask)."""
    B, N, d = h.shape
    H = sysm.n_head
    hd = d // H
    qkv = F.linear(h, p[pre + "c_attn.weight"], p[pre + "c_attn.bias"])
    q, k, v = qkv.split(d, dim=2)
    q = q.view(B, N, H, hd).transpose(1, 2)
    k = k.view(B, N, H, hd).transpose(1, 2)
    v = v.view(B, N, H, hd).transpose(1, 2)
    att = (q @ k.transpose(-2, -1)) / math.sqrt(hd)
    att = F.softmax(att, dim=-1)
    y = (att @ v).transpose(1, 2).contiguous().view(B, N, d)
    return F.linear(y, p[pre + "c_proj.weight"], p[pre + "c_proj.bias"])


def _mlp(p: Dict[str, Tensor], pre: str, h: Tensor) -> Tensor:
    """psiformer.py:73-77 (GELU with tanh approximation, :70)."""
    u = F.linear(h, p[pre + "c_fc.weight"], p[pre + "c_fc.bias"])
    u = F.gelu(u, approximate="tanh")
    return F.linear(u, p[pre + "c_proj.weight"], p[pre + "c_proj.bias"])


def backbone(sysm: OracleSystem, p: Dict[str, Tensor], feats: Tensor) -> Tensor:
    """psiformer.py:236-239 with Layer.forward :89-93 (pre-LN residual blocks)."""
    d = sysm.n_embd
    h = F.linear(feats, p["l_0.weight"], p["l_0.bias"])
    for i in range(sysm.n_layer):
        pre = f"layers.{i}."
        a = F.layer_norm(h, (d,), p[pre + "ln_1.weight"], p[pre + "ln_1.bias"], LN_EPS)
        h = h + _attention(sysm, p, pre + "attn.", a)
        m = F.layer_norm(h, (d,), p[pre + "ln_2.weight"], p[pre + "ln_2.bias"], LN_EPS)
        h = h + _mlp(p, pre + "mlp.", m)
    return h


def envelope(pi: Tensor, raw_sigma: Tensor, r_ae: Tensor) -> Tensor:
    """psiformer.py:106-120.  r_ae (B,n,natom,1) -> (B,n,K*n)."""
    sigma = torch.clamp(F.softplus(raw_sigma) + 1e-6, min=1e-3, max=1e3)
    pic = torch.clamp(pi, min=1e-3, max=1e3)
    return torch.sum(torch.exp(-r_ae * sigma) * pic, dim=2)


def orbital_matrices(sysm: OracleSystem, p: Dict[str, Tensor], h: Tensor,
                     r_ae: Tensor) -> Tuple[Tensor, Tensor]:
    """psiformer.py:150-175, 185-188: Phi[b,k,i,j] = out[b,i,k*n+j] (App. A.2)."""
    B = h.shape[0]
    K, nu, nd = sysm.n_det, sysm.n_up, sysm.n_dn
    o = "orbital_head."
    out_u = F.linear(h[:, :nu], p[o + "orb_up.weight"], p[o + "orb_up.bias"])
    out_u = out_u * envelope(p[o + "envelope_up.pi"], p[o + "envelope_up.raw_sigma"], r_ae[:, :nu])
    out_d = F.linear(h[:, nu:nu + nd], p[o + "orb_down.weight"], p[o + "orb_down.bias"])
    out_d = out_d * envelope(p[o + "envelope_down.pi"], p[o + "envelope_down.raw_sigma"], r_ae[:, nu:nu + nd])
    phi_u = out_u.view(B, nu, K, nu).transpose(1, 2)
    phi_d = out_d.view(B, nd, K, nd).transpose(1, 2)
    return phi_u, phi_d


def logdet_matmul_value(x1: Tensor, x2: Tensor, w: Tensor) -> Tuple[Tensor, Tensor]:
    """logdet_matmul.py:35-70: log|sum_k w_k det(x1_k) det(x2_k)| and its sign,
    by SVD with jitter 1e-4, singular clamp 1e-6, per-spin max shift, floor 1e-12.
    (LogDetMatmul.backward :94-120 differentiates this same function with
    autograd, so plain autograd through it is the reference's derivative.)"""
    def jit(a: Tensor) -> Tensor:  # _stabilize_matrix :21-27
        if a.shape[-1] != a.shape[-2]:
            return a
        return a + DET_JITTER * torch.eye(a.shape[-1], dtype=a.dtype, device=a.device)

    def slog(a: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
        u, s_raw, vh = torch.linalg.svd(a, full_matrices=False)
        s = torch.clamp(s_raw, min=MIN_SINGULAR)
        sgn = torch.sign(torch.linalg.det(u)) * torch.sign(torch.linalg.det(vh.transpose(-2, -1)))
        return torch.sum(torch.log(torch.clamp(s, min=DET_EPS)), dim=-1), sgn, s_raw

    l1, g1, _ = slog(jit(x1))
    l2, g2, _ = slog(jit(x2))
    m1 = l1.max(dim=-1, keepdim=True).values
    m2 = l2.max(dim=-1, keepdim=True).values
    det = torch.exp(l1 + l2 - m1 - m2) * (g1 * g2)
    out = det @ w
    log_out = torch.log(torch.clamp(out.abs(), min=OUTPUT_FLOOR)) + m1 + m2
    return log_out, torch.sign(out)


def min_singular_values(x1: Tensor, x2: Tensor) -> Tensor:
    """Smallest singular value over all jittered blocks of a walker, (B,).
    Not in the reference: exported so parity tests can mask walkers on which
    the 1e-6 clamp of logdet_matmul.py:50-51 is active."""
    def smin(a: Tensor) -> Tensor:
        a = a + DET_JITTER * torch.eye(a.shape[-1], dtype=a.dtype)
        return torch.linalg.svdvals(a).amin(dim=(-1, -2))
    return torch.minimum(smin(x1), smin(x2))


def jastrow(sysm: OracleSystem, p: Dict[str, Tensor], x: Tensor) -> Tensor:
    """jastrow.py:67-87 (with _same_spin_sum :20-42, _diff_spin_sum :45-65)."""
    if x.shape[1] < 2:
        raise ValueError("Jastrow requires at least two electrons.")
    nu, nd = sysm.n_up, sysm.n_dn
    a_par, a_anti = p["jastrow.alpha_par"], p["jastrow.alpha_anti"]

    def same(pos: Tensor) -> Tensor:
        n = pos.shape[1]
        if n < 2:
            return pos.new_zeros(pos.shape[0])
        diff = pos[:, :, None, :] - pos[:, None, :, :]
        dist = torch.sqrt(diff.pow(2).sum(-1) + JASTROW_EPS)
        i, j = torch.triu_indices(n, n, offset=1)
        return (-0.25 * a_par.pow(2) / (a_par + dist[:, i, j])).sum(1)

    up, dn = x[:, :nu], x[:, nu:nu + nd]
    tot = same(up) + same(dn)
    if nu > 0 and nd > 0:
        diff = up[:, :, None, :] - dn[:, None, :, :]
        dist = torch.sqrt(diff.pow(2).sum(-1) + JASTROW_EPS).reshape(x.shape[0], -1)
        tot = tot + (-0.5 * a_anti.pow(2) / (a_anti + dist)).sum(1)
    return tot


def log_psi_and_sign(sysm: OracleSystem, p: Dict[str, Tensor], x: Tensor) -> Tuple[Tensor, Tensor]:
    """PsiFormer.forward, psiformer.py:220-264, plus the sign the reference
    computes and drops (:191).  Raises like :228-231 and :256-257."""
    if x.dim() > 3:
        x = x.reshape(-1, x.size(-2), x.size(-1))
    if tuple(x.shape[1:]) != (sysm.n_elec, 3):
        raise ValueError("Input model shape mismatch", f"x shape: {tuple(x.shape)}")
    feats, r_ae = electron_features(sysm, x)
    h = backbone(sysm, p, feats)
    phi_u, phi_d = orbital_matrices(sysm, p, h, r_ae)
    w = torch.softmax(p["orbital_head.det_logits"], dim=-1).unsqueeze(-1)
    logdet, sign = logdet_matmul_value(phi_u, phi_d, w)
    logdet = logdet.squeeze(-1)
    if not torch.isfinite(logdet).all():
        raise ValueError("Non-finite log determinant detected")
    return logdet + jastrow(sysm, p, x), sign.squeeze(-1)


def log_psi(sysm: OracleSystem, p: Dict[str, Tensor], x: Tensor) -> Tensor:
    return log_psi_and_sign(sysm, p, x)[0]


# --------------------------------------------------------------------------
# Hamiltonian
# --------------------------------------------------------------------------
def potential(sysm: OracleSystem, x: Tensor) -> Tensor:
    """hamiltonian.py:15-35 generalised per App. A.7: softened e-n and e-e
    Coulomb, plus the constant (unsoftened) nuclear repulsion for natom > 1."""
    eps = COULOMB_EPS
    Z = sysm.charges(x.dtype)
    R = sysm.positions(x.dtype)
    d2 = (x[:, :, None, :] - R[None, None]).pow(2).sum(-1)       # (B,N,natom)
    r_i = torch.sqrt(d2 + eps)
    v = -(Z[None, None, :] * (1.0 / (r_i + eps))).sum(dim=(-1, -2))
    n = x.shape[1]
    if n >= 2:
        i, j = torch.triu_indices(n, n, offset=1)
        r_ij = torch.sqrt((x[:, i] - x[:, j]).pow(2).sum(-1) + eps)
        v = v + (1.0 / (r_ij + eps)).sum(-1)
    for a in range(sysm.natom):
        for b in range(a + 1, sysm.natom):
            v = v + Z[a] * Z[b] / torch.linalg.norm(R[a] - R[b])
    return v


def grad_and_laplacian(fn: Callable[[Tensor], Tensor], x: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    """hamiltonian.py:56-95: nested autograd; 3N second backward passes, of
    which only the diagonal entry is kept (:87-93).  Returns (y, grad, lap)."""
    xr = x.clone().detach().requires_grad_(True)
    y = fn(xr)
    (g,) = torch.autograd.grad(y, xr, grad_outputs=torch.ones_like(y), create_graph=True)
    gf = g.reshape(g.shape[0], -1)
    lap = torch.zeros_like(y)
    for j in range(gf.shape[1]):
        sec = torch.autograd.grad(gf[:, j].sum(), xr, retain_graph=True)[0]
        lap = lap + sec.reshape(g.shape[0], -1)[:, j]
    return y.detach(), g.detach(), lap.detach()


def local_energy_parts(sysm: OracleSystem, p: Dict[str, Tensor], x: Tensor) -> Dict[str, Tensor]:
    """hamiltonian.py:46-54: E_L = -1/2 (lap + |grad|^2) + V, with the pieces."""
    y, g, lap = grad_and_laplacian(lambda t: log_psi(sysm, p, t), x)
    v = potential(sysm, x)
    g2 = g.pow(2).reshape(g.shape[0], -1).sum(1)
    return {"logabs": y, "grad": g, "lap": lap, "pot": v, "e_loc": -0.5 * (lap + g2) + v}


def local_energy(sysm: OracleSystem, p: Dict[str, Tensor], x: Tensor) -> Tensor:
    return local_energy_parts(sysm, p, x)["e_loc"]


# --------------------------------------------------------------------------
# Metropolis-Hastings
# --------------------------------------------------------------------------
def mh_step(target: Callable[[Tensor], Tensor], state: Tensor, step_size: float,
            eps: Tensor, u: Tensor) -> Tuple[Tensor, Tensor]:
    """One all-electron Gaussian move with injected noise: mcmc.py:31-49.
    ``eps`` plays randn_like(state) (:33), ``u`` plays rand_like(alpha) (:42).
    Returns (new_state, accept_mask).  psi(current) is recomputed, as at :40."""
    trial = state + step_size * eps
    with torch.no_grad():
        alpha = 2 * (target(trial) - target(state))
        log_accept = torch.min(alpha, torch.zeros_like(alpha))
    accept = torch.log(u) < log_accept
    return torch.where(accept[:, None, None], trial, state), accept


def mh_run(target: Callable[[Tensor], Tensor], state: Tensor, step_size: float,
           eps: Tensor, u: Tensor) -> Tuple[Tensor, Tensor]:
    """mcmc.py:51-54 with injected noise eps[steps,B,N,3], u[steps,B]."""
    acc = []
    for s in range(max(1, eps.shape[0])):
        state, a = mh_step(target, state, step_size, eps[s], u[s])
        acc.append(a)
    return state, torch.stack(acc)


def mh_sampler(target: Callable[[Tensor], Tensor], state: Tensor | None, *, batch_size: int,
               n_elec: int, mc_len: int, burn_in: int, steps_per_sample: int, step_size: float,
               generator: torch.Generator | None = None) -> Tuple[Tensor, Tensor]:
    """MH.sampler, mcmc.py:56-84 (RNG order per SURVEY App. A.3)."""
    def run(st: Tensor, steps: int) -> Tensor:
        for _ in range(max(1, steps)):
            eps = torch.randn(st.shape, generator=generator, dtype=st.dtype)
            trial = st + step_size * eps
            with torch.no_grad():
                alpha = 2 * (target(trial) - target(st))
                la = torch.min(alpha, torch.zeros_like(alpha))
            u = torch.rand(alpha.shape, generator=generator, dtype=alpha.dtype)
            st = torch.where((torch.log(u) < la)[:, None, None], trial, st)
        return st

    if state is None:
        state = torch.randn(batch_size, n_elec, 3, generator=generator)
        state = run(state, burn_in)
    out = torch.empty(mc_len, batch_size, n_elec, 3, dtype=state.dtype)
    for i in range(mc_len):
        state = run(state, steps_per_sample)
        out[i] = state
    return out, state


# --------------------------------------------------------------------------
# deterministic synthetic weights (shared by fixtures, tests and bench)
# --------------------------------------------------------------------------
def synthetic_params(sysm: OracleSystem, seed: int) -> Dict[str, Tensor]:
    """Deterministic fp32 weights from numpy's PCG64 (platform independent),
    shaped like a lightly-trained reference model: nn.Linear-style
    U(-1/sqrt(fan_in), 1/sqrt(fan_in)); LayerNorm gamma 1+-0.1, beta +-0.05;
    envelope pi in (0.6,1.4), raw_sigma in (0.2,0.9); det_logits N(0,0.3);
    Jastrow alphas in (0.3,1.0).  Non-default LN/envelope values are
    deliberate: they exercise terms the default init (gamma=1, beta=0) hides."""
    import numpy as np

    rng = np.random.default_rng(seed)
    out: Dict[str, Tensor] = {}
    for name, shape in param_shapes(sysm).items():
        if name.endswith("ln_1.weight") or name.endswith("ln_2.weight"):
            a = 1.0 + 0.1 * rng.uniform(-1, 1, shape)
        elif name.endswith("ln_1.bias") or name.endswith("ln_2.bias"):
            a = 0.05 * rng.uniform(-1, 1, shape)
        elif name.endswith(".pi"):
            a = rng.uniform(0.6, 1.4, shape)
        elif name.endswith(".raw_sigma"):
            a = rng.uniform(0.2, 0.9, shape)
        elif name.endswith("det_logits"):
            a = 0.3 * rng.standard_normal(shape)
        elif name.startswith("jastrow."):
            a = rng.uniform(0.3, 1.0, shape)
        elif name.endswith(".weight"):
            bound = 1.0 / math.sqrt(shape[1])
            a = rng.uniform(-bound, bound, shape)
        elif name.endswith(".bias"):
            fan_in = param_shapes(sysm)[name[:-4] + "weight"][1]
            bound = 1.0 / math.sqrt(fan_in)
            a = rng.uniform(-bound, bound, shape)
        else:  # pragma: no cover
            raise KeyError(name)
        out[name] = torch.from_numpy(np.asarray(a, dtype=np.float32).copy())
    return out


def synthetic_walkers(sysm: OracleSystem, n_walkers: int, seed: int) -> Tensor:
    """x ~ N(0, I) like MH._init_state (mcmc.py:23-29); for molecules each
    electron is additionally shifted to a nucleus chosen round-robin."""
    import numpy as np

    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n_walkers, sysm.n_elec, 3)).astype(np.float32)
    if sysm.natom > 1:
        R = np.asarray([r for _, r in sysm.nuclei], dtype=np.float32)
        x = x + R[np.arange(sysm.n_elec) % sysm.natom][None]
    return torch.from_numpy(x)


def cast_params(p: Dict[str, Tensor], dtype: torch.dtype) -> Dict[str, Tensor]:
    return {k: v.to(dtype) for k, v in p.items()}


# named systems of BASELINE.json / SURVEY section 8 (C1..C5)
SYSTEMS: Dict[str, OracleSystem] = {
    "He": OracleSystem(1, 16, 64, 1, 1, 1, ((2.0, (0.0, 0.0, 0.0)),)),
    "Be": OracleSystem(4, 4, 256, 16, 2, 2, ((4.0, (0.0, 0.0, 0.0)),)),
    "LiH": OracleSystem(4, 4, 256, 16, 2, 2, ((3.0, (0.0, 0.0, 0.0)), (1.0, (0.0, 0.0, 3.015)))),
    "Ne": OracleSystem(4, 4, 256, 16, 5, 5, ((10.0, (0.0, 0.0, 0.0)),)),
    "N2": OracleSystem(4, 4, 256, 32, 7, 7, ((7.0, (0.0, 0.0, -2.0)), (7.0, (0.0, 0.0, 2.0)))),
}
