"""Forward-Laplacian restatement of the local energy, in plain torch (CPU).

TEST INFRASTRUCTURE ONLY (same rules as psiformer_oracle.py).  The reference
computes the kinetic term with 3N+2 autograd passes (hamiltonian.py:56-95).
The CUDA path instead carries, for every activation ``a``, the triple
``(a, grad a in R^{3N}, lap a)`` through the network in one pass.  This file
states those propagation rules stage by stage in torch so that

* the rules themselves are checked against ``psiformer_oracle`` (nested
  autograd, itself pinned to the reference) in ``tests/test_forward_laplacian.py``;
* every CUDA stage kernel has a same-layout CPU checker.

Payload layout (identical to the CUDA kernels): ``P[b, i, c, e]`` with
``c = 0`` the value, ``c = 1 + 3 j + alpha`` the derivative with respect to
coordinate alpha of electron j, and ``c = 3N + 1`` the Laplacian.
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import torch
import torch.nn.functional as F

from . import psiformer_oracle as O

Tensor = torch.Tensor


def n_channels(n_elec: int) -> int:
    return 3 * n_elec + 2


def embed_payload(sysm: O.OracleSystem, p: Dict[str, Tensor], x: Tensor) -> Tensor:
    """Features [r_i - R_I, |r_i - R_I|] (psiformer.py:233-234) and l_0 (:236)
    with their derivatives: grad|r| = r/|r|, lap|r| = 2/|r|; only electron i's
    own three coordinate channels are non-zero for token i."""
    B, N, _ = x.shape
    C = n_channels(N)
    na = sysm.natom
    R = sysm.positions(x.dtype)
    disp = x[:, :, None, :] - R[None, None]                    # (B,N,na,3)
    r = torch.linalg.norm(disp, dim=-1)                        # (B,N,na)
    f = torch.zeros(B, N, C, 4 * na, dtype=x.dtype)
    for a in range(na):
        f[:, :, 0, 4 * a:4 * a + 3] = disp[:, :, a]
        f[:, :, 0, 4 * a + 3] = r[:, :, a]
        f[:, :, C - 1, 4 * a + 3] = 2.0 / r[:, :, a]
        for i in range(N):
            for al in range(3):
                f[:, i, 1 + 3 * i + al, 4 * a + al] = 1.0
                f[:, i, 1 + 3 * i + al, 4 * a + 3] = disp[:, i, a, al] / r[:, i, a]
    return linear_payload(f, p["l_0.weight"], p["l_0.bias"])


def linear_payload(P: Tensor, W: Tensor, b: Tensor | None) -> Tensor:
    """y = W a + b: every channel goes through W, the bias only on channel 0."""
    Y = P @ W.t()
    if b is not None:
        Y[:, :, 0, :] = Y[:, :, 0, :] + b
    return Y


def layernorm_payload(P: Tensor, gamma: Tensor, beta: Tensor, eps: float = O.LN_EPS) -> Tensor:
    """nn.LayerNorm(d) (psiformer.py:86-87), App. B LayerNorm row."""
    a = P[:, :, 0, :]
    c = a - a.mean(-1, keepdim=True)
    s = (c.pow(2).mean(-1, keepdim=True) + eps).rsqrt()
    ah = c * s
    T = P[:, :, 1:-1, :]
    cd = T - T.mean(-1, keepdim=True)                          # (B,N,3N,d)
    m = (ah[:, :, None, :] * cd).mean(-1, keepdim=True)
    w = cd - ah[:, :, None, :] * m
    gT = s[:, :, None, :] * w
    q = cd.pow(2).mean(-1, keepdim=True)
    s2 = (s * s)[:, :, None, :]
    corr = (-2.0 * s2 * m * w - s2 * ah[:, :, None, :] * (q - m * m)).sum(2)
    lapa = P[:, :, -1, :]
    lc = lapa - lapa.mean(-1, keepdim=True)
    lap = s * (lc - ah * (ah * lc).mean(-1, keepdim=True)) + corr
    out = torch.empty_like(P)
    out[:, :, 0, :] = gamma * ah + beta
    out[:, :, 1:-1, :] = gamma * gT
    out[:, :, -1, :] = gamma * lap
    return out


def attention_payload(QKV: Tensor, n_head: int) -> Tensor:
    """softmax(q k^T / sqrt(hd)) v per head (psiformer.py:42-62) with the
    bilinear and softmax rules of App. B.  QKV: (B,N,C,3d) -> (B,N,C,d)."""
    B, N, C, d3 = QKV.shape
    d = d3 // 3
    hd = d // n_head
    scale = 1.0 / math.sqrt(hd)

    def heads(t: Tensor) -> Tensor:  # (B,N,C,d) -> (B,H,C,N,hd)
        return t.reshape(B, N, C, n_head, hd).permute(0, 3, 2, 1, 4)

    q, k, v = (heads(t) for t in QKV.split(d, dim=-1))
    q0, k0, v0 = q[:, :, 0], k[:, :, 0], v[:, :, 0]            # (B,H,N,hd)
    qT, kT, vT = q[:, :, 1:-1], k[:, :, 1:-1], v[:, :, 1:-1]  # (B,H,3N,N,hd)
    qL, kL, vL = q[:, :, -1], k[:, :, -1], v[:, :, -1]
    s0 = (q0 @ k0.transpose(-1, -2)) * scale                    # (B,H,N,N)
    sT = (qT @ k0[:, :, None].transpose(-1, -2) + q0[:, :, None] @ kT.transpose(-1, -2)) * scale
    sL = (qL @ k0.transpose(-1, -2) + q0 @ kL.transpose(-1, -2)
          + 2.0 * (qT @ kT.transpose(-1, -2)).sum(2)) * scale
    p0 = torch.softmax(s0, dim=-1)
    mbar = (p0[:, :, None] * sT).sum(-1, keepdim=True)          # (B,H,3N,N,1)
    dev = sT - mbar
    pT = p0[:, :, None] * dev
    quad = dev.pow(2).sum(2)                                     # (B,H,N,N)
    pL = p0 * ((sL - (p0 * sL).sum(-1, keepdim=True)) + quad - (p0 * quad).sum(-1, keepdim=True))
    y0 = p0 @ v0
    yT = pT @ v0[:, :, None] + p0[:, :, None] @ vT
    yL = pL @ v0 + p0 @ vL + 2.0 * (pT @ vT).sum(2)
    y = torch.cat([y0[:, :, None], yT, yL[:, :, None]], dim=2)  # (B,H,C,N,hd)
    return y.permute(0, 3, 2, 1, 4).reshape(B, N, C, d)


def attention_first_layer_payload(QKV5: Tensor, n_head: int) -> Tensor:
    """The same rule for the FIRST layer, where token i depends on x_i only: QKV5 (B,N,5,3d) holds the value row, the three
    own tangents d/dx_{i,alpha} and the Laplacian row of every token (all other tangent rows are zero), the result is the
    dense payload (B,N,3N+2,d).  Restates, in closed form, what attention_first_layer.cuh computes: for channel
    c = (j, alpha)   A_c[k] = scale dq_j[alpha].k0_k,  T_c[i] = scale q0_i.dk_j[alpha]  and
        i != j:  y_i[c] = P_ij (T_c[i] (v0_j - y_i[0]) + dv_j[alpha])
        i == j:  dS_jk = A_c[k] + delta_kj T_c[j],  dP_jk = P_jk (dS_jk - sum_m P_jm dS_jm),  y_j[c] = sum_k dP_jk v0_k + P_jj dv_j[alpha]
    and the Laplacian row from quad_ik = sum_c (dS_ik[c] - m_i[c])^2 with the foreign channels summed analytically."""
    B, N, five, d3 = QKV5.shape
    assert five == 5
    d = d3 // 3
    hd = d // n_head
    scale = 1.0 / math.sqrt(hd)

    def heads(t: Tensor) -> Tensor:  # (B,N,5,d) -> (B,H,5,N,hd)
        return t.reshape(B, N, 5, n_head, hd).permute(0, 3, 2, 1, 4)

    q, k, v = (heads(t) for t in QKV5.split(d, dim=-1))
    q0, k0, v0 = q[:, :, 0], k[:, :, 0], v[:, :, 0]                    # (B,H,N,hd)
    dq, dk, dv = q[:, :, 1:4], k[:, :, 1:4], v[:, :, 1:4]              # (B,H,3,N,hd): [alpha][j]
    lq, lk, lv = q[:, :, 4], k[:, :, 4], v[:, :, 4]
    P = torch.softmax((q0 @ k0.transpose(-1, -2)) * scale, dim=-1)     # (B,H,N,N)
    A = (dq @ k0[:, :, None].transpose(-1, -2)) * scale                # (B,H,3,j,k)
    T = (dk @ q0[:, :, None].transpose(-1, -2)) * scale                # (B,H,3,j,i)
    X = 2.0 * scale * (dq * dk).sum(-1).sum(2)                         # (B,H,j)
    eye = torch.eye(N, dtype=QKV5.dtype)
    LS = (lq @ k0.transpose(-1, -2) + q0 @ lk.transpose(-1, -2)) * scale + torch.diag_embed(X)
    Tdiag = torch.diagonal(T, dim1=-2, dim2=-1)                        # (B,H,3,j): T_c[j]
    dS_own = A + Tdiag[..., None] * eye                                # (B,H,3,j,k)
    m_own = (P[:, :, None] * dS_own).sum(-1, keepdim=True)
    dP_own = P[:, :, None] * (dS_own - m_own)                          # (B,H,3,j,k)
    y0 = P @ v0                                                        # (B,H,N,hd)
    # tangent rows y[b,h,i,(j,alpha)]
    Pij = P[:, :, None]                                                # (B,H,1,i,j)
    Tij = T.transpose(-1, -2)                                          # (B,H,3,i,j)
    foreign = Pij[..., None] * (Tij[..., None] * (v0[:, :, None, None] - y0[:, :, None, :, None]) + dv[:, :, :, None])
    own = dP_own @ v0[:, :, None] + torch.diagonal(P, dim1=-2, dim2=-1)[:, :, None, :, None] * dv     # (B,H,3,j,hd)
    yT = foreign.clone()                                               # (B,H,3,i,j,hd)
    idx = torch.arange(N)
    yT[:, :, :, idx, idx] = own
    # Laplacian row
    TT = T.pow(2).sum(2)                                               # (B,H,j,i)
    off = 1.0 - eye
    Bi = (TT.transpose(-1, -2) * P.pow(2) * off).sum(-1, keepdim=True)               # (B,H,i,1)
    quad = Bi + off * TT.transpose(-1, -2) * (1.0 - 2.0 * P) + (dS_own - m_own).pow(2).sum(2)
    WL = P * ((LS - (P * LS).sum(-1, keepdim=True)) + quad - (P * quad).sum(-1, keepdim=True))
    CX = 2.0 * Pij * Tij * (1.0 - Pij)                                 # (B,H,3,i,j), foreign entries
    CX_own = 2.0 * torch.diagonal(dP_own, dim1=-2, dim2=-1)            # (B,H,3,j)
    CX[:, :, :, idx, idx] = CX_own
    yL = WL @ v0 + P @ lv + torch.einsum("bhaij,bhajd->bhid", CX, dv)
    # assemble (B,H,C,N,hd): channel 1 + 3 j + alpha
    yTc = yT.permute(0, 1, 4, 2, 3, 5).reshape(B, n_head, 3 * N, N, hd)               # [(j,alpha)][i]
    y = torch.cat([y0[:, :, None], yTc, yL[:, :, None]], dim=2)
    return y.permute(0, 3, 2, 1, 4).reshape(B, N, 3 * N + 2, d)


def gelu_payload(P: Tensor) -> Tensor:
    """GELU(tanh) (psiformer.py:70,75): grad = g' grad u, lap = g' lap u + g'' |grad u|^2."""
    u = P[:, :, 0, :]
    kap = math.sqrt(2.0 / math.pi)
    t = torch.tanh(kap * (u + 0.044715 * u ** 3))
    q = kap * (1.0 + 3 * 0.044715 * u * u)
    sech2 = 1.0 - t * t
    g1 = 0.5 * (1.0 + t) + 0.5 * u * sech2 * q
    g2 = sech2 * q + 0.5 * u * sech2 * (kap * 6 * 0.044715 * u - 2.0 * t * q * q)
    out = torch.empty_like(P)
    out[:, :, 0, :] = 0.5 * u * (1.0 + t)
    out[:, :, 1:-1, :] = g1[:, :, None, :] * P[:, :, 1:-1, :]
    out[:, :, -1, :] = g1 * P[:, :, -1, :] + g2 * P[:, :, 1:-1, :].pow(2).sum(2)
    return out


def backbone_payload(sysm: O.OracleSystem, p: Dict[str, Tensor], x: Tensor, taps: dict | None = None) -> Tensor:
    """psiformer.py:236-239 on payloads; ``taps`` collects every intermediate."""
    h = embed_payload(sysm, p, x)
    if taps is not None:
        taps["embed"] = h
    for i in range(sysm.n_layer):
        pre = f"layers.{i}."
        a = layernorm_payload(h, p[pre + "ln_1.weight"], p[pre + "ln_1.bias"])
        qkv = linear_payload(a, p[pre + "attn.c_attn.weight"], p[pre + "attn.c_attn.bias"])
        y = attention_payload(qkv, sysm.n_head)
        h = h + linear_payload(y, p[pre + "attn.c_proj.weight"], p[pre + "attn.c_proj.bias"])
        m = layernorm_payload(h, p[pre + "ln_2.weight"], p[pre + "ln_2.bias"])
        u = linear_payload(m, p[pre + "mlp.c_fc.weight"], p[pre + "mlp.c_fc.bias"])
        gl = gelu_payload(u)
        h = h + linear_payload(gl, p[pre + "mlp.c_proj.weight"], p[pre + "mlp.c_proj.bias"])
        if taps is not None:
            taps[f"ln1.{i}"], taps[f"qkv.{i}"], taps[f"att.{i}"] = a, qkv, y
            taps[f"ln2.{i}"], taps[f"fc.{i}"], taps[f"gelu.{i}"], taps[f"h.{i}"] = m, u, gl, h
    return h


def envelope_payload(sysm: O.OracleSystem, pi: Tensor, raw_sigma: Tensor, x: Tensor, tok0: int, n: int) -> Tensor:
    """sum_I pi exp(-sigma r_iI) (psiformer.py:115-120) for tokens tok0..tok0+n,
    as a payload (B,n,C,K*n): grad = -sigma e r_hat, lap = e (sigma^2 - 2 sigma / r)."""
    B, N, _ = x.shape
    C = n_channels(N)
    sigma = torch.clamp(F.softplus(raw_sigma) + 1e-6, min=1e-3, max=1e3)   # (na,ch)
    pic = torch.clamp(pi, min=1e-3, max=1e3)
    R = sysm.positions(x.dtype)
    out = torch.zeros(B, n, C, pi.shape[1], dtype=x.dtype)
    for t in range(n):
        i = tok0 + t
        disp = x[:, i, None, :] - R[None]                        # (B,na,3)
        r = torch.linalg.norm(disp, dim=-1)                      # (B,na)
        e = pic[None] * torch.exp(-r[:, :, None] * sigma[None])  # (B,na,ch)
        out[:, t, 0] = e.sum(1)
        for al in range(3):
            out[:, t, 1 + 3 * i + al] = (-sigma[None] * e * (disp[:, :, al] / r)[:, :, None]).sum(1)
        out[:, t, C - 1] = (e * (sigma[None] ** 2 - 2.0 * sigma[None] / r[:, :, None])).sum(1)
    return out


def product_payload(a: Tensor, b: Tensor) -> Tensor:
    """Elementwise product rule: lap(uv) = u lap v + v lap u + 2 grad u . grad v."""
    out = torch.empty_like(a)
    out[:, :, 0] = a[:, :, 0] * b[:, :, 0]
    out[:, :, 1:-1] = a[:, :, 1:-1] * b[:, :, :1] + a[:, :, :1] * b[:, :, 1:-1]
    out[:, :, -1] = (a[:, :, -1] * b[:, :, 0] + a[:, :, 0] * b[:, :, -1]
                     + 2.0 * (a[:, :, 1:-1] * b[:, :, 1:-1]).sum(2))
    return out


def orbital_payload(sysm: O.OracleSystem, p: Dict[str, Tensor], h: Tensor, x: Tensor) -> Tuple[Tensor, Tensor]:
    """psiformer.py:150-175 on payloads -> Phi_sigma (B,K,n,n,C) (row = electron)."""
    B, N, C, _ = h.shape
    K, nu, nd = sysm.n_det, sysm.n_up, sysm.n_dn
    o = "orbital_head."
    lu = linear_payload(h[:, :nu], p[o + "orb_up.weight"], p[o + "orb_up.bias"])
    ld = linear_payload(h[:, nu:nu + nd], p[o + "orb_down.weight"], p[o + "orb_down.bias"])
    eu = envelope_payload(sysm, p[o + "envelope_up.pi"], p[o + "envelope_up.raw_sigma"], x, 0, nu)
    ed = envelope_payload(sysm, p[o + "envelope_down.pi"], p[o + "envelope_down.raw_sigma"], x, nu, nd)
    pu = product_payload(lu, eu).reshape(B, nu, C, K, nu).permute(0, 3, 1, 4, 2)
    pd = product_payload(ld, ed).reshape(B, nd, C, K, nd).permute(0, 3, 1, 4, 2)
    return pu, pd


def slogdet_payload(phi_u: Tensor, phi_d: Tensor, w: Tensor) -> Dict[str, Tensor]:
    """log|sum_k w_k det(A_up_k) det(A_dn_k)|, A = Phi + 1e-4 I
    (logdet_matmul.py:41-42, 58-69), by LU; App. B rows "log|det A|" and
    "multi-det combine".  The 1e-6 singular-value clamp (:50-51) is *not*
    applied here: walkers where it is active are identified by ``smin``."""
    def per_spin(phi: Tensor):
        n = phi.shape[-2]
        A = phi[..., 0] + O.DET_JITTER * torch.eye(n, dtype=phi.dtype)
        sgn, ld = torch.linalg.slogdet(A)
        Ai = torch.linalg.inv(A)
        dA = phi[..., 1:-1].permute(0, 1, 4, 2, 3)                # (B,K,3N,n,n)
        M = Ai[:, :, None] @ dA
        g = M.diagonal(dim1=-1, dim2=-2).sum(-1)                   # (B,K,3N)
        tr2 = (M * M.transpose(-1, -2)).sum(dim=(-1, -2))          # tr(M M)
        lap = (Ai * phi[..., -1].transpose(-1, -2)).sum(dim=(-1, -2)) - tr2.sum(-1)
        return ld, sgn, g, lap

    l1, s1, g1, p1 = per_spin(phi_u)
    l2, s2, g2, p2 = per_spin(phi_d)
    m1 = l1.max(-1, keepdim=True).values
    m2 = l2.max(-1, keepdim=True).values
    wk = w.reshape(1, -1)
    D = wk * s1 * s2 * torch.exp(l1 + l2 - m1 - m2)               # (B,K)
    S = D.sum(-1, keepdim=True)
    ck = D / S
    gk = g1 + g2
    G = (ck[:, :, None] * gk).sum(1)                               # (B,3N)
    lap = (ck * (p1 + p2 + gk.pow(2).sum(-1))).sum(1) - G.pow(2).sum(-1)
    logabs = torch.log(torch.clamp(S.abs(), min=O.OUTPUT_FLOOR)).squeeze(-1) + (m1 + m2).squeeze(-1)
    return {"logabs": logabs, "sign": torch.sign(S).squeeze(-1), "grad": G, "lap": lap}


def jastrow_payload(sysm: O.OracleSystem, p: Dict[str, Tensor], x: Tensor) -> Dict[str, Tensor]:
    """jastrow.py:67-87 with closed-form gradient / Laplacian (App. B Jastrow row)."""
    B, N, _ = x.shape
    nu = sysm.n_up
    val = torch.zeros(B, dtype=x.dtype)
    grad = torch.zeros(B, N, 3, dtype=x.dtype)
    lap = torch.zeros(B, dtype=x.dtype)
    for i in range(N):
        for j in range(i + 1, N):
            same = (i < nu) == (j < nu)
            c, al = (-0.25, p["jastrow.alpha_par"]) if same else (-0.5, p["jastrow.alpha_anti"])
            al = al.to(x.dtype).reshape(())
            dvec = x[:, i] - x[:, j]
            d2 = dvec.pow(2).sum(-1)
            rt = torch.sqrt(d2 + O.JASTROW_EPS)
            den = al + rt
            f1 = -c * al * al / den ** 2
            f2 = 2.0 * c * al * al / den ** 3
            val = val + c * al * al / den
            gi = (f1 / rt)[:, None] * dvec
            grad[:, i] += gi
            grad[:, j] -= gi
            lap = lap + 2.0 * (f2 * d2 / rt ** 2 + f1 * (3.0 / rt - d2 / rt ** 3))
    return {"val": val, "grad": grad.reshape(B, -1), "lap": lap}


def local_energy_forward(sysm: O.OracleSystem, p: Dict[str, Tensor], x: Tensor, taps: dict | None = None) -> Dict[str, Tensor]:
    """E_L = -1/2 (lap + |grad|^2) + V (hamiltonian.py:52-54) in one forward pass."""
    h = backbone_payload(sysm, p, x, taps)
    phi_u, phi_d = orbital_payload(sysm, p, h, x)
    w = torch.softmax(p["orbital_head.det_logits"], dim=-1)
    det = slogdet_payload(phi_u, phi_d, w)
    jas = jastrow_payload(sysm, p, x)
    if taps is not None:
        taps["phi_up"], taps["phi_dn"] = phi_u, phi_d
    g = det["grad"] + jas["grad"]
    lap = det["lap"] + jas["lap"]
    v = O.potential(sysm, x)
    return {"logabs": det["logabs"] + jas["val"], "sign": det["sign"], "grad": g.reshape(x.shape),
            "lap": lap, "pot": v, "e_loc": -0.5 * (lap + g.pow(2).sum(-1)) + v}
