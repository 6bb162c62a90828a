"""``logdet_matmul`` / ``LogDetMatmul``: the drop-in for logdet_matmul.py:73-137.

Value path only calls ``psif_slogdet_multi`` (pivoted Gauss-Jordan in fp64 per block, the
reference's 1e-4 jitter, exact 1e-6 singular-value clamp on suspicious blocks, per-spin max shift,
1e-12 floor).  Inside ``PsiFormer`` the determinant never goes through this function: the fused
pipeline evaluates it together with its derivatives (csrc/slogdet.cuh).
"""
from __future__ import annotations

import torch
from torch import Tensor
from torch.autograd import Function

from . import _lib as L


def _value(x1: Tensor, x2: Tensor, w: Tensor):
    if x1.shape[:-2] != x2.shape[:-2]:
        raise ValueError("x1 and x2 must share leading dimensions.")
    if w.shape[0] != x1.shape[-3]:
        raise ValueError("Number of determinants must match w's first dimension.")
    if x1.device.type != "cuda":
        raise RuntimeError("psiformer_torch_b200.logdet_matmul runs on CUDA tensors only (no CPU fallback)")
    lib = L.load()
    lead = x1.shape[:-3]
    K, nu, nd = x1.shape[-3], x1.shape[-1], x2.shape[-1]
    a = x1.detach().to(torch.float32).reshape(-1, K, nu, nu).contiguous()
    b = x2.detach().to(torch.float32).reshape(-1, K, nd, nd).contiguous()
    B = a.shape[0]
    w2 = w.detach().to(torch.float32).reshape(K, -1)
    M = w2.shape[1]
    logs, signs = [], []
    stream = torch.cuda.current_stream(a.device).cuda_stream
    with torch.cuda.device(a.device):
        for m in range(M):
            wm = w2[:, m].contiguous()
            la = torch.empty(B, dtype=torch.float32, device=a.device)
            sg = torch.empty_like(la)
            L.check(lib.psif_slogdet_multi(L.ptr(a), L.ptr(b), L.ptr(wm), B, K, nu, nd, L.ptr(la), L.ptr(sg), None, stream))
            logs.append(la)
            signs.append(sg)
    log_out = torch.stack(logs, -1).reshape(*lead, M).to(x1.dtype)
    sign_out = torch.stack(signs, -1).reshape(*lead, M).to(x1.dtype)
    return log_out, sign_out


class LogDetMatmul(Function):
    """log|sum_k w_k det(x1_k) det(x2_k)| with sign tracking (value on the GPU library)."""

    @staticmethod
    def forward(ctx, x1: Tensor, x2: Tensor, w: Tensor):
        return _value(x1, x2, w)

    @staticmethod
    def backward(ctx, grad_log: Tensor, grad_sign: Tensor):
        raise NotImplementedError(
            "derivatives of the determinant are produced by the fused pipeline (PsiFormer / Hamiltonian); "
            "the standalone logdet_matmul op is value-only in this build")


def logdet_matmul(x1: Tensor, x2: Tensor, w: Tensor):
    return LogDetMatmul.apply(x1, x2, w)


__all__ = ["logdet_matmul", "LogDetMatmul"]
