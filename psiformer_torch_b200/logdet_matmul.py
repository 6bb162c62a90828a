"""``logdet_matmul`` / ``LogDetMatmul``: the drop-in for logdet_matmul.py:73-137.

    log_abs, sign = logdet_matmul(x1, x2, w)      x1 (..., K, nu, nu), x2 (..., K, nd, nd), w (K, M) -> (..., M)

The value calls ``psif_slogdet_multi`` (pivoted Gauss-Jordan in fp64 per block, the reference's 1e-4 jitter, exact
1e-6 singular-value clamp on suspicious blocks, per-spin max shift, 1e-12 floor).  Like the reference's op
(logdet_matmul.py:94-120: its backward re-runs the value function under autograd with ``create_graph``) it is
differentiable TWICE: ``backward`` is ``psif_logdet_matmul_grad`` wrapped in an autograd Function of its own, whose
backward is ``psif_logdet_matmul_grad_grad`` (closed forms in csrc/logdet_grad.cuh; clamped singular values carry no
gradient, exactly like ``torch.clamp`` in the reference).  Inside ``PsiFormer`` the determinant never goes through this
function: the fused pipeline evaluates it together with its derivatives (csrc/slogdet.cuh).
"""
from __future__ import annotations

import torch
from torch import Tensor
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _lib as L


def _prep(x1: Tensor, x2: Tensor, w: Tensor):
    if x1.shape[:-2] != x2.shape[:-2]:
        raise ValueError("x1 and x2 must share leading dimensions.")
    if w.shape[0] != x1.shape[-3]:
        raise ValueError("Number of determinants must match w's first dimension.")
    if x1.device.type != "cuda":
        raise RuntimeError("psiformer_torch_b200.logdet_matmul runs on CUDA tensors only (no CPU fallback)")
    K, nu, nd = x1.shape[-3], x1.shape[-1], x2.shape[-1]
    a = x1.detach().to(torch.float32).reshape(-1, K, nu, nu).contiguous()
    b = x2.detach().to(torch.float32).reshape(-1, K, nd, nd).contiguous()
    w2 = w.detach().to(device=a.device, dtype=torch.float32).reshape(K, -1)
    return a, b, w2, K, nu, nd


def _value(x1: Tensor, x2: Tensor, w: Tensor):
    a, b, w2, K, nu, nd = _prep(x1, x2, w)
    lib = L.load()
    lead = x1.shape[:-3]
    B, M = a.shape[0], w2.shape[1]
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


class _LogDetMatmulGrad(Function):
    """(x1, x2, w, grad_log) -> (dx1, dx2, dw): the backward of ``LogDetMatmul`` as a differentiable op."""

    @staticmethod
    def forward(ctx, x1: Tensor, x2: Tensor, w: Tensor, grad_log: Tensor):
        a, b, w2, K, nu, nd = _prep(x1, x2, w)
        B, M = a.shape[0], w2.shape[1]
        g = grad_log.detach().to(torch.float32).reshape(B, M)
        d1, d2 = torch.zeros_like(a), torch.zeros_like(b)
        dw = torch.empty(K, M, dtype=torch.float32, device=a.device)
        lib = L.load()
        stream = torch.cuda.current_stream(a.device).cuda_stream
        with torch.cuda.device(a.device):
            for m in range(M):
                o1, o2 = torch.empty_like(a), torch.empty_like(b)
                ow = torch.empty(B, K, dtype=torch.float32, device=a.device)
                # named, so that they outlive the call: the caching allocator hands a freed temporary's block to the
                # very next allocation, and two temporaries built inside one argument list would alias
                wm, gm = w2[:, m].contiguous(), g[:, m].contiguous()
                L.check(lib.psif_logdet_matmul_grad(L.ptr(a), L.ptr(b), L.ptr(wm), L.ptr(gm),
                                                    B, K, nu, nd, L.ptr(o1), L.ptr(o2), L.ptr(ow), stream))
                d1 += o1
                d2 += o2
                dw[:, m] = ow.double().sum(0).float()
        ctx.save_for_backward(x1, x2, w, grad_log)
        return d1.reshape(x1.shape).to(x1.dtype), d2.reshape(x2.shape).to(x2.dtype), dw.reshape(w.shape).to(w.dtype)

    @staticmethod
    @once_differentiable
    def backward(ctx, v1: Tensor, v2: Tensor, vw: Tensor):
        x1, x2, w, grad_log = ctx.saved_tensors
        a, b, w2, K, nu, nd = _prep(x1, x2, w)
        B, M = a.shape[0], w2.shape[1]
        g = grad_log.detach().to(torch.float32).reshape(B, M)
        V1 = (torch.zeros_like(a) if v1 is None else v1.detach().to(torch.float32).reshape(a.shape).contiguous())
        V2 = (torch.zeros_like(b) if v2 is None else v2.detach().to(torch.float32).reshape(b.shape).contiguous())
        VW = (torch.zeros_like(w2) if vw is None else vw.detach().to(torch.float32).reshape(K, M))
        h1, h2 = torch.zeros_like(a), torch.zeros_like(b)
        hw = torch.empty(K, M, dtype=torch.float32, device=a.device)
        hg = torch.empty(B, M, dtype=torch.float32, device=a.device)
        lib = L.load()
        stream = torch.cuda.current_stream(a.device).cuda_stream
        with torch.cuda.device(a.device):
            for m in range(M):
                o1, o2 = torch.empty_like(a), torch.empty_like(b)
                ow = torch.empty(B, K, dtype=torch.float32, device=a.device)
                og = torch.empty(B, dtype=torch.float32, device=a.device)
                wm, gm, vwm = w2[:, m].contiguous(), g[:, m].contiguous(), VW[:, m].contiguous()
                L.check(lib.psif_logdet_matmul_grad_grad(
                    L.ptr(a), L.ptr(b), L.ptr(wm), L.ptr(gm), L.ptr(V1), L.ptr(V2),
                    L.ptr(vwm), B, K, nu, nd, L.ptr(og), L.ptr(o1), L.ptr(o2), L.ptr(ow), stream))
                h1 += o1
                h2 += o2
                hw[:, m] = ow.double().sum(0).float()
                hg[:, m] = og
        need = ctx.needs_input_grad
        return (h1.reshape(x1.shape).to(x1.dtype) if need[0] else None,
                h2.reshape(x2.shape).to(x2.dtype) if need[1] else None,
                hw.reshape(w.shape).to(w.dtype) if need[2] else None,
                hg.reshape(grad_log.shape).to(grad_log.dtype) if need[3] else None)


class LogDetMatmul(Function):
    """log|sum_k w_k det(x1_k) det(x2_k)| with sign tracking; twice differentiable like logdet_matmul.py:73-137."""

    @staticmethod
    def forward(ctx, x1: Tensor, x2: Tensor, w: Tensor):
        log_out, sign_out = _value(x1, x2, w)
        ctx.save_for_backward(x1, x2, w)
        ctx.mark_non_differentiable(sign_out)
        return log_out, sign_out

    @staticmethod
    def backward(ctx, grad_log: Tensor, grad_sign: Tensor):
        del grad_sign
        x1, x2, w = ctx.saved_tensors
        dx1, dx2, dw = _LogDetMatmulGrad.apply(x1, x2, w, grad_log)
        need = ctx.needs_input_grad
        return (dx1 if need[0] else None, dx2 if need[1] else None, dw if need[2] else None)


def logdet_matmul(x1: Tensor, x2: Tensor, w: Tensor):
    return LogDetMatmul.apply(x1, x2, w)


__all__ = ["logdet_matmul", "LogDetMatmul"]
