"""CPU check of the mathematics behind psif_logdet_matmul_grad / _grad_grad (csrc/logdet_math.cuh, host/device code)
against torch autograd through the reference's SVD formulation (oracle.logdet_matmul_value restates
logdet_matmul.py:35-70), including blocks with singular values under the 1e-6 clamp."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import psiformer_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("ldhost") / "logdet_host.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++",
                    os.path.join(ROOT, "tests", "host", "logdet_host.cpp"), "-o", so], check=True)
    return C.CDLL(so)


def _dp(a):
    return a.ctypes.data_as(C.c_void_p)


def _f(A):
    s = torch.linalg.svdvals(A)
    return torch.log(torch.clamp(s, min=O.MIN_SINGULAR)).sum()


def _block_with_singular_values(n, svals, g):
    U, _ = torch.linalg.qr(torch.randn(n, n, generator=g, dtype=torch.float64))
    V, _ = torch.linalg.qr(torch.randn(n, n, generator=g, dtype=torch.float64))
    return U @ torch.diag(torch.tensor(svals, dtype=torch.float64)) @ V.T


@pytest.mark.parametrize("n,svals", [(3, [1.3, 0.4, 0.05]), (4, [2.0, 0.7, 0.1, 3e-8]), (5, [1.0, 0.5, 0.2, 4e-7, 1e-9]),
                                     (7, [3.0, 1.1, 0.8, 0.3, 0.05, 1e-3, 2e-7]), (2, [0.9, 1e-10]), (1, [0.3])])
def test_block_gradient_and_hessian_follow_the_clamped_svd(host, n, svals):
    g = torch.Generator().manual_seed(n * 17 + len(svals))
    A = _block_with_singular_values(n, svals, g).requires_grad_(True)
    E = torch.randn(n, n, generator=g, dtype=torch.float64)
    f = _f(A)
    (G,) = torch.autograd.grad(f, A, create_graph=True)
    (H,) = torch.autograd.grad((G * E).sum(), A)
    An, En = np.ascontiguousarray(A.detach().numpy()), np.ascontiguousarray(E.numpy())
    fo, so, svd = C.c_double(), C.c_double(), C.c_int()
    Go, Ho = np.zeros((n, n)), np.zeros((n, n))
    host.ld_host_block(_dp(An), n, _dp(En), C.byref(fo), C.byref(so), _dp(Go), _dp(Ho), C.byref(svd))
    assert svd.value == (1 if min(svals) < O.MIN_SINGULAR else 0)
    assert abs(fo.value - f.item()) < 1e-10 * max(1.0, abs(f.item()))
    if n > 1 or svals[0] >= O.MIN_SINGULAR:
        assert so.value == torch.sign(torch.linalg.det(A.detach())).item()
    scale_g, scale_h = G.abs().max().item(), H.abs().max().item()
    assert np.abs(Go - G.detach().numpy()).max() < 1e-8 * scale_g
    assert np.abs(Ho - H.numpy()).max() < 1e-6 * scale_h


def _case(B, K, nu, nd, seed, near_singular):
    g = torch.Generator().manual_seed(seed)
    x1 = torch.randn(B, K, nu, nu, generator=g, dtype=torch.float64)
    x2 = torch.randn(B, K, nd, nd, generator=g, dtype=torch.float64)
    if near_singular:
        x1[0, 0] = _block_with_singular_values(nu, [1.0] * (nu - 1) + [3e-8], g) - O.DET_JITTER * torch.eye(nu, dtype=torch.float64)
        x2[1 % B, K - 1] = _block_with_singular_values(nd, [0.5] * (nd - 1) + [1e-9], g) - O.DET_JITTER * torch.eye(nd, dtype=torch.float64)
    w = torch.softmax(torch.randn(K, generator=g, dtype=torch.float64), 0)
    gbar = torch.randn(B, generator=g, dtype=torch.float64)
    # the kernels take fp32 inputs: compare on the fp32-rounded values
    return [t.float().double() for t in (x1, x2, w, gbar)]


@pytest.mark.parametrize("B,K,nu,nd,near", [(6, 3, 4, 2, False), (5, 16, 2, 2, False), (4, 4, 5, 5, True), (3, 2, 7, 7, True),
                                            (4, 1, 1, 1, False)])
def test_whole_op_backward_and_double_backward(host, B, K, nu, nd, near):
    x1, x2, w, gbar = _case(B, K, nu, nd, 100 + B + K, near)
    x1.requires_grad_(True); x2.requires_grad_(True); w.requires_grad_(True); gbar.requires_grad_(True)
    la, _ = O.logdet_matmul_value(x1, x2, w.unsqueeze(-1))
    d1, d2, dw = torch.autograd.grad(la.squeeze(-1), (x1, x2, w), grad_outputs=gbar, create_graph=True)
    gen = torch.Generator().manual_seed(7)
    v1 = torch.randn(x1.shape, generator=gen, dtype=torch.float64).float().double()
    v2 = torch.randn(x2.shape, generator=gen, dtype=torch.float64).float().double()
    vw = torch.randn(K, generator=gen, dtype=torch.float64).float().double()
    s = (d1 * v1).sum() + (d2 * v2).sum() + (dw * vw).sum()
    h1, h2, hw, hg = torch.autograd.grad(s, (x1, x2, w, gbar))

    f32 = lambda t: np.ascontiguousarray(t.detach().float().numpy())
    X1, X2, W, GB, V1, V2, VW = map(f32, (x1, x2, w, gbar, v1, v2, vw))
    o1, o2, ow = np.zeros_like(X1), np.zeros_like(X2), np.zeros((B, K), dtype=np.float32)
    host.ld_host_grad(_dp(X1), _dp(X2), _dp(W), _dp(GB), C.c_longlong(B), K, nu, nd, _dp(o1), _dp(o2), _dp(ow))

    def close(got, ref, tol=2e-5):
        ref = ref.detach().numpy()
        return np.abs(got - ref).max() <= tol * max(1e-30, np.abs(ref).max())
    assert close(o1, d1) and close(o2, d2) and close(ow.astype(np.float64).sum(0), dw)
    q1, q2, qw, qg = np.zeros_like(X1), np.zeros_like(X2), np.zeros((B, K), dtype=np.float32), np.zeros(B, dtype=np.float32)
    host.ld_host_grad_grad(_dp(X1), _dp(X2), _dp(W), _dp(GB), _dp(V1), _dp(V2), _dp(VW), C.c_longlong(B), K, nu, nd,
                           _dp(qg), _dp(q1), _dp(q2), _dp(qw))
    assert close(qg, hg) and close(q1, h1) and close(q2, h2) and close(qw.astype(np.float64).sum(0), hw)


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 6, 7, 8])
def test_register_sized_svd_and_inverse(host, n):
    """smallmat.cuh (host/device): the one-sided Jacobi SVD and the pivoted Gauss-Jordan inverse that the determinant kernels
    run per thread, on well-conditioned blocks and on blocks with a singular value far under the 1e-6 clamp."""
    g = torch.Generator().manual_seed(40 + n)
    for trial in range(4):
        svals = (0.2 + torch.rand(n, generator=g, dtype=torch.float64) * 2.0).tolist()
        if trial >= 2 and n > 1:
            svals[-1] = [3e-8, 1e-11][trial - 2]
        A = np.ascontiguousarray(_block_with_singular_values(n, svals, g).numpy())
        W, V = np.zeros((n, n)), np.zeros((n, n))
        assert host.sm_host_jacobi_svd(_dp(A), n, _dp(W), _dp(V)) == 0
        assert np.abs(W @ V.T - A).max() < 1e-13 * max(svals)
        assert np.abs(V.T @ V - np.eye(n)).max() < 1e-13
        s = np.linalg.norm(W, axis=0)
        ref = np.linalg.svd(A, compute_uv=False)
        assert np.abs(np.sort(s)[::-1] - ref).max() < 1e-12 * max(svals)
        gram = W.T @ W
        off = gram - np.diag(np.diag(gram))
        assert np.abs(off).max() < 1e-12 * max(svals) ** 2          # columns u_j s_j are orthogonal
        if trial < 2:
            X = np.zeros((n, n))
            ld, sg = C.c_double(), C.c_double()
            assert host.sm_host_gj_inverse(_dp(A), n, _dp(X), C.byref(ld), C.byref(sg)) == 0
            assert np.abs(X @ A - np.eye(n)).max() < 1e-10
            sign, logabs = np.linalg.slogdet(A)
            assert sg.value == sign and abs(ld.value - logabs) < 1e-11


@pytest.mark.parametrize("n,svals", [(4, [2.0, 0.7, 0.1, 3e-8]), (5, [1.0, 0.5, 0.2, 4e-7, 1e-9]), (7, [3.0, 1.1, 0.8, 0.3, 0.05, 1e-3, 2e-7]),
                                     (2, [0.9, 1e-10]), (3, [0.6, 2e-7, 5e-8])])
def test_clamped_block_forms_without_the_hessian(host, n, svals):
    """det_clamp_fixup_kernel never forms H[E]: with P = U^T E V it takes <G, E> = sum_unclamped P_ii / s_i and
    <H[E], E> = <M(P), P>.  Both must equal what torch autograd gives through svd -> clamp -> log."""
    g = torch.Generator().manual_seed(n * 31 + 5)
    A = _block_with_singular_values(n, svals, g).requires_grad_(True)
    E = torch.randn(n, n, generator=g, dtype=torch.float64)
    (G,) = torch.autograd.grad(_f(A), A, create_graph=True)
    gdot_ref = (G * E).sum()
    (H,) = torch.autograd.grad(gdot_ref, A)
    quad_ref = (H * E).sum().item()
    An, En = np.ascontiguousarray(A.detach().numpy()), np.ascontiguousarray(E.numpy())
    gd, qd = C.c_double(), C.c_double()
    assert host.ld_host_clamped_forms(_dp(An), n, _dp(En), C.byref(gd), C.byref(qd)) == 1
    assert abs(gd.value - gdot_ref.item()) < 1e-8 * max(1.0, abs(gdot_ref.item()))
    assert abs(qd.value - quad_ref) < 1e-6 * max(1.0, abs(quad_ref))
