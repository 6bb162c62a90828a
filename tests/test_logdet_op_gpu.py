"""The standalone ``logdet_matmul`` op (the drop-in for logdet_matmul.py:73-137) on the GPU: value, backward and
double backward through the C ABI against torch autograd through the oracle's restatement of the reference's SVD
formulation, including the reference's own stability case (tests/logdet_matmul_stability_test.py:6-25) and blocks on
which the 1e-6 singular-value clamp is active."""
import pytest
import torch

from oracle import psiformer_oracle as O

pytestmark = pytest.mark.gpu


def _rel(got, ref):
    ref = ref.detach().double().cpu()
    return ((got.detach().double().cpu() - ref).abs().max() / ref.abs().max().clamp_min(1e-30)).item()


def _near_singular(n, smin, g):
    U, _ = torch.linalg.qr(torch.randn(n, n, generator=g, dtype=torch.float64))
    V, _ = torch.linalg.qr(torch.randn(n, n, generator=g, dtype=torch.float64))
    s = torch.linspace(1.0, 0.3, n, dtype=torch.float64)
    s[-1] = smin
    return (U @ torch.diag(s) @ V.T - O.DET_JITTER * torch.eye(n, dtype=torch.float64)).float()


@pytest.mark.parametrize("B,K,nu,nd,M,near", [(33, 4, 4, 2, 1, False), (70, 16, 2, 2, 3, False), (9, 16, 5, 5, 1, True),
                                              (5, 32, 7, 7, 2, True), (64, 1, 1, 1, 1, False)])
def test_value_backward_and_double_backward(B, K, nu, nd, M, near):
    from psiformer_torch_b200.logdet_matmul import logdet_matmul
    g = torch.Generator().manual_seed(B * 7 + K)
    x1 = torch.randn(B, K, nu, nu, generator=g)
    x2 = torch.randn(B, K, nd, nd, generator=g)
    if near:
        x1[0, 0] = _near_singular(nu, 3e-8, g)
        x2[1, K - 1] = _near_singular(nd, 1e-9, g)
    w = torch.softmax(torch.randn(K, M, generator=g), 0)
    gbar = torch.randn(B, M, generator=g)
    v1, v2, vw = torch.randn(x1.shape, generator=g), torch.randn(x2.shape, generator=g), torch.randn(K, M, generator=g)

    def run(fn, dev, dt):
        a, b, c, gb = (t.to(dev, dt).requires_grad_(True) for t in (x1, x2, w, gbar))
        la, sg = fn(a, b, c)
        d1, d2, dw = torch.autograd.grad(la, (a, b, c), grad_outputs=gb, create_graph=True)
        s = (d1 * v1.to(dev, dt)).sum() + (d2 * v2.to(dev, dt)).sum() + (dw * vw.to(dev, dt)).sum()
        h = torch.autograd.grad(s, (a, b, c, gb))
        return la, sg, (d1, d2, dw), h
    la, sg, d, h = run(logdet_matmul, "cuda", torch.float32)
    rla, rsg, rd, rh = run(O.logdet_matmul_value, "cpu", torch.float64)
    assert torch.allclose(la.double().cpu(), rla, rtol=1e-5, atol=1e-5) and torch.equal(sg.double().cpu(), rsg)
    for got, ref in zip(d, rd):
        assert _rel(got, ref) < 2e-5
    for got, ref in zip(h, rh):
        assert _rel(got, ref) < 5e-5


def test_reference_stability_case_is_finite_and_matches():
    """logdet_matmul_stability_test.py:6-25: value, first and second derivative of a nearly singular 2x2."""
    from psiformer_torch_b200.logdet_matmul import logdet_matmul
    base = torch.tensor([[1.0, 2.0], [2.0001, 4.0]])

    def run(fn, dev, dt):
        b = base.to(dev, dt).requires_grad_(True)
        x1 = b[None, None]
        x2 = torch.ones(1, 1, 1, 1, device=dev, dtype=dt, requires_grad=True)
        out, _ = fn(x1, x2, torch.ones(1, 1, device=dev, dtype=dt))
        g = torch.autograd.grad(out.sum(), b, create_graph=True)[0]
        gg = torch.autograd.grad(g.sum(), b)[0]
        return out, g, gg
    out, g, gg = run(logdet_matmul, "cuda", torch.float32)
    assert torch.isfinite(out).all() and torch.isfinite(g).all() and torch.isfinite(gg).all()
    ro, rg, rgg = run(O.logdet_matmul_value, "cpu", torch.float64)
    assert _rel(out, ro) < 1e-5 and _rel(g, rg) < 1e-4 and _rel(gg, rgg) < 1e-3


def test_shape_errors_match_the_reference():
    from psiformer_torch_b200.logdet_matmul import logdet_matmul
    with pytest.raises(ValueError):
        logdet_matmul(torch.zeros(2, 3, 2, 2, device="cuda"), torch.zeros(4, 3, 2, 2, device="cuda"), torch.ones(3, 1, device="cuda"))
    with pytest.raises(ValueError):
        logdet_matmul(torch.zeros(2, 3, 2, 2, device="cuda"), torch.zeros(2, 3, 2, 2, device="cuda"), torch.ones(4, 1, device="cuda"))
