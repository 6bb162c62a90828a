"""The forward-Laplacian propagation rules (oracle/forward_laplacian.py) against
the nested-autograd oracle and the reference's fp64 golden vectors."""
import pytest
import torch

from conftest import ALL_CASES
from oracle import forward_laplacian as FL
from oracle import psiformer_oracle as O


@pytest.mark.parametrize("name", ["debug", "he_small", "large", "be", "lih"])
def test_forward_laplacian_matches_golden_fp64(golden, name):
    sysm, params, data = golden(name)
    with torch.no_grad():
        out = FL.local_energy_forward(sysm, O.cast_params(params, torch.float64), data["x"].double())
    ok = data["ref64_smin"] > 10 * O.MIN_SINGULAR   # clamp of logdet_matmul.py:50-51 inactive
    assert ok.float().mean() > 0.8, "fixture walkers should mostly sit away from the singular-value clamp"
    for k, ref in (("logabs", "ref64_logabs"), ("lap", "ref64_lap"), ("pot", "ref64_pot"), ("e_loc", "ref64_eloc")):
        scale = max(1.0, data[ref].abs().max().item())
        assert (out[k] - data[ref])[ok].abs().max().item() <= 1e-8 * scale, k
    gscale = max(1.0, data["ref64_grad"].abs().max().item())
    assert (out["grad"] - data["ref64_grad"])[ok].abs().max().item() <= 1e-8 * gscale
    assert torch.equal(out["sign"], data["ref64_sign"])


def test_stage_rules_against_autograd_hessian():
    """LayerNorm / softmax-attention / GELU rules vs torch.autograd.functional on a tiny case."""
    torch.manual_seed(0)
    N, d, H = 2, 8, 2
    C = FL.n_channels(N)
    W = torch.randn(3 * d, 3 * N, dtype=torch.float64) * 0.5
    gamma, beta = torch.rand(d, dtype=torch.float64) + 0.5, torch.randn(d, dtype=torch.float64) * 0.1

    def pre(xf):                      # smooth map R^{3N} -> (N, 3d)
        t = torch.tanh(W @ xf)
        return torch.stack([t * (i + 1) + torch.sin(xf).sum() * 0.1 for i in range(N)])

    def chain(xf):
        qkv = pre(xf)                                            # (N,3d)
        ln = torch.nn.functional.layer_norm(qkv[:, :d], (d,), gamma, beta, O.LN_EPS)
        q, k, v = qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:]
        hd = d // H
        qh, kh, vh = (t.view(N, H, hd).transpose(0, 1) for t in (q, k, v))
        att = torch.softmax(qh @ kh.transpose(-1, -2) / hd ** 0.5, -1) @ vh
        att = att.transpose(0, 1).reshape(N, d)
        return torch.cat([ln, att, torch.nn.functional.gelu(q, approximate="tanh")], -1)   # (N,3d)

    x0 = torch.randn(3 * N, dtype=torch.float64)
    J = torch.autograd.functional.jacobian(pre, x0)              # (N,3d,3N)
    lap_pre = torch.stack([torch.stack([torch.autograd.functional.hessian(lambda z: pre(z)[i, e], x0).trace()
                                        for e in range(3 * d)]) for i in range(N)])
    P = torch.zeros(1, N, C, 3 * d, dtype=torch.float64)
    P[0, :, 0] = pre(x0)
    P[0, :, 1:-1] = J.permute(0, 2, 1)
    P[0, :, -1] = lap_pre
    got = torch.cat([FL.layernorm_payload(P[..., :d], gamma, beta), FL.attention_payload(P, H),
                     FL.gelu_payload(P[..., :d])], -1)[0]
    Jc = torch.autograd.functional.jacobian(chain, x0)
    lap_c = torch.stack([torch.stack([torch.autograd.functional.hessian(lambda z: chain(z)[i, e], x0).trace()
                                      for e in range(3 * d)]) for i in range(N)])
    assert torch.allclose(got[:, 0], chain(x0), atol=1e-12)
    assert torch.allclose(got[:, 1:-1], Jc.permute(0, 2, 1), atol=1e-11)
    assert torch.allclose(got[:, -1], lap_c, atol=1e-10)


@pytest.mark.parametrize("B,N,d,H", [(2, 4, 32, 2), (3, 5, 64, 1), (2, 7, 48, 3), (1, 2, 16, 2), (2, 10, 64, 1), (1, 14, 128, 2)])
def test_first_layer_attention_closed_forms(B, N, d, H):
    """In front of the first attention token i depends on x_i only.  The closed forms that attention_first_layer.cuh
    evaluates on the compact payload (value, own tangents, Laplacian per token) must be the dense attention rule applied
    to the zero-expanded payload."""
    g = torch.Generator().manual_seed(7 * N + d)
    q5 = torch.randn(B, N, 5, 3 * d, generator=g, dtype=torch.float64) * 0.7
    dense = torch.zeros(B, N, 3 * N + 2, 3 * d, dtype=torch.float64)
    dense[:, :, 0] = q5[:, :, 0]
    dense[:, :, -1] = q5[:, :, 4]
    for i in range(N):
        dense[:, i, 1 + 3 * i:4 + 3 * i] = q5[:, i, 1:4]
    ref = FL.attention_payload(dense, H)
    out = FL.attention_first_layer_payload(q5, H)
    assert (out - ref).abs().max().item() < 1e-12 * ref.abs().max().item()
