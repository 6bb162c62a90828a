"""Pin the CPU oracle to vectors produced by the unmodified reference
(tests/golden/*.npz, written by oracle/make_golden.py)."""
import hashlib

import numpy as np
import pytest
import torch

from conftest import PINNED_CASES, ALL_CASES, GOLDEN_DIR
from oracle import psiformer_oracle as O


def _sha(p):
    h = hashlib.sha256()
    for k in p:
        h.update(k.encode())
        h.update(p[k].numpy().tobytes())
    return h.hexdigest()


@pytest.mark.parametrize("name", ALL_CASES)
def test_synthetic_params_are_reproducible(golden, name):
    sysm, params, data = golden(name)
    assert _sha(params) == str(data["params_sha256"])
    assert list(params.keys()) == list(O.param_shapes(sysm).keys())


@pytest.mark.parametrize("name", PINNED_CASES)
def test_value_and_sign_match_reference_fp64(golden, name):
    sysm, params, data = golden(name)
    p = O.cast_params(params, torch.float64)
    with torch.no_grad():
        la, sg = O.log_psi_and_sign(sysm, p, data["x"].double())
    assert torch.allclose(la, data["ref64_logabs"], rtol=1e-10, atol=1e-10)
    assert torch.equal(sg, data["ref64_sign"])


@pytest.mark.parametrize("name", ["debug", "he_small", "be"])
def test_local_energy_matches_reference_fp64(golden, name):
    sysm, params, data = golden(name)
    parts = O.local_energy_parts(sysm, O.cast_params(params, torch.float64), data["x"].double())
    for k, ref in (("grad", "ref64_grad"), ("lap", "ref64_lap"), ("pot", "ref64_pot"), ("e_loc", "ref64_eloc")):
        scale = max(1.0, data[ref].abs().max().item())
        assert (parts[k] - data[ref]).abs().max().item() <= 1e-9 * scale, k


@pytest.mark.parametrize("name", ["debug", "he_small", "be"])
def test_fp32_oracle_tracks_reference_fp32(golden, name):
    """Same ops, same order of magnitude of rounding: fp32 oracle vs fp32 reference."""
    sysm, params, data = golden(name)
    parts = O.local_energy_parts(sysm, params, data["x"])
    assert torch.allclose(parts["logabs"], data["ref32_logabs"], rtol=2e-5, atol=2e-5)
    assert torch.allclose(parts["pot"], data["ref32_pot"], rtol=1e-6, atol=1e-5)
    err = (parts["e_loc"].double() - data["ref64_eloc"]).abs().median().item()
    assert err < 1e-4


@pytest.mark.parametrize("name", PINNED_CASES)
def test_mh_decisions_match_reference(golden, name):
    """mcmc.py:31-49 replayed with the stored proposals/uniforms (fp32, bit-exact masks)."""
    sysm, params, data = golden(name)
    with torch.no_grad():
        state, acc = O.mh_run(lambda t: O.log_psi(sysm, params, t), data["x"], float(data["step_size"]),
                              data["mh_eps"], data["mh_u"])
    assert torch.equal(acc, data["mh_accept"])
    assert torch.equal(state, data["mh_final"])


def test_logdet_known_answers():
    z = np.load(f"{GOLDEN_DIR}/logdet_kat.npz")
    la, sg = O.logdet_matmul_value(torch.from_numpy(z["x1"]), torch.from_numpy(z["x2"]), torch.from_numpy(z["w"]))
    assert torch.allclose(la, torch.from_numpy(z["logabs"]), rtol=1e-12, atol=1e-12)
    assert torch.equal(sg, torch.from_numpy(z["sign"]))
    la2, sg2 = O.logdet_matmul_value(torch.from_numpy(z["near"]), torch.ones(1, 1, 1, 1, dtype=torch.float64),
                                     torch.ones(1, 1, dtype=torch.float64))
    assert torch.allclose(la2, torch.from_numpy(z["near_logabs"]), rtol=1e-12, atol=1e-12)
    assert torch.equal(sg2, torch.from_numpy(z["near_sign"]))


def test_near_singular_derivatives_are_finite():
    """logdet_matmul_stability_test.py:6-25 restated for the oracle."""
    base = torch.tensor([[1.0, 2.0], [2.0001, 4.0]], requires_grad=True)
    x1 = base[None, None]
    x2 = torch.tensor([[[[1.0]]]], requires_grad=True)
    out, _ = O.logdet_matmul_value(x1, x2, torch.ones(1, 1))
    assert torch.isfinite(out).all()
    g = torch.autograd.grad(out.sum(), x1, create_graph=True)[0]
    assert torch.isfinite(g).all()
    assert torch.isfinite(torch.autograd.grad(g.sum(), x1)[0]).all()


def test_hydrogen_1s_local_energy_is_minus_half():
    """local_energy_test.py:5-11 in the batched signature: E_L = -0.5 Ha for exact 1s."""
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1024, 1, 3, generator=g, dtype=torch.float64)
    sysm = O.OracleSystem(1, 1, 4, 1, 1, 0, ((1.0, (0.0, 0.0, 0.0)),))
    y, gr, lap = O.grad_and_laplacian(lambda t: -torch.linalg.norm(t[:, 0], dim=-1), x)
    e = -0.5 * (lap + gr.pow(2).sum(dim=(1, 2))) + O.potential(sysm, x)
    assert abs(e.mean().item() + 0.5) < 1e-3
    assert (e + 0.5).abs().median().item() < 5e-5


def test_single_nucleus_extension_reduces_to_reference_form():
    """App. A.7 with one nucleus at the origin must be the reference's features / potential."""
    sysm = O.SYSTEMS["Be"]
    x = O.synthetic_walkers(sysm, 8, 1).double()
    feats, r_ae = O.electron_features(sysm, x)
    r = torch.linalg.norm(x, dim=-1, keepdim=True)
    assert torch.equal(feats, torch.cat([x, r], -1))
    eps = 1e-5
    v = -4 * (1 / (torch.sqrt(x.pow(2).sum(-1) + eps) + eps)).sum(-1)
    i, j = torch.triu_indices(4, 4, offset=1)
    v = v + (1 / (torch.sqrt((x[:, i] - x[:, j]).pow(2).sum(-1) + eps) + eps)).sum(-1)
    assert torch.allclose(O.potential(sysm, x), v, rtol=0, atol=1e-12)
