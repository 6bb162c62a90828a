"""The masking rules of the training loop's energy evaluation (train.py:60-101) with non-finite inputs, on the
reference-shaped ``_batched_energy_eval`` and on the fused ``_fused_energy_eval``, and the sampler's
private workspace (a captured Metropolis graph must survive larger calls on the same engine)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _trainer(batch=8, mc_len=1, ebs=4, preset="small"):
    from psiformer_torch_b200.psiformer import PsiFormer
    from psiformer_torch_b200.train import Trainer, wrapper
    torch.manual_seed(0)
    mcfg, tcfg = wrapper(preset, wand_mode="disabled")
    tcfg.batch_size, tcfg.monte_carlo_length, tcfg.mh_steps_per_sample, tcfg.burn_in_steps = batch, mc_len, 2, 4
    tcfg.energy_batch_size, tcfg.seed, tcfg.train_steps = ebs, 7, 4
    return Trainer(PsiFormer(mcfg), tcfg, False)


def test_batched_energy_eval_skips_and_masks_like_the_reference():
    tr = _trainer()
    g = torch.Generator().manual_seed(1)
    samples = torch.randn(1, 8, 2, 3, generator=g).cuda()
    clean_lp, clean_e = tr._batched_energy_eval(samples)
    assert clean_lp.shape == (8,) and clean_e.shape == (8,) and torch.isfinite(clean_e).all()
    bad = samples.clone()
    bad[0, 1, 0, 0] = float("nan")          # chunk 0 (walkers 0-3): log|psi| non-finite -> the whole chunk is skipped
    bad[0, 6, 1] = 0.0                       # chunk 1 (walkers 4-7): an electron ON the nucleus -> E_L non-finite -> dropped
    lp, e = tr._batched_energy_eval(bad)
    assert lp.shape == (3,) and e.shape == (3,) and torch.isfinite(e).all() and torch.isfinite(lp).all()
    keep = [4, 5, 7]
    assert torch.allclose(e, clean_e[keep]) and torch.allclose(lp, clean_lp[keep])
    allbad = torch.full_like(samples, float("nan"))
    assert tr._batched_energy_eval(allbad) == (None, None)


def test_fused_energy_eval_applies_the_same_rules(monkeypatch):
    from psiformer_torch_b200 import _lib
    tr = _trainer(batch=8, mc_len=1, ebs=4)
    res = tr.mh.sample_energies()
    e0, la0, x0 = res["e_loc"].clone(), res["logabs"].clone(), res["samples"].clone()
    assert res["e_loc"].shape == (1, 8) and torch.isfinite(e0).all() and (res["status"] == 0).all()

    def fake(accum=None, keep_samples=True):
        e, la, st = e0.clone(), la0.clone(), torch.zeros_like(res["status"])
        la[0, 1] = float("nan"); st[0, 1] = _lib.ST_NONFINITE_LOGDET        # chunk 0 skipped
        e[0, 6] = float("inf"); st[0, 6] = _lib.ST_NONFINITE_ELOC           # entry dropped
        return {"e_loc": e, "logabs": la, "status": st, "samples": x0.clone()}
    monkeypatch.setattr(tr.mh, "sample_energies", fake)
    lp, e, acc = tr._fused_energy_eval()
    keep = [4, 5, 7]
    assert torch.equal(e, e0[0, keep]) and torch.equal(lp.detach(), la0[0, keep])
    assert acc[2].item() == 3 and abs(acc[0].item() - e0[0, keep].double().sum().item()) < 1e-9
    lp.sum().backward()                        # the cached log|psi| carries the parameter backward
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in tr.model.parameters())


def test_fused_path_matches_sampler_plus_batched_eval():
    """sample_energies == sampler() followed by the energy evaluation of the same samples (same Philox stream)."""
    a, b = _trainer(batch=64, mc_len=3, ebs=64), _trainer(batch=64, mc_len=3, ebs=64)
    b.model.load_state_dict(a.model.state_dict())
    res = a.mh.sample_energies()
    samples = b.mh.sampler()
    assert torch.equal(res["samples"], samples)
    lp, e = b._batched_energy_eval(samples)
    assert torch.allclose(res["e_loc"].reshape(-1), e, rtol=1e-5, atol=1e-5)
    assert torch.allclose(res["logabs"].reshape(-1), lp.detach(), rtol=1e-5, atol=1e-5)
    r = a.mh.acceptance_rate
    assert 0.0 < r <= 1.0 and a.mh.window_acceptance() == pytest.approx(r)
    s0 = a.mh.config.step_size
    s1 = a.mh.adapt_step_size(target=0.0 if r > 0.5 else 1.0, tolerance=0.0)    # forces a change in a known direction
    assert (s1 > s0) == (r > 0.5) or r != r


def test_metropolis_graph_survives_larger_calls_on_the_same_engine():
    """ADVICE (round 1): the captured MH graph used to hold the address of the engine's shared workspace, which a later
    larger call replaced.  sampler() -> bigger model(x) / local_energy(x) -> sampler() must keep producing the same chain
    as an undisturbed sampler."""
    from psiformer_torch_b200.hamiltonian import Hamiltonian
    a, b = _trainer(batch=32, mc_len=2, ebs=32), _trainer(batch=32, mc_len=2, ebs=32)
    b.model.load_state_dict(a.model.state_dict())
    s_a1, s_b1 = a.mh.sampler(), b.mh.sampler()
    assert torch.equal(s_a1, s_b1)
    big = torch.randn(4096, 2, 3, device="cuda")
    with torch.no_grad():
        a.model(big)                                       # grows / replaces the engine's shared VALUE workspace
    Hamiltonian(a.model, n_elec=2, Z=2).local_energy(big)  # and the ENERGY one
    junk = torch.empty(64 * 1024 * 1024, device="cuda").normal_()   # recycle freed blocks
    s_a2, s_b2 = a.mh.sampler(), b.mh.sampler()
    del junk
    assert torch.equal(s_a2, s_b2) and torch.isfinite(s_a2).all()
