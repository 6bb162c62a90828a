"""Parameter gradients of log|psi| (psif_logpsi_backward) against torch.autograd through the CPU
oracle, and the reference train step (train.py:128-152) end to end on the CUDA path."""
import pytest
import torch

from oracle import psiformer_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["debug", "he_small", "large", "be", "lih"])
def test_parameter_gradients_match_autograd(golden, name):
    from gpu_util import make_engine
    sysm, params, data = golden(name)
    x = data["x"]
    g = torch.Generator().manual_seed(1)
    gbar = torch.randn(x.shape[0], generator=g, dtype=torch.float64)
    p64 = {k: v.double().requires_grad_(True) for k, v in params.items()}
    loss = (gbar * O.log_psi(sysm, p64, x.double())).sum()
    ref = torch.autograd.grad(loss, list(p64.values()), allow_unused=True)   # He has no same-spin pair: alpha_par unused
    ref = [torch.zeros_like(v) if r is None else r for r, v in zip(ref, p64.values())]
    eng = make_engine(sysm, params)
    flat = eng.logpsi_backward(x.cuda(), gbar.float().cuda()).double().cpu()
    o = 0
    worst = 0.0
    for (k, v), r in zip(params.items(), ref):
        got = flat[o:o + v.numel()].view(v.shape)
        o += v.numel()
        scale = r.abs().max().clamp_min(1e-6)
        err = ((got - r).abs().max() / scale).item()
        worst = max(worst, err)
        assert err < 5e-4, (k, err, scale.item())
    assert o == flat.numel()
    print(f"\n[{name}] worst relative gradient error {worst:.2e}")


def test_backward_through_module_and_chunking(golden):
    from psiformer_torch_b200.config import Model_Config
    from psiformer_torch_b200.psiformer import PsiFormer
    sysm, params, data = golden("large")
    cfg = Model_Config(n_layer=4, n_head=32, n_embd=256, n_determinants=4, n_electron_num=6, n_spin_up=4, n_spin_down=2,
                       nuclear_charge=6)
    model = PsiFormer(cfg)
    model.load_state_dict(params, strict=True)
    model = model.cuda()
    x = data["x"].cuda()
    wts = torch.linspace(-1, 1, x.shape[0], device="cuda")
    (wts * model(x)).sum().backward()
    g_all = torch.cat([p.grad.reshape(-1) for p in model.parameters()]).clone()
    model.zero_grad()
    (wts[:7] * model(x[:7])).sum().backward()
    (wts[7:] * model(x[7:])).sum().backward()       # gradients accumulate across calls like any autograd op
    g_two = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    assert torch.allclose(g_all, g_two, rtol=1e-4, atol=1e-5 * g_all.abs().max().item())
    assert torch.isfinite(g_all).all() and g_all.abs().max() > 0


def test_reference_train_loop_runs_on_the_cuda_path():
    """Trainer.train_step = sample -> energies -> score-function loss -> backward -> AdamW (train.py:128-152)."""
    from psiformer_torch_b200.psiformer import PsiFormer
    from psiformer_torch_b200.train import Trainer, wrapper
    torch.manual_seed(0)
    mcfg, tcfg = wrapper("small", wand_mode="disabled")
    tcfg.batch_size, tcfg.monte_carlo_length, tcfg.mh_steps_per_sample, tcfg.burn_in_steps = 256, 4, 8, 32
    tcfg.train_steps, tcfg.lr, tcfg.seed = 12, 2e-3, 3
    trainer = Trainer(PsiFormer(mcfg), tcfg, False)
    before = [p.detach().clone() for p in trainer.model.parameters()]
    trainer.train()
    hist = trainer.history
    assert len(hist) == 12 and all(torch.isfinite(torch.tensor(h["Energy"])) for h in hist)
    assert any(not torch.equal(a, b) for a, b in zip(before, trainer.model.parameters()))
    e_first = sum(h["Energy"] for h in hist[:3]) / 3
    e_last = sum(h["Energy"] for h in hist[-3:]) / 3
    print(f"\n[train] He small: E {e_first:.4f} -> {e_last:.4f} Ha, acceptance {hist[-1]['mh_acceptance']:.2f}")
    assert e_last < e_first + 0.3 and -4.5 < e_last < -1.0     # variational energy of He is -2.9037 Ha


def test_checkpoint_resume_is_exact(tmp_path):
    """chains + Philox counters + optimiser state are saved, so a resumed run repeats the original."""
    from psiformer_torch_b200.psiformer import PsiFormer
    from psiformer_torch_b200.train import Trainer, wrapper

    def make():
        torch.manual_seed(0)
        mcfg, tcfg = wrapper("small", wand_mode="disabled")
        tcfg.batch_size, tcfg.monte_carlo_length, tcfg.mh_steps_per_sample, tcfg.burn_in_steps = 128, 2, 4, 16
        tcfg.train_steps, tcfg.lr, tcfg.seed, tcfg.checkpoint_step = 6, 1e-3, 5, 1
        tcfg.checkpoint_dir, tcfg.checkpoint_name = str(tmp_path), "ck.pth"
        return Trainer(PsiFormer(mcfg), tcfg, False)

    a = make()
    for step in range(3):
        a.train_step(step)
    a.save_checkpoint(3)
    ref = [a.train_step(s)["Energy"].item() for s in (3, 4)]
    b = make()
    assert b.load_checkpoint() == 3
    got = [b.train_step(s)["Energy"].item() for s in (3, 4)]
    assert got == ref
    assert all(torch.equal(p, q) for p, q in zip(a.model.parameters(), b.model.parameters()))


def test_evaluation_callers(golden):
    """compute_energy (utils/EVAL.py:37-61) and the two-electron landscape (helium_landscape.py:140-176)."""
    from psiformer_torch_b200.config import PSIFORMER_TORCH_SMALL_MODEL
    from psiformer_torch_b200.evaluate import compute_energy, landscape
    from psiformer_torch_b200.psiformer import PsiFormer
    sysm, params, data = golden("he_small")
    model = PsiFormer(PSIFORMER_TORCH_SMALL_MODEL)
    model.load_state_dict(params, strict=True)
    model = model.cuda()
    e, se = compute_energy(model, monte_carlo=20, burn_in=50, step_size=0.6, batch_size=512, mh_steps_per_sample=8, seed=1)
    e2, _ = compute_energy(model, monte_carlo=20, burn_in=50, step_size=0.6, batch_size=512, mh_steps_per_sample=8, seed=1)
    # same chains for a fixed seed; the device-side double accumulator adds in arbitrary order
    assert abs(e - e2) < 1e-9 and se > 0 and abs(e) < 50
    xs, dens, ener = landscape(model, -1.5, 1.5, 16)
    assert dens.shape == (16, 16) and torch.isfinite(dens).all() and (dens >= 0).all()
    # spot check four grid points against the fp64 oracle
    pts = torch.zeros(4, 2, 3)
    idx = [(1, 3), (5, 12), (9, 2), (14, 7)]
    for k, (i2, i1) in enumerate(idx):
        pts[k, 0, 0], pts[k, 1, 0] = xs[i1], xs[i2]
    ref = O.local_energy_parts(sysm, O.cast_params(params, torch.float64), pts.double())
    got = torch.stack([ener[i2, i1] for i2, i1 in idx]).double()
    assert (got - ref["e_loc"]).abs().max() < 1e-3
