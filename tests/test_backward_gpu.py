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
