"""The drop-in claim, proven with the reference's OWN training loop: ``oracle/_ref/psiformer_torch/train.py`` (the
unmodified file, vendored by oracle/build_ref.py) is executed on top of ``install_as("psiformer_torch")``, so every
``from psiformer_torch.x import y`` in it resolves to this package.  Its ``Trainer.train()`` must run as is, and the
energies / loss its ``_batched_energy_eval`` produces on a fixed sample tensor must equal what the real reference
(run in a separate CPU process from the same vendored copy, same weights, same samples) produces."""
import importlib.util
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu

_CPU_SIDE = r"""
import json, os, sys, torch
sys.path.insert(0, sys.argv[1])                      # oracle/_ref: the real reference package
os.environ["WANDB_MODE"] = "disabled"
os.environ["CUDA_VISIBLE_DEVICES"] = ""              # the reference falls back to the CPU (psiformer.py:13-16)
from psiformer_torch.psiformer import PsiFormer
from psiformer_torch.train import Trainer, wrapper
blob = torch.load(sys.argv[2])
mcfg, tcfg = wrapper(blob["preset"], wand_mode="disabled")
tcfg.energy_batch_size = blob["energy_batch_size"]
model = PsiFormer(mcfg)
model.load_state_dict(blob["state_dict"], strict=True)
tr = Trainer(model, tcfg, False)
lp, e = tr._batched_energy_eval(blob["samples"])
E = e.mean().detach()
loss = 2 * ((e.detach() - E) * lp).mean()
loss.backward()
g = torch.cat([p.grad.reshape(-1) for p in model.parameters() if p.grad is not None])
print(json.dumps({"E_mean": E.item(), "loss": loss.item(), "n": e.numel(), "grad_norm": g.norm().item(),
                  "e_loc": e.tolist()}))
"""


def _ref_dir():
    from oracle import build_ref
    if not build_ref.available():
        pytest.skip("oracle/_ref has not been vendored (python oracle/build_ref.py in the build container)")
    assert build_ref.verify(), "oracle/_ref differs from its manifest"
    return build_ref.REF_DIR


@pytest.mark.parametrize("preset", ["debug", "small"])
def test_reference_train_loop_runs_unmodified_on_this_package(preset, tmp_path, monkeypatch):
    ref_dir = _ref_dir()
    import psiformer_torch_b200 as P
    saved = {k: v for k, v in sys.modules.items() if k == "psiformer_torch" or k.startswith("psiformer_torch.")}
    try:
        P.install_as("psiformer_torch")
        spec = importlib.util.spec_from_file_location("_reference_train_py", os.path.join(ref_dir, "psiformer_torch", "train.py"))
        ref_train = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref_train)                        # the reference's file, our modules underneath
        from psiformer_torch_b200.psiformer import PsiFormer
        assert ref_train.PsiFormer is PsiFormer and ref_train.MH.__module__.startswith("psiformer_torch_b200")
        torch.manual_seed(0)
        mcfg, tcfg = ref_train.wrapper(preset, wand_mode="disabled")
        tcfg.train_steps, tcfg.checkpoint_dir, tcfg.seed = 3, str(tmp_path), 5
        if preset == "small":
            tcfg.batch_size, tcfg.monte_carlo_length, tcfg.mh_steps_per_sample = 16, 4, 4
        trainer = ref_train.Trainer(PsiFormer(mcfg), tcfg, False)
        before = [p.detach().clone() for p in trainer.model.parameters()]
        trainer.train()                                            # train.py:115-194, unmodified
        assert any(not torch.equal(a, b) for a, b in zip(before, trainer.model.parameters()))
        assert all(torch.isfinite(p).all() for p in trainer.model.parameters())

        # same weights, same samples: this package (GPU) vs the real reference (CPU subprocess)
        samples = trainer.mh.sampler().clone()
        trainer.model.zero_grad()
        lp, e = trainer._batched_energy_eval(samples)
        E = e.mean().detach()
        loss = 2 * ((e.detach() - E) * lp).mean()
        loss.backward()
        gn = torch.cat([p.grad.reshape(-1) for p in trainer.model.parameters() if p.grad is not None]).norm().item()
        blob = {"preset": preset, "energy_batch_size": tcfg.energy_batch_size, "samples": samples.cpu(),
                "state_dict": {k: v.detach().cpu() for k, v in trainer.model.state_dict().items()}}
        torch.save(blob, tmp_path / "blob.pt")
        out = subprocess.run([sys.executable, "-c", _CPU_SIDE, ref_dir, str(tmp_path / "blob.pt")], capture_output=True,
                             text=True, timeout=900)
        assert out.returncode == 0, out.stderr[-2000:]
        ref = json.loads(out.stdout.strip().splitlines()[-1])
        assert ref["n"] == e.numel()
        e_ref = torch.tensor(ref["e_loc"], dtype=torch.float64)
        err = (e.detach().double().cpu() - e_ref).abs()
        print(f"\n[{preset}] E_mean {E.item():.6f} vs reference {ref['E_mean']:.6f}; per-sample |dE_L| max {err.max():.2e}; "
              f"loss {loss.item():.5f} vs {ref['loss']:.5f}; |grad| {gn:.4f} vs {ref['grad_norm']:.4f}")
        # both sides are fp32 evaluations of the same function: the reference's own fp32 noise is the yardstick
        assert err.median() < 1e-4 and abs(E.item() - ref["E_mean"]) < 2e-4 * max(1.0, abs(ref["E_mean"]))
        assert abs(loss.item() - ref["loss"]) < 2e-3 * max(1.0, abs(ref["loss"]))
        assert abs(gn - ref["grad_norm"]) < 2e-2 * max(1e-3, ref["grad_norm"])
    finally:
        for k in [k for k in sys.modules if k == "psiformer_torch" or k.startswith("psiformer_torch.")]:
            del sys.modules[k]
        sys.modules.update(saved)
