"""Evaluation callers of the hot path: the two other consumers of ``Hamiltonian.local_energy`` in
the reference, ``utils/EVAL.py:17-61`` (energy of a checkpoint) and ``utils/helium_landscape.py:140-176``
(|psi|^2 and E_L on a grid), re-pointed at the fused kernels.  Plotting stays out of scope."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import sharding
from .config import Model_Config, Train_Config
from .hamiltonian import Hamiltonian
from .mcmc import MH
from .psiformer import PsiFormer, get_device


def load_checkpoint(path: str, model_config: Model_Config, device: Optional[torch.device] = None) -> PsiFormer:
    """Accepts both formats the reference writes/reads (EVAL.py:23-31): ``{"model_state_dict", "step"}``
    (train.py:106-112) or a raw state_dict."""
    loaded = torch.load(path, map_location="cpu")
    state = loaded["model_state_dict"] if isinstance(loaded, dict) and "model_state_dict" in loaded else loaded
    model = PsiFormer(model_config)
    model.load_state_dict(state)
    return model.eval().to(device or get_device())


@torch.no_grad()
def compute_energy(model: PsiFormer, monte_carlo: int = 100, burn_in: int = 10, step_size: float = 1.0,
                   batch_size: int = 1024, mh_steps_per_sample: int = 32, seed: Optional[int] = None
                   ) -> Tuple[float, float]:
    """Variational energy of ``model`` (EVAL.py:37-61): ``monte_carlo`` stored samples per chain from ``batch_size``
    chains.  Returns (mean, standard error); sums are accumulated on the device by the local-energy kernel and
    all-reduced over ranks when torch.distributed is initialised."""
    dev = next(model.parameters()).device
    cfg = Train_Config(batch_size=batch_size, monte_carlo_length=monte_carlo, burn_in_steps=burn_in, step_size=step_size,
                       mh_steps_per_sample=mh_steps_per_sample, seed=seed)
    shard = sharding.current_shard(1)
    mh = MH(model, cfg, model.config.n_electron_num, device=dev, walker_id0=shard.rank * batch_size)
    eng = model.ready_engine(dev)
    acc = torch.zeros(3, dtype=torch.float64, device=dev)
    mh._run_steps(mh._init_state(), burn_in)
    for _ in range(monte_carlo):
        mh._run_steps(mh._state, mh_steps_per_sample)
        eng.local_energy(mh._state, accum=acc)
    sharding.allreduce_energy_stats(acc)
    mean, var = sharding.energy_mean_and_variance(acc)
    n = max(1.0, float(acc[2].item()))
    return float(mean.item()), float((var / n).sqrt().item())


@torch.no_grad()
def landscape(model: PsiFormer, xmin: float, xmax: float, n: int, batch_size: int = 8192
              ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Two-electron landscape of helium_landscape.py:140-176: electrons on the x axis at (x1,0,0), (x2,0,0).
    Returns (xs (n,), |psi|^2 (n,n), E_L (n,n)) on the CPU; non-finite local energies are kept as such."""
    if model.config.n_electron_num != 2:
        raise ValueError("the landscape grid is defined for two electrons")
    dev = next(model.parameters()).device
    xs = torch.linspace(xmin, xmax, n, dtype=torch.float32)
    x1, x2 = torch.meshgrid(xs, xs, indexing="xy")
    R = torch.zeros(n * n, 2, 3)
    R[:, 0, 0], R[:, 1, 0] = x1.reshape(-1), x2.reshape(-1)
    ham = Hamiltonian(model, n_elec=2, Z=model.config.nuclear_charge)
    dens, ener = [], []
    for chunk in R.split(batch_size):
        out = ham._run(chunk.to(dev))
        dens.append(torch.exp(2.0 * out["logabs"]).cpu())
        ener.append(out["e_loc"].cpu())
    return xs, torch.cat(dens).reshape(n, n), torch.cat(ener).reshape(n, n)
