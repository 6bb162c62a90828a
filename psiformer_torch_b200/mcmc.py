"""``MH``: the drop-in for mcmc.py:7-84 (batched all-electron Metropolis-Hastings).

Differences from the reference, none visible in the samples' distribution:
  * log|psi(current)| is cached between steps instead of recomputed (mcmc.py:40 evaluates it twice);
  * proposals / uniforms come from a device Philox4x32-10 stream keyed by (seed, global walker id,
    step), so chains are identical however the walkers are sharded over GPUs;
  * ``target`` must be backed by a ``PsiFormer`` (TypeError otherwise; single backend).
"""
from __future__ import annotations

from typing import Callable, Optional

import torch

from . import _lib
from .config import Train_Config
from .hamiltonian import _resolve_model
from .psiformer import get_device

_INIT_STEP = 2**64 - 1      # Philox step index reserved for the initial positions


class MH():
    def __init__(self, target: Callable[[torch.Tensor], torch.Tensor], config: Train_Config, n_elec: int,
                 device: torch.device | None = None, walker_id0: int = 0):
        self.target = target
        self.config = config
        self.n_elec = n_elec
        self.device = torch.device(device) if device is not None else get_device()
        if self.device.type == "cuda" and self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.model = _resolve_model(target)
        self.walker_id0 = int(walker_id0)          # global id of this rank's first walker
        self._state: torch.Tensor | None = None
        self._logabs: torch.Tensor | None = None
        self._sign: torch.Tensor | None = None
        self._step = 0                              # Philox step counter (host mirror)
        self._seed: Optional[int] = getattr(config, "seed", None)
        self.n_accept = None
        self.n_proposed = 0
        # the `mh_steps_per_sample` loop is ~35 stream-ordered launches per step with no host dependency: it is
        # captured once per step count into a CUDA graph and replayed (the Philox step counter lives on the device)
        self.use_graph = True
        self._graphs: dict = {}
        self._counter: torch.Tensor | None = None
        self._param_key = None

    def _init_state(self) -> torch.Tensor:
        """N(0,1) start (mcmc.py:58-60), drawn from the same Philox stream at a reserved step index so that the
        chains do not depend on how walkers are sharded."""
        B, n_e, dim = self.config.batch_size, self.n_elec, self.config.dim
        if dim != 3:
            return torch.randn(B, n_e, dim, device=self.device)
        out = torch.empty(B, n_e, 3, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().psif_philox_normal(self._ensure_seed(), self.walker_id0, _INIT_STEP, B, n_e,
                                                      _lib.ptr(out), None,
                                                      torch.cuda.current_stream(self.device).cuda_stream))
        return out

    def _ensure_seed(self) -> int:
        if self._seed is None:
            self._seed = int(torch.randint(0, 2**62, (1,)).item())
        return self._seed

    def _run_steps(self, state: torch.Tensor, steps: int) -> torch.Tensor:
        """``max(1, steps)`` Metropolis steps (mcmc.py:51-54), in place on a private copy of ``state``."""
        eng = self.model.ready_engine(self.device)
        fresh = state is not self._state or self._logabs is None
        if fresh:
            state = state.detach().to(self.device, torch.float32).contiguous().clone()
            self._logabs = torch.empty(state.shape[0], dtype=torch.float32, device=self.device)
            self._sign = torch.empty_like(self._logabs)
        if self.n_accept is None:
            self.n_accept = torch.zeros(1, dtype=torch.int64, device=self.device)
        n = max(1, int(steps))
        if self._counter is None:
            self._counter = torch.zeros(1, dtype=torch.int64, device=self.device)
        key = eng._param_key
        if key != self._param_key:          # parameters changed: the cached log|psi(current)| is stale, and graphs
            self._param_key = key           # are re-captured to be safe (weight tensor maps are baked in)
            self._graphs.clear()
            fresh = True
        if fresh or not self.use_graph:
            self._graphs.clear()
            self._counter.fill_(self._step)
            eng.mh_steps(state, self._logabs, self._sign, n, float(self.config.step_size), have_logabs=not fresh,
                         seed=self._ensure_seed(), walker_id0=self.walker_id0, step_counter=self._counter,
                         n_accept=self.n_accept)
        else:
            g = self._graphs.get(n)
            if g is None:
                self._counter.fill_(self._step)
                torch.cuda.synchronize(self.device)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    eng.mh_steps(state, self._logabs, self._sign, n, float(self.config.step_size), have_logabs=True,
                                 seed=self._ensure_seed(), walker_id0=self.walker_id0, step_counter=self._counter,
                                 n_accept=self.n_accept)
                self._graphs[n] = g
            g.replay()
        self._step += n
        self.n_proposed += n * state.shape[0]
        self._state = state
        return state

    def state_dict(self) -> dict:
        """Everything needed to continue the chains exactly: positions, Philox seed and step counter."""
        return {"state": None if self._state is None else self._state.detach().clone().cpu(), "step": self._step,
                "seed": self._ensure_seed(), "walker_id0": self.walker_id0}

    def load_state_dict(self, sd: dict) -> None:
        self._seed, self._step, self.walker_id0 = int(sd["seed"]), int(sd["step"]), int(sd["walker_id0"])
        self._graphs.clear()
        self._logabs = None                   # recomputed on the next call (same value: log|psi| is deterministic)
        self._state = None if sd["state"] is None else sd["state"].to(self.device, torch.float32).contiguous().clone()

    @property
    def acceptance_rate(self) -> float:
        if not self.n_proposed:
            return float("nan")
        return float(self.n_accept.item()) / self.n_proposed

    @torch.inference_mode()
    def sampler(self) -> torch.Tensor:
        """(monte_carlo_length, batch_size, n_elec, dim) samples; the chain persists across calls and is
        burnt in only on the first one (mcmc.py:56-84)."""
        if self._state is None:
            self._state = None
            st = self._init_state()
            self._run_steps(st, self.config.burn_in_steps)
        B, n_e, dim = self.config.batch_size, self.n_elec, self.config.dim
        samples_eq = torch.empty(self.config.monte_carlo_length, B, n_e, dim, device=self.device)
        for i in range(self.config.monte_carlo_length):
            self._run_steps(self._state, self.config.mh_steps_per_sample)
            samples_eq[i] = self._state
        return samples_eq
