"""``MH``: the drop-in for mcmc.py:7-84 (batched all-electron Metropolis-Hastings).

Differences from the reference, none visible in the samples' distribution:
  * log|psi(current)| is cached between steps instead of recomputed (mcmc.py:40 evaluates it twice);
  * proposals / uniforms come from a device Philox4x32-10 stream keyed by (seed, global walker id,
    step), so chains are identical however the walkers are sharded over GPUs;
  * ``target`` must be backed by a ``PsiFormer`` (TypeError otherwise; single backend).

Additions: ``sample_energies`` runs the sampler loop and evaluates the local energy of
every stored sample on the resident chain state (one fused library call per sample, nothing but the
tiny per-sample results leaves the device), ``acceptance_rate`` / ``window_acceptance`` and
``adapt_step_size`` expose the step-size control loop the reference leaves to the user.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional

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
        self._window = (0, 0)                       # (accepted, proposed) at the last window_acceptance() call
        # the `mh_steps_per_sample` loop is ~35 stream-ordered launches per step with no host dependency: it is
        # captured once per step count into a CUDA graph and replayed (the Philox step counter lives on the device).
        # Captured graphs hold raw addresses, so the sampler OWNS its workspace (never the engine's shared one, which
        # a later, larger call may replace) and drops the graphs whenever that workspace or the chain tensors change.
        self.use_graph = True
        # look at the library's fp16-range event after every batch of steps (one stream synchronise) and repeat the batch
        # in tf32 mode if it was raised; False keeps _run_steps asynchronous (the event then only rejects the proposal)
        self.guard_range = True
        self._graphs: dict = {}
        self._counter: torch.Tensor | None = None
        self._param_key = None
        self._ws: torch.Tensor | None = None
        self._ws_B = -1

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

    # ---- chain bookkeeping shared by _run_steps and sample_energies ------------------------------------------------
    def _prepare(self, state: torch.Tensor):
        """Adopt ``state`` as the chain (private copy unless it already is), make sure the cache / counters / private
        workspace exist.  Returns (engine, state, fresh) with fresh = log|psi(current)| must be recomputed."""
        eng = self.model.ready_engine(self.device)
        fresh = state is not self._state or self._logabs is None
        if fresh:
            state = state.detach().to(self.device, torch.float32).contiguous().clone()
            self._logabs = torch.empty(state.shape[0], dtype=torch.float32, device=self.device)
            self._sign = torch.empty_like(self._logabs)
            self._graphs.clear()
        if self.n_accept is None:
            self.n_accept = torch.zeros(1, dtype=torch.int64, device=self.device)
        if self._counter is None:
            self._counter = torch.zeros(1, dtype=torch.int64, device=self.device)
        B = state.shape[0]
        if self._ws is None or self._ws_B != B:
            self._ws = eng.new_workspace(B, _lib.MODE_VALUE, _lib.MODE_ENERGY)
            self._ws_B = B
            self._graphs.clear()
        key = eng._param_key
        if key != self._param_key:          # parameters changed: the cached log|psi(current)| is stale
            self._param_key = key
            self._graphs.clear()
            fresh = True
        return eng, state, fresh

    def _range_backup(self, eng, state: torch.Tensor):
        """What an exact repeat of the coming steps needs (see _run_steps); also drops a stale range event."""
        eng._range_event(clear_only=True)
        return (state.clone(), self._logabs.clone(), self._sign.clone(), self.n_accept.clone())

    def _range_restore(self, backup, state: torch.Tensor) -> None:
        state.copy_(backup[0])
        self._logabs.copy_(backup[1])
        self._sign.copy_(backup[2])
        self.n_accept.copy_(backup[3])

    def _run_steps(self, state: torch.Tensor, steps: int) -> torch.Tensor:
        """``max(1, steps)`` Metropolis steps (mcmc.py:51-54), in place on a private copy of ``state``."""
        eng, state, fresh = self._prepare(state)
        n = max(1, int(steps))
        backup = self._range_backup(eng, state) if self.guard_range else None
        if fresh or not self.use_graph:
            self._counter.fill_(self._step)
            eng.mh_steps(state, self._logabs, self._sign, n, float(self.config.step_size), have_logabs=not fresh,
                         seed=self._ensure_seed(), walker_id0=self.walker_id0, step_counter=self._counter,
                         n_accept=self.n_accept, ws=self._ws)
        else:
            key = (n, float(self.config.step_size))
            g = self._graphs.get(key)
            if g is None:
                self._counter.fill_(self._step)
                torch.cuda.synchronize(self.device)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    eng.mh_steps(state, self._logabs, self._sign, n, float(self.config.step_size), have_logabs=True,
                                 seed=self._ensure_seed(), walker_id0=self.walker_id0, step_counter=self._counter,
                                 n_accept=self.n_accept, ws=self._ws)
                self._graphs[key] = g
            g.replay()
        if backup is not None and eng._range_event():
            # An activation left fp16's range in one of the forwards: inside the loop that proposal was simply rejected
            # (log|psi| = NaN), which is not the reference's accept rule.  Rare enough (coordinates beyond ~1e5 bohr) to
            # afford the exact answer: restore the chains and repeat the steps with tf32-split GEMMs.
            self._range_restore(backup, state)
            self._counter.fill_(self._step)
            eng.set_gemm_mode(_lib.GEMM_TF32_SPLIT)
            try:
                eng.mh_steps(state, self._logabs, self._sign, n, float(self.config.step_size), have_logabs=False,
                             seed=self._ensure_seed(), walker_id0=self.walker_id0, step_counter=self._counter,
                             n_accept=self.n_accept, ws=self._ws)
            finally:
                eng.set_gemm_mode(_lib.GEMM_FP16_SPLIT)
        self._step += n
        self.n_proposed += n * state.shape[0]
        self._state = state
        return state

    def state_dict(self) -> dict:
        """Everything needed to continue the chains exactly: positions, Philox seed and step counter."""
        return {"state": None if self._state is None else self._state.detach().clone().cpu(), "step": self._step,
                "seed": self._ensure_seed(), "walker_id0": self.walker_id0, "step_size": float(self.config.step_size)}

    def load_state_dict(self, sd: dict) -> None:
        self._seed, self._step, self.walker_id0 = int(sd["seed"]), int(sd["step"]), int(sd["walker_id0"])
        if "step_size" in sd:
            self.config.step_size = float(sd["step_size"])
        self._graphs.clear()
        self._logabs = None                   # recomputed on the next call (same value: log|psi| is deterministic)
        self._state = None if sd["state"] is None else sd["state"].to(self.device, torch.float32).contiguous().clone()

    # ---- acceptance statistics and step-size control ------------------------------------------------------------------
    @property
    def acceptance_rate(self) -> float:
        """Accepted / proposed moves since construction (one device->host read)."""
        if not self.n_proposed:
            return float("nan")
        return float(self.n_accept.item()) / self.n_proposed

    def window_acceptance(self) -> float:
        """Acceptance rate since the previous call of this method (NaN if nothing was proposed in between)."""
        acc = int(self.n_accept.item()) if self.n_accept is not None else 0
        a0, p0 = self._window
        self._window = (acc, self.n_proposed)
        return float("nan") if self.n_proposed == p0 else (acc - a0) / (self.n_proposed - p0)

    def adapt_step_size(self, target: float = 0.5, tolerance: float = 0.05, factor: float = 1.1,
                        bounds: tuple = (1e-3, 10.0)) -> float:
        """Step-size adaptation hook (not in the reference, whose step size is a constant of the preset): widen the
        Gaussian proposal when the windowed acceptance rate is above ``target + tolerance``, narrow it when below
        ``target - tolerance``.  Returns the step size now in effect; the sampler picks it up on its next call."""
        rate = self.window_acceptance()
        if rate == rate:
            if rate > target + tolerance:
                self.config.step_size = min(bounds[1], float(self.config.step_size) * factor)
            elif rate < target - tolerance:
                self.config.step_size = max(bounds[0], float(self.config.step_size) / factor)
        return float(self.config.step_size)

    # ---- the reference's public call -------------------------------------------------------------------------------------
    def _burn_in(self) -> None:
        if self._state is None:
            self._run_steps(self._init_state(), self.config.burn_in_steps)

    @torch.inference_mode()
    def sampler(self) -> torch.Tensor:
        """(monte_carlo_length, batch_size, n_elec, dim) samples; the chain persists across calls and is
        burnt in only on the first one (mcmc.py:56-84)."""
        self._burn_in()
        B, n_e, dim = self.config.batch_size, self.n_elec, self.config.dim
        samples_eq = torch.empty(self.config.monte_carlo_length, B, n_e, dim, device=self.device)
        for i in range(self.config.monte_carlo_length):
            self._run_steps(self._state, self.config.mh_steps_per_sample)
            samples_eq[i] = self._state
        return samples_eq

    @torch.inference_mode()
    def sample_energies(self, accum: Optional[torch.Tensor] = None, keep_samples: bool = True) -> Dict[str, torch.Tensor]:
        """The fused form of ``sampler()`` followed by ``Trainer._batched_energy_eval`` (mcmc.py:56-84 +
        train.py:60-101): for each of the ``monte_carlo_length`` stored samples, ``mh_steps_per_sample`` Metropolis
        steps and ONE local-energy pass on the resident chain state, in one library call (psif_sample_energy).  Two model
        evaluations per sample where the reference spends 2 * steps + 3; no host synchronisation inside the loop.

        Returns ``e_loc`` / ``logabs`` / ``status`` of shape (mc_len, B) and, if ``keep_samples`` (the parameter
        backward needs them), ``samples`` (mc_len, B, n_elec, 3).  ``accum`` (fp64[3] on the device) receives
        {sum E_L, sum E_L^2, n} over walkers with status 0: the operand of the energy all-reduce."""
        self._burn_in()
        cfg = self.config
        eng, state, fresh = self._prepare(self._state)
        n = max(1, int(cfg.mh_steps_per_sample))
        M, B = cfg.monte_carlo_length, state.shape[0]
        out = {"e_loc": torch.empty(M, B, dtype=torch.float32, device=self.device),
               "logabs": torch.empty(M, B, dtype=torch.float32, device=self.device),
               "status": torch.empty(M, B, dtype=torch.int32, device=self.device)}
        if keep_samples:
            out["samples"] = torch.empty(M, B, self.n_elec, 3, dtype=torch.float32, device=self.device)
        backup = self._range_backup(eng, state) if self.guard_range else None
        acc0 = accum.clone() if (backup is not None and accum is not None) else None

        def run(have_first: bool) -> None:
            self._counter.fill_(self._step)
            have = have_first
            for i in range(M):
                r = eng.sample_energy(state, self._logabs, self._sign, n, float(cfg.step_size), have_logabs=have,
                                      seed=self._ensure_seed(), walker_id0=self.walker_id0, step_counter=self._counter,
                                      n_accept=self.n_accept, accum=accum, ws=self._ws)
                have = True
                out["e_loc"][i], out["logabs"][i], out["status"][i] = r["e_loc"], r["logabs"], r["status"]
                if keep_samples:
                    out["samples"][i] = state

        run(not fresh)
        if backup is not None and eng._range_event():
            # an activation left fp16's range in a Metropolis forward or an energy pass: repeat the whole call exactly, with
            # tf32-split GEMMs, from the chains as they were (one synchronise per call; the caller reads the results next)
            self._range_restore(backup, state)
            if acc0 is not None:
                accum.copy_(acc0)
            eng.set_gemm_mode(_lib.GEMM_TF32_SPLIT)
            try:
                run(False)
            finally:
                eng.set_gemm_mode(_lib.GEMM_FP16_SPLIT)
        self._step += n * M
        self.n_proposed += n * B * M
        self._state = state
        return out
