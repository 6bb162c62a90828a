"""``Trainer`` / ``wrapper``: the caller of the hot path, API-compatible with train.py:24-246.

The loop is the reference's (sample -> chunked energy evaluation -> score-function loss -> AdamW +
cosine schedule + clipping), re-expressed over the fused kernels, plus what a walker-sharded run
needs: the mean energy and the gradients are averaged over ranks when ``torch.distributed`` is
initialised.  Logging to wandb happens only if ``wand_mode`` is not "disabled".
"""
from __future__ import annotations

import logging
import time
from dataclasses import replace
from typing import Optional, Tuple

import torch
import torch.optim as optim

from . import _lib, sharding
from .config import DEBUG_CONF, LARGE_CONF, SMALL_CONF, Train_Config
from .hamiltonian import Hamiltonian
from .mcmc import MH
from .psiformer import PsiFormer, get_device

logger = logging.getLogger("psiformer_torch_b200.train")


class Trainer():
    def __init__(self, model: PsiFormer, config: Train_Config, push: bool):
        self.device = get_device()
        self.model = model.to(self.device)
        self.config = config
        self.push = push
        self.optimizer = optim.AdamW(self.model.parameters(), lr=config.lr, betas=(0.9, 0.95),
                                     weight_decay=1e-4, amsgrad=True)
        self.scheduler = optim.lr_scheduler.CosineAnnealingLR(self.optimizer, T_max=config.train_steps,
                                                              eta_min=config.lr * 0.1)
        rank = sharding.current_shard(1).rank   # every rank owns `batch_size` walkers (weak scaling)
        sharding.broadcast_parameters(self.model.parameters())       # replicas start from rank 0's initialisation
        self.mh = MH(self.log_psi, self.config, self.model.config.n_electron_num, device=self.device,
                     walker_id0=rank * config.batch_size)
        self.hamilton = Hamiltonian(self.log_psi, n_elec=self.model.config.n_electron_num,
                                    Z=self.model.config.nuclear_charge)
        self.history: list[dict] = []

    def log_psi(self, x: torch.Tensor) -> torch.Tensor:
        if x.device != self.device:
            x = x.to(self.device)
        return self.model(x)

    def _batched_energy_eval(self, samples: torch.Tensor) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
        """samples (mc_steps, B, n_elec, 3) -> (log|psi| (M,), E_L (M,)) with the reference's
        skip/mask rules (train.py:60-101): chunks whose log|psi| is non-finite are skipped, non-finite
        E_L entries are dropped."""
        flat = samples.reshape(-1, samples.size(-2), samples.size(-1))
        logpsis, local_es = [], []
        for chunk in flat.split(self.config.energy_batch_size):
            try:
                logpsi = self.log_psi(chunk)
            except ValueError as e:
                logger.warning(f"Skipping chunk due to log_psi error: {e}")
                continue
            if not torch.isfinite(logpsi).all():
                logger.warning("Skipping chunk with non-finite log_psi")
                continue
            local_energy = self.hamilton.local_energy(chunk)
            finite_mask = torch.isfinite(local_energy)
            if not finite_mask.all():
                logger.warning("Dropping non-finite local_energy entries")
                logpsi = logpsi[finite_mask]
                local_energy = local_energy[finite_mask]
            if logpsi.numel() == 0:
                continue
            logpsis.append(logpsi)
            local_es.append(local_energy)
        if not logpsis:
            return None, None
        return torch.cat(logpsis, dim=0), torch.cat(local_es, dim=0)

    def save_checkpoint(self, step):
        """train.py:103-113 ({"model_state_dict", "step"}) plus what an exact resume needs and the reference omits:
        optimiser / scheduler state and the sampler's chains with their Philox counters.  Under torch.distributed rank 0
        writes the model / optimiser file and every rank writes its own chains next to it (``<name>.mh<rank>``)."""
        if step % self.config.checkpoint_step == 0:
            path = self.config.init_checkpoint()
            shard = sharding.current_shard(1)
            if shard.rank == 0:
                torch.save({"model_state_dict": self.model.state_dict(), "step": step,
                            "optimizer_state_dict": self.optimizer.state_dict(),
                            "scheduler_state_dict": self.scheduler.state_dict(), "mh_state": self.mh.state_dict()}, path)
                print(f"Saved checkpoint at step {step}")
            else:
                torch.save({"mh_state": self.mh.state_dict(), "step": step}, f"{path}.mh{shard.rank}")

    def load_checkpoint(self, path: Optional[str] = None) -> int:
        """Resume from ``save_checkpoint`` output (or from a reference checkpoint: weights only).  Returns the step.
        Checkpoints hold tensors and plain containers only, so they are read with ``weights_only=True``."""
        path = path or self.config.init_checkpoint()
        ck = torch.load(path, map_location="cpu", weights_only=True)
        state = ck["model_state_dict"] if isinstance(ck, dict) and "model_state_dict" in ck else ck
        self.model.load_state_dict(state)
        if isinstance(ck, dict):
            if "optimizer_state_dict" in ck:
                self.optimizer.load_state_dict(ck["optimizer_state_dict"])
            if "scheduler_state_dict" in ck:
                self.scheduler.load_state_dict(ck["scheduler_state_dict"])
            rank = sharding.current_shard(1).rank
            mh_state = ck.get("mh_state")
            if rank > 0:
                import os
                mine = f"{path}.mh{rank}"
                mh_state = torch.load(mine, map_location="cpu", weights_only=True)["mh_state"] if os.path.exists(mine) else None
            if mh_state is not None:
                self.mh.load_state_dict(mh_state)
            return int(ck.get("step", 0))
        return 0

    def _fused_energy_eval(self) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor], torch.Tensor]:
        """``mh.sampler()`` + ``_batched_energy_eval`` (train.py:128-130) as ONE pass over the resident chains
        (``MH.sample_energies``): per stored sample, the Metropolis steps and a single local-energy evaluation whose
        log|psi| is the operand of the loss, so the extra forward of train.py:70 disappears.  Returns (log|psi| with the
        parameter backward attached, E_L, fp64[3] {sum E_L, sum E_L^2, n} of this rank) under the reference's masking
        rules (train.py:73-95)."""
        dev = self.device
        acc = torch.zeros(3, dtype=torch.float64, device=dev)
        res = self.mh.sample_energies(accum=acc)
        N = self.model.config.n_electron_num
        xs = res["samples"].reshape(-1, N, 3)
        e, la, st = res["e_loc"].reshape(-1), res["logabs"].reshape(-1), res["status"].reshape(-1)
        bits = int(torch.bitwise_or(st.amax(), st.amin()).item()) if st.numel() else 0     # the one host sync of the step
        if bits & _lib.ST_FP16_RANGE:
            # an activation left fp16's range somewhere: re-evaluate the samples through the guarded call (tf32 split)
            acc.zero_()
            eng = self.model.ready_engine(dev)
            parts = [eng.local_energy(c, accum=acc) for c in xs.split(self.config.energy_batch_size)]
            e = torch.cat([p["e_loc"] for p in parts])
            la = torch.cat([p["logabs"] for p in parts])
            st = torch.cat([p["status"] for p in parts])
            bits = int(torch.bitwise_or(st.amax(), st.amin()).item())
        if bits & (_lib.ST_NONFINITE_LOGDET | _lib.ST_NONFINITE_ELOC):
            # rare: apply the reference's rules literally.  A chunk of `energy_batch_size` consecutive samples holding a
            # non-finite log|psi| is skipped as a whole (train.py:73-83), non-finite E_L entries are dropped (:86-90)
            chunk = torch.arange(e.numel(), device=dev) // self.config.energy_batch_size
            bad_chunk = torch.zeros(int(chunk[-1].item()) + 1, dtype=torch.bool, device=dev)
            bad_chunk[chunk[~torch.isfinite(la)]] = True
            keep = torch.isfinite(e) & ~bad_chunk[chunk]
            if bad_chunk.any():
                logger.warning("Skipping chunk with non-finite log_psi")
            if not torch.isfinite(e).all():
                logger.warning("Dropping non-finite local_energy entries")
            xs, e, la = xs[keep], e[keep], la[keep]
            e64 = e.double()
            acc = torch.stack([e64.sum(), (e64 * e64).sum(), torch.tensor(float(e.numel()), dtype=torch.float64, device=dev)])
        if e.numel() == 0:
            return None, None, acc
        return self.model.log_psi_cached(xs, la), e.clone(), acc

    def train_step(self, step: int) -> Optional[dict]:
        t0 = time.perf_counter()
        log_psi_vals, local_energies, acc = self._fused_energy_eval()
        # the energy statistics of all ranks in one all-reduce (3 doubles, accumulated on the device by the kernel);
        # the sample count rides along, so "no valid samples anywhere" is decided collectively and no rank is left
        # waiting in a collective (a rank without samples still takes part with zeros)
        sharding.allreduce_energy_stats(acc)
        if float(acc[2].item()) == 0.0:
            logger.warning(f"No valid samples at step {step}; resampling next step.")
            return None
        E_mean = sharding.energy_mean_and_variance(acc)[0].to(torch.float32)
        self.optimizer.zero_grad()
        if log_psi_vals is not None:
            loss = 2 * ((local_energies.detach() - E_mean) * log_psi_vals).mean()
            loss.backward()
        else:
            loss = torch.zeros((), device=self.device)
            for p in self.model.parameters():
                p.grad = torch.zeros_like(p)
        sharding.allreduce_mean_gradients(self.model.parameters())
        grad_norm = torch.nn.utils.clip_grad_norm_(self.model.parameters(), max_norm=10.0)
        self.optimizer.step()
        self.scheduler.step()
        env_up, env_down = self.model.orbital_head.envelope_up, self.model.orbital_head.envelope_down
        metrics = {
            "Energy": E_mean, "loss": loss, "step_time_sec": time.perf_counter() - t0,
            "grad_norm": grad_norm.item() if grad_norm is not None else .0,
            "lr": self.optimizer.param_groups[0]["lr"],
            "env_up_pi_norm": env_up.pi.detach().norm().item(),
            "env_up_sigma_norm": env_up.raw_sigma.detach().norm().item(),
            "env_down_pi_norm": env_down.pi.detach().norm().item(),
            "env_down_sigma_n": env_down.raw_sigma.detach().norm().item(),
            "mh_acceptance": self.mh.window_acceptance(),
            "energy_variance": float(sharding.energy_mean_and_variance(acc)[1].item()),
        }
        if getattr(self.config, "adapt_step_size", False):
            metrics["mh_step_size"] = self.mh.adapt_step_size()
        logger.info(f"Step {step}: E_mean = {E_mean.item():.6f}")
        logger.info(f"Loss = {loss.item():.6f}")
        return metrics

    def train(self):
        run = None
        if self.config.wand_mode != "disabled":
            run = self.config.init_wandb(self.model.config)
        train_start = time.perf_counter()
        torch.cuda.reset_peak_memory_stats(self.device)
        for step in range(self.config.train_steps):
            metrics = self.train_step(step)
            if metrics is None:
                continue
            torch.cuda.synchronize(self.device)
            metrics["gpu/mem_allocated_mb"] = torch.cuda.memory_allocated(self.device) / 2**20
            metrics["gpu/mem_reserved_mb"] = torch.cuda.memory_reserved(self.device) / 2**20
            self.history.append({k: (float(v) if torch.is_tensor(v) else v) for k, v in metrics.items()})
            if run is not None:
                run.log(metrics)
        total_time = time.perf_counter() - train_start
        logger.info(f"Total train time:{total_time/60:.2f} min ({total_time:.1f}sec)")
        if run is not None:
            run.log({"total_training_time_sec": total_time})
            run.finish()


def wrapper(preset: str, run_name: str = "", checkpoint_name: str = "", wand_mode: str = "") -> tuple:
    """Preset selector of train.py:197-246: returns *copies* of (model_config, train_config)."""
    key = (preset or "debug").lower()
    table = {"large": (LARGE_CONF, "_LARGE"), "small": (SMALL_CONF, "_SMALL"), "debug": (DEBUG_CONF, "_DEBUG"),
             "": (DEBUG_CONF, "_DEBUG")}
    if key not in table:
        raise ValueError(f"Unknown preset {preset!r}; expected 'debug', 'small' or 'large'.")
    (base_model, base_train), suffix = table[key]
    model_config, train_config = replace(base_model), replace(base_train)
    train_config.run_name = f"{run_name or train_config.run_name}{suffix}"
    base_ckpt = checkpoint_name or train_config.checkpoint_name
    train_config.checkpoint_name = f"{base_ckpt}{suffix}" if base_ckpt else f"{train_config.run_name}"
    if wand_mode:
        train_config.wand_mode = wand_mode
    logger.info("Selected preset=%s run_name=%s checkpoint_name=%s", key, train_config.run_name,
                train_config.checkpoint_name)
    return model_config, train_config


if __name__ == "__main__":
    logging.basicConfig(level=logging.INFO, format="%(asctime)s %(name)s %(levelname)s: %(message)s")
    model_configs = wrapper("small", run_name="Helium", checkpoint_name="Helium", wand_mode="offline")
    Trainer(PsiFormer(model_configs[0]), model_configs[1], True).train()
