"""Model / training configuration: the drop-in counterpart of the reference's config.py.

Field names and defaults follow config.py:6-43 and the three presets config.py:77-156 so that
code written against the reference keeps working.  Additions (all optional, defaults reproduce the
reference): ``Model_Config.nuclei`` for molecules and ``Train_Config.seed``.
``wandb`` is imported only when ``init_wandb`` is called.
"""
from __future__ import annotations

import os
from dataclasses import asdict, dataclass
from typing import Optional, Tuple

Nucleus = Tuple[float, Tuple[float, float, float]]


@dataclass
class Model_Config():
    n_layer: int = 4
    n_head: int = 32
    n_embd: int = 512
    n_features: int = 3
    n_determinants: int = 1
    n_electron_num: int = 2
    n_spin_up: int = 1
    n_spin_down: int = 1
    nuclear_charge: int = 6
    # ((Z, (x, y, z)), ...) in bohr; None = one nucleus of charge `nuclear_charge` at the origin
    nuclei: Optional[Tuple[Nucleus, ...]] = None

    def resolved_nuclei(self) -> Tuple[Nucleus, ...]:
        if self.nuclei is None:
            return ((float(self.nuclear_charge), (0.0, 0.0, 0.0)),)
        return tuple((float(z), (float(r[0]), float(r[1]), float(r[2]))) for z, r in self.nuclei)


@dataclass
class Train_Config():
    train_steps: int = 1000
    checkpoint_step: int = 333
    batch_size: int = 2
    checkpoint_name: str = ""
    energy_batch_size: int = 128
    dim: int = 3
    lr: float = 3e-4

    entity: str = "alvaro18ml-university-of-minnesota"
    project: str = "Psiformer"
    run_name: str = "Train"
    wand_mode: str = "online"

    monte_carlo_length: int = 1024
    burn_in_steps: int = 4
    step_size: float = 1.0
    mh_steps_per_sample: int = 32

    repo_id: str = "jorgemunozl/psiformer_torch"
    checkpoint_dir: str = "./checkpoints"

    # Philox seed of the device sampler; None = drawn from torch's global generator at first use
    seed: Optional[int] = None
    # step-size adaptation hook of the sampler (MH.adapt_step_size) after every training step; off = the reference
    adapt_step_size: bool = False

    def init_checkpoint(self):
        os.makedirs(self.checkpoint_dir, exist_ok=True)
        return os.path.join(self.checkpoint_dir, self.checkpoint_name)

    def check_name(self):
        name = self.checkpoint_name
        return name if name.endswith(".pth") else name + ".pth"

    def init_wandb(self, model_config: Model_Config):
        import wandb

        full = asdict(model_config) | asdict(self)
        mode = self.wand_mode if self.wand_mode in ("online", "offline", "disabled") else "online"
        return wandb.init(entity=self.entity, project=self.project, name=self.run_name, config=full, mode=mode)


PSIFORMER_TORCH_SMALL_MODEL = Model_Config(
    n_layer=1, n_head=16, n_embd=64, n_determinants=1, n_electron_num=2, n_spin_up=1, n_spin_down=1,
    nuclear_charge=2)
PSIFORMER_TORCH_SMALL_CONF = Train_Config(
    batch_size=2, checkpoint_step=33, train_steps=301, energy_batch_size=2048, monte_carlo_length=128,
    burn_in_steps=4, step_size=1.0, mh_steps_per_sample=128)

PSIFORMER_TORCH_LARGE_MODEL = Model_Config(
    n_layer=4, n_head=32, n_embd=256, n_determinants=4, n_electron_num=6, n_spin_up=4, n_spin_down=2,
    nuclear_charge=6)
PSIFORMER_TORCH_LARGE_CONF = Train_Config(
    batch_size=2, checkpoint_step=50, train_steps=175, energy_batch_size=1024, monte_carlo_length=1024,
    burn_in_steps=4, step_size=0.8, mh_steps_per_sample=128)

PSIFORMER_TORCH_DEBUG_MODEL = Model_Config(
    n_layer=1, n_head=2, n_embd=4, n_determinants=1, n_electron_num=3, n_spin_up=2, n_spin_down=1,
    nuclear_charge=3)
PSIFORMER_TORCH_DEBUG_CONF = Train_Config(
    batch_size=1, checkpoint_step=2, train_steps=10, energy_batch_size=4, monte_carlo_length=4,
    burn_in_steps=4, step_size=1.0, mh_steps_per_sample=4)

LARGE_CONF = (PSIFORMER_TORCH_LARGE_MODEL, PSIFORMER_TORCH_LARGE_CONF)
SMALL_CONF = (PSIFORMER_TORCH_SMALL_MODEL, PSIFORMER_TORCH_SMALL_CONF)
DEBUG_CONF = (PSIFORMER_TORCH_DEBUG_MODEL, PSIFORMER_TORCH_DEBUG_CONF)

# BASELINE.json systems C1..C5: (Model_Config, walkers per GPU, MH step size)
BENCH_SYSTEMS = {
    "He": (PSIFORMER_TORCH_SMALL_MODEL, 1024, 1.0),
    "Be": (Model_Config(4, 4, 256, 3, 16, 4, 2, 2, 4), 4096, 0.8),
    "LiH": (Model_Config(4, 4, 256, 3, 16, 4, 2, 2, 3, nuclei=((3.0, (0.0, 0.0, 0.0)), (1.0, (0.0, 0.0, 3.015)))), 8192, 0.8),
    "Ne": (Model_Config(4, 4, 256, 3, 16, 10, 5, 5, 10), 2048, 0.8),
    "N2": (Model_Config(4, 4, 256, 3, 32, 14, 7, 7, 7, nuclei=((7.0, (0.0, 0.0, -2.0)), (7.0, (0.0, 0.0, 2.0)))), 4096, 0.8),
}
