"""psiformer_torch_b200: B200-native (sm_100a) VMC hot path behind psiformer_torch's Python API.

Public surface (same names as the reference package):
    config.Model_Config / Train_Config / presets, psiformer.PsiFormer / get_device,
    logdet_matmul / LogDetMatmul, jastrow.Jastrow, hamiltonian.Hamiltonian / Potential,
    mcmc.MH, train.Trainer / wrapper.
"""
from __future__ import annotations

import sys
from typing import Any

__all__ = ["logdet_matmul", "LogDetMatmul", "install_as"]


def __getattr__(name: str) -> Any:  # lazy, like the reference's __init__.py:12-16
    if name in ("logdet_matmul", "LogDetMatmul"):
        import importlib

        mod = importlib.import_module(__name__ + ".logdet_matmul")
        return getattr(mod, name)
    raise AttributeError(name)


def install_as(alias: str = "psiformer_torch") -> None:
    """Make ``import psiformer_torch...`` resolve to this package (drop-in switch for user code)."""
    import importlib

    sys.modules[alias] = sys.modules[__name__]
    for sub in ("config", "psiformer", "logdet_matmul", "jastrow", "hamiltonian", "mcmc", "train"):
        sys.modules[f"{alias}.{sub}"] = importlib.import_module(f"{__name__}.{sub}")
