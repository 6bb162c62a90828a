"""``Potential`` / ``Hamiltonian``: drop-ins for hamiltonian.py:6-95.

``Hamiltonian.local_energy`` is ONE fused forward-Laplacian pass (psif_local_energy) instead of the
reference's 3N+2 autograd passes.  ``log_psi_fn`` must be backed by a ``PsiFormer`` (the module, its
``forward``, or a bound method of an object with a ``.model`` such as ``Trainer.log_psi``); any other
callable raises ``TypeError``: there is a single backend and no autograd fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Optional

import torch

from . import _lib as L


def _resolve_model(fn):
    from .psiformer import PsiFormer

    if isinstance(fn, PsiFormer):
        return fn
    owner = getattr(fn, "__self__", None)
    if isinstance(owner, PsiFormer):
        return owner
    model = getattr(owner, "model", None)
    if isinstance(model, PsiFormer):
        return model
    raise TypeError("Hamiltonian needs a log_psi_fn backed by a psiformer_torch_b200.PsiFormer "
                    "(module, module.forward or Trainer.log_psi); arbitrary callables are not supported")


class Potential():
    """Softened Coulomb potential of hamiltonian.py:15-35 (single nucleus at the origin by default)."""

    def __init__(self, coords: torch.Tensor, Z: int = 2, nuclei=None):
        self.coords = coords
        self.Z = Z
        self.nuclei = nuclei if nuclei is not None else ((float(Z), (0.0, 0.0, 0.0)),)

    def potential(self) -> torch.Tensor:
        x = self.coords
        if x.device.type != "cuda":
            raise RuntimeError("psiformer_torch_b200.Potential runs on CUDA tensors only (no CPU fallback)")
        xc = x.detach().to(torch.float32).contiguous()
        B, n = xc.shape[0], xc.shape[1]
        na = len(self.nuclei)
        Z = (C.c_double * na)(*[float(z) for z, _ in self.nuclei])
        R = (C.c_double * (3 * na))(*[float(c) for _, r in self.nuclei for c in r])
        out = torch.empty(B, dtype=torch.float32, device=xc.device)
        with torch.cuda.device(xc.device):
            L.check(L.load().psif_potential(L.ptr(xc), B, n, na, Z, R, L.ptr(out),
                                            torch.cuda.current_stream(xc.device).cuda_stream))
        return out.to(x.dtype)


class Hamiltonian():
    def __init__(self, log_psi_fn: Callable[[torch.Tensor], torch.Tensor], n_elec: int = 2, Z: int = 2):
        self.log_psi_fn = log_psi_fn
        self.n_elec = n_elec
        self.Z = Z
        self.model = _resolve_model(log_psi_fn)
        self.last: Optional[dict] = None

    def _device(self, x: torch.Tensor) -> torch.device:
        if x.device.type == "cuda":
            return x.device
        p = next(self.model.parameters())
        if p.device.type != "cuda":
            raise RuntimeError("model and samples are on the CPU; psiformer_torch_b200 has no CPU path")
        return p.device

    def _run(self, x: torch.Tensor, **want) -> dict:
        x = self.model._flatten(x.to(self._device(x)))
        eng = self.model.ready_engine(x.device)
        out = eng.local_energy(x, **want)
        nuc = self.model.config.resolved_nuclei()
        if len(nuc) == 1 and float(self.Z) != float(nuc[0][0]):
            # the reference lets Hamiltonian.Z differ from Model_Config.nuclear_charge (hamiltonian.py:48)
            v_model = Potential(x, nuclei=nuc).potential()
            v_here = Potential(x, self.Z).potential()
            out["e_loc"] = out["e_loc"] - v_model + v_here
            if "pot" in out:
                out["pot"] = v_here
        self.last = out
        return out

    def local_energy(self, sample: torch.Tensor, accum: Optional[torch.Tensor] = None) -> torch.Tensor:
        """E_L = -1/2 (lap log psi + |grad log psi|^2) + V, (B,).  Non-finite entries stay non-finite so
        that train.py:86-90 masks them exactly as with the reference.  ``accum`` (optional, not in the reference:
        fp64[3] on the device) is incremented by {sum E_L, sum E_L^2, n} over the finite entries by the kernel itself --
        the operand of the energy all-reduce of a walker-sharded run (train.py:138)."""
        return self._run(sample, accum=accum)["e_loc"]

    def grad_log_psi(self, x: torch.Tensor) -> torch.Tensor:
        return self._run(x, want_grad=True)["grad"]

    def laplacian_log_psi(self, x: torch.Tensor) -> torch.Tensor:
        return self._run(x, want_lap=True)["lap"]
