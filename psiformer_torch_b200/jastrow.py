"""``Jastrow``: parameter container + standalone evaluation (jastrow.py:5-87 of the reference)."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib as L


class Jastrow(nn.Module):
    def __init__(self, spin_up: int, spin_down: int):
        super().__init__()
        self.alpha_anti = nn.Parameter(torch.rand(1))
        self.alpha_par = nn.Parameter(torch.rand(1))
        self.spin_up = spin_up
        self.spin_down = spin_down

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if x.dim() == 2:
            x = x.unsqueeze(0)
        if x.size(1) < 2:
            raise ValueError("Jastrow requires at least two electrons.")
        if x.device.type != "cuda":
            raise RuntimeError("psiformer_torch_b200.Jastrow runs on CUDA tensors only (no CPU fallback)")
        xc = x.detach().to(torch.float32).contiguous()
        out = torch.empty(xc.shape[0], dtype=torch.float32, device=xc.device)
        n_dn = min(self.spin_down, x.size(1) - self.spin_up)
        with torch.cuda.device(xc.device):
            L.check(L.load().psif_jastrow(L.ptr(xc), xc.shape[0], self.spin_up, n_dn, float(self.alpha_par.detach()),
                                          float(self.alpha_anti.detach()), L.ptr(out),
                                          torch.cuda.current_stream(xc.device).cuda_stream))
        return out.to(x.dtype)
