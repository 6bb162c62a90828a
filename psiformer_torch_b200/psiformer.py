"""``PsiFormer``: the drop-in for psiformer.py:196-264 of the reference.

The module tree exists to hold parameters under the reference's state_dict names;
the arithmetic of ``forward`` is one call into libpsiformer_b200.so (fused embed -> L x [LN, QKV,
attention, proj, LN, MLP] -> orbital*envelope -> multi-determinant slogdet -> + Jastrow).  The
sub-modules therefore have no ``forward`` of their own: nothing in this package computes the
wavefunction with PyTorch ops.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.nn as nn

from . import _lib as L
from .config import Model_Config
from .engine import Engine
from .jastrow import Jastrow

try:  # the reference mixes this in (psiformer.py:196-199); optional here
    from huggingface_hub import PyTorchModelHubMixin as _HubMixin
except Exception:  # pragma: no cover
    class _HubMixin:  # type: ignore
        def __init_subclass__(cls, **kw):
            super().__init_subclass__()


def get_device() -> torch.device:
    """psiformer.py:13-16, minus the CPU branch: this package has no CPU path."""
    if not torch.cuda.is_available():
        raise RuntimeError("psiformer_torch_b200 needs a CUDA device (there is no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


class _ParamsOnly(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError(f"{type(self).__name__} only holds parameters; call PsiFormer.forward (fused CUDA path)")


class MHA(_ParamsOnly):
    def __init__(self, config: Model_Config):
        super().__init__()
        assert config.n_embd % config.n_head == 0
        self.c_attn = nn.Linear(config.n_embd, 3 * config.n_embd)
        self.c_proj = nn.Linear(config.n_embd, config.n_embd)
        self.n_head, self.n_embd = config.n_head, config.n_embd


class MLP(_ParamsOnly):
    def __init__(self, config: Model_Config):
        super().__init__()
        self.c_fc = nn.Linear(config.n_embd, 4 * config.n_embd)
        self.c_proj = nn.Linear(4 * config.n_embd, config.n_embd)


class Layer(_ParamsOnly):
    def __init__(self, config: Model_Config):
        super().__init__()
        self.attn = MHA(config)
        self.mlp = MLP(config)
        self.ln_1 = nn.LayerNorm(config.n_embd)
        self.ln_2 = nn.LayerNorm(config.n_embd)


class Envelope(_ParamsOnly):
    def __init__(self, natom: int, det_spin: int, sigma_init: float = 0.5):
        super().__init__()
        self.pi = nn.Parameter(torch.ones(natom, det_spin))
        self.raw_sigma = nn.Parameter(torch.full((natom, det_spin), sigma_init))


class Orbital_Head(_ParamsOnly):
    def __init__(self, config: Model_Config) -> None:
        super().__init__()
        self.n_det = config.n_determinants
        self.n_spin_up, self.n_spin_down = config.n_spin_up, config.n_spin_down
        self.n_atom = len(config.resolved_nuclei())
        self.envelope_up = Envelope(self.n_atom, self.n_det * self.n_spin_up)
        self.envelope_down = Envelope(self.n_atom, self.n_det * self.n_spin_down)
        self.n_embd = config.n_embd
        self.orb_up = nn.Linear(self.n_embd, self.n_det * self.n_spin_up)
        self.orb_down = nn.Linear(self.n_embd, self.n_det * self.n_spin_down)
        nn.init.constant_(self.orb_up.bias, 1e-3)
        nn.init.constant_(self.orb_down.bias, 1e-3)
        self.det_logits = nn.Parameter(torch.zeros(self.n_det))


class _LogPsi(torch.autograd.Function):
    """log|psi| with the parameter backward that train.py:148 needs (psif_logpsi_backward)."""

    @staticmethod
    def forward(ctx, model: "PsiFormer", x: torch.Tensor, *params: torch.Tensor):
        eng = model.engine(x.device)
        eng.sync_params(params)
        logabs, sign, status = eng.logpsi(x)
        ctx.model = model
        ctx.shapes = [tuple(p.shape) for p in params]
        ctx.needs = [p.requires_grad for p in params]
        # MH.sampler() hands out inference tensors (mcmc.py:56); those cannot be saved for backward as they are
        ctx.save_for_backward(x.clone() if x.is_inference() else x)
        ctx.mark_non_differentiable(sign, status)
        return logabs, sign, status

    @staticmethod
    def backward(ctx, g_log, g_sign, g_status):
        (x,) = ctx.saved_tensors
        flat = ctx.model.engine(x.device).logpsi_backward(x, g_log)
        grads, o = [], 0
        for shape, need in zip(ctx.shapes, ctx.needs):
            n = 1
            for s in shape:
                n *= s
            grads.append(flat[o:o + n].view(shape) if need else None)
            o += n
        return (None, None, *grads)


class _LogPsiCached(torch.autograd.Function):
    """log|psi| values that the fused sampler/energy pass has ALREADY computed for ``x``, with the same parameter
    backward as ``_LogPsi``: the score-function loss of train.py:141 needs d log|psi| / d params, not another forward."""

    @staticmethod
    def forward(ctx, model: "PsiFormer", x: torch.Tensor, logabs: torch.Tensor, *params: torch.Tensor):
        ctx.model = model
        ctx.shapes = [tuple(p.shape) for p in params]
        ctx.needs = [p.requires_grad for p in params]
        ctx.save_for_backward(x.clone() if x.is_inference() else x)
        return logabs.clone()

    @staticmethod
    def backward(ctx, g_log):
        (x,) = ctx.saved_tensors
        eng = ctx.model.engine(x.device)
        flat = eng.logpsi_backward(x, g_log)
        grads, o = [], 0
        for shape, need in zip(ctx.shapes, ctx.needs):
            n = 1
            for s in shape:
                n *= s
            grads.append(flat[o:o + n].view(shape) if need else None)
            o += n
        return (None, None, None, *grads)


class PsiFormer(nn.Module, _HubMixin, repo_url=""):
    """Convention (psiformer.py:213-217): the first ``n_spin_up`` electrons are spin-up."""

    def __init__(self, config: Model_Config):
        super().__init__()
        self.config = config
        natom = len(config.resolved_nuclei())
        self.l_0 = nn.Linear((config.n_features + 1) * natom, config.n_embd)
        self.layers = nn.ModuleList([Layer(config) for _ in range(config.n_layer)])
        self.orbital_head = Orbital_Head(config)
        self.jastrow = Jastrow(config.n_spin_up, config.n_spin_down)
        self.spin_up_idx = list(range(config.n_spin_up))
        self.spin_down_idx = list(range(config.n_spin_up, config.n_spin_up + config.n_spin_down))
        self._engines = {}
        self.last_sign: Optional[torch.Tensor] = None
        self.last_status: Optional[torch.Tensor] = None

    # ---- engine plumbing ----------------------------------------------------------------------
    def engine(self, device: torch.device) -> Engine:
        device = torch.device(device)
        if device.type == "cuda" and device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        eng = self._engines.get(device)
        if eng is None:
            c = self.config
            eng = Engine(n_layer=c.n_layer, n_head=c.n_head, n_embd=c.n_embd, n_det=c.n_determinants,
                         n_up=c.n_spin_up, n_dn=c.n_spin_down, nuclei=c.resolved_nuclei(), device=device)
            self._engines[device] = eng
        return eng

    def ready_engine(self, device: torch.device) -> Engine:
        eng = self.engine(device)
        eng.sync_params(list(self.parameters()))
        return eng

    def _flatten(self, x: torch.Tensor) -> torch.Tensor:
        if x.dim() > 3:
            x = x.reshape(-1, x.size(-2), x.size(-1))
        if x.dim() != 3 or tuple(x.shape[1:]) != (self.config.n_electron_num, self.config.n_features):
            error = f"x shape: {tuple(x.shape)}{self.config.n_electron_num, self.config.n_features}"
            raise ValueError("Input model shape mismatch", error)
        if x.device.type != "cuda":
            raise RuntimeError("psiformer_torch_b200 evaluates on CUDA tensors only (no CPU fallback)")
        return x

    # ---- the reference's public call ------------------------------------------------------------
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x: (..., n_electron, 3) -> log|psi| (B,).  Raises like psiformer.py:228-231, 256-257."""
        x = self._flatten(x)
        params = list(self.parameters())
        if torch.is_grad_enabled() and any(p.requires_grad for p in params):
            logabs, sign, status = _LogPsi.apply(self, x, *params)
        else:
            eng = self.engine(x.device)
            eng.sync_params(params)
            logabs, sign, status = eng.logpsi(x)
        self.last_sign, self.last_status = sign, status
        # one device->host sync per call, as the reference's isfinite check incurs (psiformer.py:256)
        if bool((status & L.ST_NONFINITE_LOGDET).any()):
            raise ValueError("Non-finite log determinant detected")
        return logabs

    def log_psi_cached(self, x: torch.Tensor, logabs: torch.Tensor) -> torch.Tensor:
        """``logabs`` = log|psi|(x) as already evaluated by the library on the current parameters (e.g. by
        ``MH.sample_energies``), returned as a tensor whose backward is the parameter gradient of log|psi| at ``x``."""
        x = self._flatten(x)
        params = list(self.parameters())
        self.engine(x.device).sync_params(params)
        return _LogPsiCached.apply(self, x, logabs.reshape(-1), *params)

    def log_psi_and_sign(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """The (log|psi|, sign) pair; the reference computes the sign and drops it (psiformer.py:191)."""
        with torch.no_grad():
            la = self.forward(x)
        return la, self.last_sign
