"""Helpers shared by the -m gpu tests: build an Engine straight from an oracle system."""
import torch

from psiformer_torch_b200 import _lib as L
from psiformer_torch_b200.engine import Engine


def make_engine(sysm, params, device=None):
    device = device or torch.device("cuda", 0)
    eng = Engine(n_layer=sysm.n_layer, n_head=sysm.n_head, n_embd=sysm.n_embd, n_det=sysm.n_det, n_up=sysm.n_up,
                 n_dn=sysm.n_dn, nuclei=sysm.nuclei, device=device)
    eng.set_params(torch.cat([v.reshape(-1) for v in params.values()]).to(device))
    return eng


def stream():
    return torch.cuda.current_stream().cuda_stream


def rel_err(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()
