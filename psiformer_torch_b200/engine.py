"""One ``Engine`` per PsiFormer module: owns the C handle, the packed-parameter upload and the
torch-allocated workspace, and exposes the three hot-path calls as tensor-in / tensor-out methods.

torch is used for device memory and streams only; every number is produced by
libpsiformer_b200.so.  Nothing here falls back to PyTorch math.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence, Tuple

import torch

from . import _lib as L


def _stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


class Engine:
    def __init__(self, *, n_layer: int, n_head: int, n_embd: int, n_det: int, n_up: int, n_dn: int,
                 nuclei: Sequence[Tuple[float, Tuple[float, float, float]]], device: torch.device):
        if device.type != "cuda":
            raise RuntimeError("psiformer_torch_b200 runs on CUDA devices only (no CPU fallback)")
        self.lib = L.load()
        self.device = device
        self.n_elec = n_up + n_dn
        cfg = L.PsifConfig()
        cfg.n_layer, cfg.n_head, cfg.n_embd, cfg.n_det = n_layer, n_head, n_embd, n_det
        cfg.n_up, cfg.n_dn, cfg.natom = n_up, n_dn, len(nuclei)
        if len(nuclei) > L.MAX_ATOMS:
            raise ValueError(f"at most {L.MAX_ATOMS} nuclei are supported")
        for a, (z, r) in enumerate(nuclei):
            cfg.Z[a] = float(z)
            for k in range(3):
                cfg.R[a][k] = float(r[k])
        self._handle = C.c_void_p()
        with torch.cuda.device(device):
            L.check(self.lib.psif_create(C.byref(cfg), C.byref(self._handle)))
        n = C.c_size_t()
        L.check(self.lib.psif_param_count(self._handle, C.byref(n)))
        self.n_params = int(n.value)
        self._ws: Dict[object, torch.Tensor] = {}
        self._param_key = None
        self._flat: Optional[torch.Tensor] = None
        # launch-bound regime (few payload rows, e.g. the SMALL preset): the ~40 launches of a local-energy pass are
        # replayed from a CUDA graph with static buffers instead of being issued one by one
        self.graph_max_rows = 1 << 16
        self._gemm_mode = L.GEMM_FP16_SPLIT
        self._energy_graphs: Dict[tuple, dict] = {}

    def __del__(self):
        try:
            if getattr(self, "_handle", None) is not None and self._handle.value:
                self.lib.psif_destroy(self._handle)
                self._handle = C.c_void_p()
        except Exception:
            pass

    # ---- parameters -------------------------------------------------------------------------
    def set_params(self, flat: torch.Tensor) -> None:
        flat = flat.detach().to(device=self.device, dtype=torch.float32).contiguous()
        if flat.numel() != self.n_params:
            raise ValueError(f"parameter blob has {flat.numel()} floats, expected {self.n_params}")
        with torch.cuda.device(self.device):
            L.check(self.lib.psif_set_params(self._handle, L.ptr(flat), flat.numel(), _stream_ptr(self.device)))
        self._flat = flat

    def sync_params(self, params: Sequence[torch.Tensor]) -> None:
        """Re-upload when any parameter tensor was replaced or modified in place."""
        key = tuple((p.data_ptr(), p._version) for p in params)
        if key != self._param_key:
            self.set_params(torch.cat([p.detach().reshape(-1).to(self.device, torch.float32) for p in params]))
            self._param_key = key

    # ---- workspace --------------------------------------------------------------------------
    def workspace_bytes(self, B: int, mode: int) -> int:
        need = C.c_size_t()
        L.check(self.lib.psif_workspace_bytes(self._handle, int(B), mode, C.byref(need)))
        return int(need.value)

    def workspace(self, B: int, mode: int) -> torch.Tensor:
        """Scratch shared by the eager calls of this engine.  Anything that outlives a call (a captured CUDA graph) must
        own its workspace instead (``new_workspace``): this tensor is replaced when a later call needs more room."""
        need = self.workspace_bytes(B, mode)
        ws = self._ws.get(mode)
        if ws is None or ws.numel() < need:
            self._ws[mode] = ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return ws

    def new_workspace(self, B: int, *modes: int) -> torch.Tensor:
        """A private workspace large enough for every mode in ``modes`` (owned by the caller, never replaced)."""
        need = max(self.workspace_bytes(B, m) for m in modes)
        return torch.empty(need, dtype=torch.uint8, device=self.device)

    def _check_x(self, x: torch.Tensor) -> torch.Tensor:
        if x.device != self.device:
            x = x.to(self.device)
        return x.detach().to(torch.float32).contiguous()

    # ---- precision guard ----------------------------------------------------------------------
    def set_gemm_mode(self, mode: int) -> None:
        """L.GEMM_FP16_SPLIT (default) or L.GEMM_TF32_SPLIT for the tensor-core Linear layers."""
        L.check(self.lib.psif_set_gemm_mode(self._handle, int(mode)))
        self._gemm_mode = int(mode)

    def _range_event(self, clear_only: bool = False) -> bool:
        """True when a forward pass of this handle saw an activation beyond fp16's range since the last look.  The
        library mirrors the event into pinned host memory, so this is ONE stream synchronise and a host read -- not a
        reduction over the status array plus a device->host copy (psif_take_range_event).  ``clear_only`` drops a stale
        event (raised by an unguarded call) without synchronising."""
        if not clear_only:
            torch.cuda.current_stream(self.device).synchronize()
        out = C.c_int32(0)
        L.check(self.lib.psif_take_range_event(self._handle, C.byref(out)))
        return bool(out.value)

    # ---- hot path ---------------------------------------------------------------------------
    def logpsi(self, x: torch.Tensor, *, guard: bool = True) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """``guard``: repeat the call with tf32-split GEMMs if an activation did not fit the fp16 split."""
        if guard:
            self._range_event(clear_only=True)
        out = self._logpsi(x)
        if guard and self._range_event():
            self.set_gemm_mode(L.GEMM_TF32_SPLIT)
            try:
                out = self._logpsi(x)
            finally:
                self.set_gemm_mode(L.GEMM_FP16_SPLIT)
        return out

    def _logpsi(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        x = self._check_x(x)
        B = x.shape[0]
        logabs = torch.empty(B, dtype=torch.float32, device=self.device)
        sign = torch.empty_like(logabs)
        status = torch.zeros(B, dtype=torch.int32, device=self.device)
        ws = self.workspace(B, L.MODE_VALUE)
        with torch.cuda.device(self.device):
            L.check(self.lib.psif_logpsi(self._handle, L.ptr(x), B, L.ptr(logabs), L.ptr(sign), L.ptr(status),
                                         L.ptr(ws), ws.numel(), _stream_ptr(self.device)))
        return logabs, sign, status

    def local_energy(self, x: torch.Tensor, *, want_grad: bool = False, want_lap: bool = False, want_pot: bool = False,
                     accum: Optional[torch.Tensor] = None, guard: bool = True) -> Dict[str, torch.Tensor]:
        """``guard``: look at the status words (one sync) and repeat the call with tf32-split GEMMs if an activation
        did not fit the fp16 split; ``guard=False`` keeps the call asynchronous and leaves PSIF_ST_FP16_RANGE to the
        caller."""
        before = accum.clone() if (guard and accum is not None) else None
        if guard:
            self._range_event(clear_only=True)
        out = self._local_energy_maybe_graphed(x, want_grad, want_lap, want_pot, accum)
        if guard and self._range_event():
            if accum is not None:
                accum.copy_(before)
            self.set_gemm_mode(L.GEMM_TF32_SPLIT)
            try:
                out = self._local_energy(x, want_grad, want_lap, want_pot, accum)
            finally:
                self.set_gemm_mode(L.GEMM_FP16_SPLIT)
        return out

    def _local_energy_maybe_graphed(self, x, want_grad, want_lap, want_pot, accum):
        x = self._check_x(x)
        B = x.shape[0]
        rows = B * self.n_elec * (3 * self.n_elec + 2)
        if B == 0 or rows > self.graph_max_rows or torch.cuda.is_current_stream_capturing():
            return self._local_energy(x, want_grad, want_lap, want_pot, accum)
        key = (B, want_grad, want_lap, want_pot, self._gemm_mode)      # a graph replays the kernels it was captured with
        g = self._energy_graphs.get(key)
        if g is None:
            g = {"x": torch.empty_like(x), "acc": torch.zeros(3, dtype=torch.float64, device=self.device),
                 "ws": self.new_workspace(B, L.MODE_ENERGY)}
            g["x"].copy_(x)
            self._local_energy(g["x"], want_grad, want_lap, want_pot, g["acc"], ws=g["ws"])      # warm-up (attributes, maps)
            torch.cuda.synchronize(self.device)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                g["acc"].zero_()
                g["out"] = self._local_energy(g["x"], want_grad, want_lap, want_pot, g["acc"], ws=g["ws"])
            g["graph"] = graph
            self._energy_graphs[key] = g
        g["x"].copy_(x)
        g["graph"].replay()
        if accum is not None:
            accum += g["acc"]
        return {k: v.clone() for k, v in g["out"].items()}

    def _local_energy(self, x: torch.Tensor, want_grad: bool, want_lap: bool, want_pot: bool,
                      accum: Optional[torch.Tensor], ws: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        x = self._check_x(x)
        B = x.shape[0]
        dev = self.device
        out = {
            "e_loc": torch.empty(B, dtype=torch.float32, device=dev),
            "logabs": torch.empty(B, dtype=torch.float32, device=dev),
            "sign": torch.empty(B, dtype=torch.float32, device=dev),
            "status": torch.zeros(B, dtype=torch.int32, device=dev),
        }
        if want_grad:
            out["grad"] = torch.empty(B, self.n_elec, 3, dtype=torch.float32, device=dev)
        if want_lap:
            out["lap"] = torch.empty(B, dtype=torch.float32, device=dev)
        if want_pot:
            out["pot"] = torch.empty(B, dtype=torch.float32, device=dev)
        if accum is not None:
            assert accum.dtype == torch.float64 and accum.numel() == 3 and accum.device == dev
        if ws is None:
            ws = self.workspace(B, L.MODE_ENERGY)
        with torch.cuda.device(dev):
            L.check(self.lib.psif_local_energy(
                self._handle, L.ptr(x), B, L.ptr(out["e_loc"]), L.ptr(out["logabs"]), L.ptr(out["sign"]),
                L.ptr(out.get("grad")), L.ptr(out.get("lap")), L.ptr(out.get("pot")), L.ptr(accum),
                L.ptr(out["status"]), L.ptr(ws), ws.numel(), _stream_ptr(dev)))
        return out

    def mh_steps(self, x: torch.Tensor, logabs: torch.Tensor, sign: Optional[torch.Tensor], n_steps: int,
                 step_size: float, *, have_logabs: bool, seed: int = 0, walker_id0: int = 0, step0: int = 0,
                 step_counter: Optional[torch.Tensor] = None, noise: Optional[torch.Tensor] = None,
                 uniforms: Optional[torch.Tensor] = None, accept_out: Optional[torch.Tensor] = None,
                 n_accept: Optional[torch.Tensor] = None, status: Optional[torch.Tensor] = None,
                 ws: Optional[torch.Tensor] = None) -> None:
        """In place on ``x`` (B,N,3) / ``logabs`` (B,) / ``sign`` (B,).  ``ws``: a workspace owned by the caller
        (``new_workspace``); required when the call is captured into a CUDA graph."""
        assert x.is_contiguous() and x.dtype == torch.float32 and x.device == self.device
        B = x.shape[0]
        if noise is not None:
            assert noise.shape == (max(1, n_steps), B, self.n_elec, 3) and noise.dtype == torch.float32
        if uniforms is not None:
            assert uniforms.shape == (max(1, n_steps), B) and uniforms.dtype == torch.float32
        if accept_out is not None:
            assert accept_out.shape == (max(1, n_steps), B) and accept_out.dtype == torch.uint8
        if ws is None:
            ws = self.workspace(B, L.MODE_VALUE)
        with torch.cuda.device(self.device):
            L.check(self.lib.psif_mh_steps(
                self._handle, L.ptr(x), L.ptr(logabs), L.ptr(sign), B, int(n_steps), float(step_size),
                1 if have_logabs else 0, int(seed) & (2**64 - 1), int(walker_id0), int(step0), L.ptr(step_counter),
                L.ptr(noise), L.ptr(uniforms), L.ptr(accept_out), L.ptr(n_accept), L.ptr(status), L.ptr(ws),
                ws.numel(), _stream_ptr(self.device)))

    def sample_energy(self, x: torch.Tensor, logabs: torch.Tensor, sign: Optional[torch.Tensor], n_steps: int,
                      step_size: float, *, have_logabs: bool, seed: int, walker_id0: int, step_counter: torch.Tensor,
                      n_accept: Optional[torch.Tensor], accum: Optional[torch.Tensor], ws: torch.Tensor
                      ) -> Dict[str, torch.Tensor]:
        """``n_steps`` Metropolis steps in place on the chain state, then one local-energy pass on it
        (psif_sample_energy).  Returns e_loc / logabs (of the energy pass) / status for the new sample."""
        assert x.is_contiguous() and x.dtype == torch.float32 and x.device == self.device
        B = x.shape[0]
        out = {"e_loc": torch.empty(B, dtype=torch.float32, device=self.device),
               "logabs": torch.empty(B, dtype=torch.float32, device=self.device),
               "status": torch.zeros(B, dtype=torch.int32, device=self.device)}
        with torch.cuda.device(self.device):
            L.check(self.lib.psif_sample_energy(
                self._handle, L.ptr(x), L.ptr(logabs), L.ptr(sign), B, int(n_steps), float(step_size),
                1 if have_logabs else 0, int(seed) & (2**64 - 1), int(walker_id0), 0, L.ptr(step_counter),
                L.ptr(n_accept), L.ptr(out["e_loc"]), L.ptr(out["logabs"]), L.ptr(accum), L.ptr(out["status"]),
                L.ptr(ws), ws.numel(), _stream_ptr(self.device)))
        return out

    def logpsi_backward(self, x: torch.Tensor, grad_out: torch.Tensor, *, tf32: bool = False) -> torch.Tensor:
        """Flat fp32 gradient (state_dict order) of sum_b grad_out[b] * log|psi|(x_b) with respect to the parameters.
        The forward recompute runs through the same split GEMMs as the forward: if an activation leaves fp16's range
        the library says so and the call is repeated with tf32-split operands (``tf32=True`` starts there)."""
        x = self._check_x(x)
        B = x.shape[0]
        g = grad_out.detach().to(self.device, torch.float32).reshape(-1).contiguous()
        assert g.numel() == B
        need = C.c_size_t()
        L.check(self.lib.psif_backward_workspace_bytes(self._handle, B, C.byref(need)))
        ws = self._ws.get("bwd")
        if ws is None or ws.numel() < need.value:
            self._ws["bwd"] = ws = torch.empty(int(need.value), dtype=torch.uint8, device=self.device)
        out = torch.empty(self.n_params, dtype=torch.float32, device=self.device)
        flag = torch.zeros(1, dtype=torch.int32, device=self.device)

        def run():
            with torch.cuda.device(self.device):
                L.check(self.lib.psif_logpsi_backward(self._handle, L.ptr(x), L.ptr(g), B, L.ptr(out), L.ptr(flag),
                                                      L.ptr(ws), ws.numel(), _stream_ptr(self.device)))
        if not tf32:
            run()
            tf32 = bool(flag.item() & L.ST_FP16_RANGE)
        if tf32:
            self.set_gemm_mode(L.GEMM_TF32_SPLIT)
            try:
                run()
            finally:
                self.set_gemm_mode(L.GEMM_FP16_SPLIT)
        return out

    # ---- stage hooks (tests) ----------------------------------------------------------------
    def stage_embed(self, x: torch.Tensor, C_: int, d: int) -> torch.Tensor:
        x = self._check_x(x)
        out = torch.empty(x.shape[0], self.n_elec, C_, d, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            L.check(self.lib.psif_stage_embed(self._handle, L.ptr(x), x.shape[0], C_, L.ptr(out), _stream_ptr(self.device)))
        return out
