"""ctypes binding of libpsiformer_b200.so (include/psiformer_b200.h).

There is no fallback of any kind: if the shared library is missing or a call fails, this module
raises.  Build it with ``python -m psiformer_torch_b200.build``.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

MAX_ATOMS = 8
MODE_VALUE, MODE_ENERGY = 0, 1
ST_NONFINITE_LOGDET, ST_CLAMP, ST_NONFINITE_ELOC, ST_FLOOR, ST_FP16_RANGE = 1, 2, 4, 8, 16
GEMM_FP16_SPLIT, GEMM_TF32_SPLIT = 0, 1
E_INVALID, E_CUDA, E_WORKSPACE, E_STATE = -1, -2, -3, -4

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "libpsiformer_b200.so")


class PsifConfig(C.Structure):
    _fields_ = [
        ("n_layer", C.c_int32), ("n_head", C.c_int32), ("n_embd", C.c_int32), ("n_det", C.c_int32),
        ("n_up", C.c_int32), ("n_dn", C.c_int32), ("natom", C.c_int32), ("reserved", C.c_int32),
        ("Z", C.c_double * MAX_ATOMS), ("R", (C.c_double * 3) * MAX_ATOMS),
    ]


class PsifError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libpsiformer_b200 error {code}: {msg}")
        self.code = code
        self.msg = msg


_vp, _i32, _i64, _u64, _sz, _f = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_size_t, C.c_float

# name -> (restype, argtypes); every symbol declared in include/psiformer_b200.h
SIGNATURES = {
    "psif_create": (_i32, [C.POINTER(PsifConfig), C.POINTER(_vp)]),
    "psif_destroy": (_i32, [_vp]),
    "psif_param_count": (_i32, [_vp, C.POINTER(_sz)]),
    "psif_set_params": (_i32, [_vp, _vp, _sz, _vp]),
    "psif_set_gemm_mode": (_i32, [_vp, _i32]),
    "psif_workspace_bytes": (_i32, [_vp, _i64, _i32, C.POINTER(_sz)]),
    "psif_logpsi": (_i32, [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _sz, _vp]),
    "psif_local_energy": (_i32, [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "psif_mh_steps": (_i32, [_vp, _vp, _vp, _vp, _i64, _i32, _f, _i32, _u64, _u64, _u64, _vp, _vp, _vp, _vp, _vp,
                             _vp, _vp, _sz, _vp]),
    "psif_sample_energy": (_i32, [_vp, _vp, _vp, _vp, _i64, _i32, _f, _i32, _u64, _u64, _u64, _vp, _vp, _vp, _vp, _vp,
                                  _vp, _vp, _sz, _vp]),
    "psif_slogdet_multi": (_i32, [_vp, _vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp, _vp, _vp]),
    "psif_logdet_matmul_grad": (_i32, [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp, _vp, _vp]),
    "psif_logdet_matmul_grad_grad": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp, _vp,
                                            _vp, _vp]),
    "psif_jastrow": (_i32, [_vp, _i64, _i32, _i32, _f, _f, _vp, _vp]),
    "psif_potential": (_i32, [_vp, _i64, _i32, _i32, C.POINTER(C.c_double), C.POINTER(C.c_double), _vp, _vp]),
    "psif_philox_normal": (_i32, [_u64, _u64, _u64, _i64, _i32, _vp, _vp, _vp]),
    "psif_logpsi_backward": (_i32, [_vp, _vp, _vp, _i64, _vp, _vp, _vp, _sz, _vp]),
    "psif_backward_workspace_bytes": (_i32, [_vp, _i64, C.POINTER(_sz)]),
    "psif_stage_embed": (_i32, [_vp, _vp, _i64, _i32, _vp, _vp]),
    "psif_stage_linear": (_i32, [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _vp, _vp]),
    "psif_stage_linear_tc": (_i32, [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp]),
    "psif_stage_pack": (_i32, [_vp, _i64, _i32, _vp, _vp, _vp]),
    "psif_stage_det_energy": (_i32, [_vp, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "psif_stage_layernorm": (_i32, [_vp, _vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp]),
    "psif_stage_attention": (_i32, [_vp, _i64, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    "psif_stage_attention_first_layer": (_i32, [_vp, _i64, _i32, _i32, _i32, _i32, _vp, _vp]),
    "psif_stage_gelu": (_i32, [_vp, _i64, _i32, _i32, _vp, _vp]),
    "psif_take_range_event": (_i32, [_vp, C.POINTER(C.c_int32)]),
    "psif_profile_enable": (_i32, [_vp, _i32]),
    "psif_profile_read": (_i32, [_vp, C.POINTER(C.c_double), _i32]),
    "psif_launch_count": (_i64, []),
    "psif_last_error": (C.c_char_p, []),
    "psif_version": (C.c_char_p, []),
}

_lock = threading.Lock()
_lib = None


def load() -> C.CDLL:
    """Load the shared library (once) and attach prototypes.  Raises if it has not been built."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise PsifError(E_STATE, f"{LIB_PATH} not found: run `python -m psiformer_torch_b200.build` "
                                         "(there is no CPU or PyTorch fallback)")
            lib = C.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)
                fn.restype = res
                fn.argtypes = args
            _lib = lib
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise PsifError(rc, load().psif_last_error().decode(errors="replace"))


def ptr(t) -> int | None:
    """Device pointer of a contiguous torch tensor (or None)."""
    if t is None:
        return None
    assert t.is_contiguous(), "tensor handed to libpsiformer_b200 must be contiguous"
    return t.data_ptr()


def launch_count() -> int:
    return int(load().psif_launch_count())


PROFILE_CLASSES = ("gemm", "attention", "layernorm", "gelu", "embed", "orbital", "det", "jastrow", "mh")


def profile_enable(handle, on: bool) -> None:
    """Per-handle kernel-class timing (``handle`` = Engine._handle)."""
    check(load().psif_profile_enable(handle, 1 if on else 0))


def profile_read(handle) -> dict:
    buf = (C.c_double * (len(PROFILE_CLASSES) * 4))()
    check(load().psif_profile_read(handle, buf, len(PROFILE_CLASSES)))
    return {name: {"groups": buf[4 * i], "ms": buf[4 * i + 1], "flops": buf[4 * i + 2], "bytes": buf[4 * i + 3]}
            for i, name in enumerate(PROFILE_CLASSES)}
