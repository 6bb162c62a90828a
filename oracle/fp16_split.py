"""TEST INFRASTRUCTURE (CPU oracle) -- never imported by the product package.

numpy restatement of the operand format of the tensor-core Linear layers ("fp16 split"):

    x  =  h0 + 2^-11 * h1,      h0 = rn_fp16(x),   h1 = rn_fp16((x - h0) * 2^11)

and of the three-pass product the kernel accumulates in fp32 (psiformer_torch_b200/csrc/gemm_tcgen05.cuh: the
splitter, `tc_split_weights_h_kernel`, and the epilogue's `main + corr / 2048`):

    x . w  ~=  sum h0x h0w  +  2^-11 * sum (h0x h1w + h1x h0w)          (the h1x h1w term, 2^-22, is dropped)

The reference computes these layers as plain fp32 `nn.Linear` (psiformer.py:32-93); this file exists to pin down
why three fp16 passes reproduce that to fp32 accuracy and where the range guard has to trip.
"""
import numpy as np

LO_SCALE = 2048.0           # H_LO_SCALE in gemm_tcgen05.cuh
FP16_MAX = 65504.0


def split(x):
    """fp32 array -> (h0, h1) fp16 arrays, round to nearest even like __float2half_rn."""
    x = np.asarray(x, dtype=np.float32)
    with np.errstate(over="ignore"):
        h0 = x.astype(np.float16)
        h1 = ((x - h0.astype(np.float32)) * np.float32(LO_SCALE)).astype(np.float16)
    return h0, h1


def join(h0, h1):
    return h0.astype(np.float64) + h1.astype(np.float64) / LO_SCALE


def in_range(x):
    """What the kernel's range guard (PSIF_ST_FP16_RANGE) protects: h0 must not overflow."""
    return bool(np.all(np.abs(np.asarray(x, dtype=np.float32)) <= FP16_MAX))


def linear_three_pass(x, w):
    """Y = X W^T from the three fp16 products, fp32 accumulation in K order emulated by float32 matmuls of exactly
    representable fp16 inputs (every product of two fp16 numbers is exact in fp32)."""
    x0, x1 = split(x)
    w0, w1 = split(w)
    f = np.float32
    main = x0.astype(f) @ w0.astype(f).T
    corr = x0.astype(f) @ w1.astype(f).T + x1.astype(f) @ w0.astype(f).T
    return main + corr * f(1.0 / LO_SCALE)
