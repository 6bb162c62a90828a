"""CPU checks of the fp16-split operand format restated in oracle/fp16_split.py (the format the tensor-core Linear
kernel uses on the GPU; the GPU kernel itself is checked in tests/test_stages_gpu.py)."""
import numpy as np
import pytest

from oracle import fp16_split as S


def test_pair_carries_22_bits_above_an_absolute_floor():
    rng = np.random.default_rng(0)
    x = (rng.standard_normal(400_000) * np.exp(rng.uniform(-20, 10, 400_000))).astype(np.float32)
    x = x[np.abs(x) <= S.FP16_MAX]
    assert S.in_range(x)
    h0, h1 = S.split(x)
    ax = np.abs(x.astype(np.float64))
    err = np.abs(S.join(h0, h1) - x.astype(np.float64))
    # 11 bits in h0 and 11 more in h1 (2^-22 relative, a factor 4 above fp32's own 2^-24), plus an ABSOLUTE floor:
    # the low half is an fp16 too, its subnormal step 2^-24 is worth 2^-24 / 2048 / 2 = 2^-36 after scaling back.
    # So the pair is not a relative format for tiny |x|; activations after LayerNorm / weights are O(1e-3 .. 10).
    assert np.all(err <= 2.0 ** -22 * ax + 2.0 ** -36)
    normal = ax >= 2.0 ** -13
    assert (err[normal] / ax[normal]).max() <= 2.0 ** -22


def test_small_values_survive_through_the_scaled_low_half():
    x = np.array([1e-7, -3e-8, 5e-9], dtype=np.float32)      # at / below fp16's subnormal step for h0
    h0, h1 = S.split(x)
    assert np.allclose(S.join(h0, h1), x, rtol=2e-3, atol=0)


def test_range_guard_boundary():
    assert S.in_range(np.array([65504.0, -65504.0], dtype=np.float32))
    assert not S.in_range(np.array([65536.0], dtype=np.float32))
    h0, _ = S.split(np.array([1e5], dtype=np.float32))
    assert np.isinf(h0[0])                                   # what the guard exists to catch


@pytest.mark.parametrize("M,K,N", [(640, 256, 256), (512, 1024, 256), (300, 256, 160)])
def test_three_pass_product_is_fp32_grade(M, K, N):
    rng = np.random.default_rng(M + K + N)
    x = rng.standard_normal((M, K)).astype(np.float32)
    w = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    ref = x.astype(np.float64) @ w.astype(np.float64).T
    scale = np.abs(ref).max()
    err_split = np.abs(S.linear_three_pass(x, w) - ref).max() / scale
    err_fp32 = np.abs(x @ w.T - ref).max() / scale
    # same normalisation and the same bound the GPU stage test uses for the kernel (2e-6); and no worse than a
    # handful of plain-fp32 errors
    assert err_split < 2e-6
    assert err_split < 8 * err_fp32 + 1e-7
