"""numpy restatement of the device random streams (csrc/mh.cuh).  TEST INFRASTRUCTURE ONLY.

Philox4x32-10 (Salmon et al., SC'11; the Random123 constants) with the counter / key layout of
``mh_random``: counter = (walker_lo, walker_hi, step_lo, (step_hi << 8) | slot), key = seed.
The reference draws from torch's global generator (mcmc.py:29,33,42); a sharding-invariant stream
is new in this build, so this file is the oracle for it.
"""
from __future__ import annotations

import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(ctr: np.ndarray, key: np.ndarray) -> np.ndarray:
    """ctr (...,4) uint32, key (...,2) uint32 -> (...,4) uint32."""
    c = [ctr[..., i].astype(np.uint64) for i in range(4)]
    k0 = key[..., 0].astype(np.uint32).copy()
    k1 = key[..., 1].astype(np.uint32).copy()
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0, p1 = M0 * c[0], M1 * c[2]
            hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & MASK, p1 >> np.uint64(32), p1 & MASK
            c = [hi1 ^ c[1] ^ k0.astype(np.uint64), lo1, hi0 ^ c[3] ^ k1.astype(np.uint64), lo0]
            k0 = (k0 + W0).astype(np.uint32)
            k1 = (k1 + W1).astype(np.uint32)
    return np.stack([v.astype(np.uint32) for v in c], axis=-1)


def mh_random(seed: int, walker: np.ndarray, step: int, slot: np.ndarray) -> np.ndarray:
    walker = np.asarray(walker, dtype=np.uint64)
    slot = np.asarray(slot, dtype=np.uint64)
    walker, slot = np.broadcast_arrays(walker, slot)
    ctr = np.stack([walker & MASK, walker >> np.uint64(32), np.full(walker.shape, step & 0xFFFFFFFF, np.uint64),
                    np.full(walker.shape, ((step >> 32) << 8) & 0xFFFFFFFF, np.uint64) | (slot & np.uint64(0xFF))],
                   axis=-1).astype(np.uint32)
    key = np.broadcast_to(np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32), walker.shape + (2,))
    return philox4x32_10(ctr, key)


def mh_normals(seed: int, walker_id0: int, step: int, n_walkers: int, n_elec: int) -> np.ndarray:
    """(n_walkers, n_elec, 3) float32 proposals, Box-Muller as in mh_normals3."""
    w = (np.arange(n_walkers, dtype=np.uint64) + np.uint64(walker_id0))[:, None]
    r = mh_random(seed, w, step, np.arange(n_elec, dtype=np.uint64)[None, :])
    k = np.float32(2.0 ** -24)
    u0 = ((r[..., 0] >> 8).astype(np.float32) + np.float32(1)) * k
    u1 = (r[..., 1] >> 8).astype(np.float32) * k
    u2 = ((r[..., 2] >> 8).astype(np.float32) + np.float32(1)) * k
    u3 = (r[..., 3] >> 8).astype(np.float32) * k
    ra = np.sqrt(np.float32(-2) * np.log(u0))
    rb = np.sqrt(np.float32(-2) * np.log(u2))
    tp = np.float32(6.283185307179586)
    return np.stack([ra * np.cos(tp * u1), ra * np.sin(tp * u1), rb * np.cos(tp * u3)], axis=-1).astype(np.float32)


def mh_uniforms(seed: int, walker_id0: int, step: int, n_walkers: int) -> np.ndarray:
    w = np.arange(n_walkers, dtype=np.uint64) + np.uint64(walker_id0)
    r = mh_random(seed, w, step, np.full(n_walkers, 0xFF, dtype=np.uint64))
    return ((r[..., 0] >> 8).astype(np.float32) * np.float32(2.0 ** -24)).astype(np.float32)


def known_answer() -> bool:
    """Random123 kat_vectors: philox4x32-10, counter = key = 0 and all-ones."""
    z = philox4x32_10(np.zeros(4, np.uint32), np.zeros(2, np.uint32))
    o = philox4x32_10(np.full(4, 0xFFFFFFFF, np.uint32), np.full(2, 0xFFFFFFFF, np.uint32))
    return (list(z) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
            and list(o) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD])
