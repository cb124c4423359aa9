"""Synthetic-input generator shared by the oracle and (bit-for-bit) by the CUDA library.

TEST INFRASTRUCTURE.  The CUDA twin is ``csrc/synth.cuh``; ``tests/test_synth_parity.py`` checks
the two produce identical bits.  SURVEY.md §8(d) asks for Philox-4x32-10 with counter = linear
index ``i + j*n``.  To be *bit-reproducible* across libm / CUDA math libraries the "normal"
variates avoid transcendental functions: one Philox call gives 128 random bits = eight uint16
words whose centred sum (Irwin-Hall, k=8) is scaled to unit variance with ONE fp64 multiply.
Excess kurtosis is -0.15; for benchmark inputs that is indistinguishable from N(0,1).

Streams (Philox counter word 2): 0 = matrix A, 1 = x_true mask, 2 = x_true values,
3 = label uniforms / noise, 4 = x0.
"""
from __future__ import annotations

import numpy as np

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = 0x9E3779B9
_W1 = 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)

# std of a sum of eight independent uniform{0..65535}: sqrt(8*(65536^2-1)/12)
IH8_MEAN = 4.0 * 65535.0
IH8_INV_STD = 1.0 / np.sqrt(8.0 * (65536.0 * 65536.0 - 1.0) / 12.0)

STREAM_A, STREAM_XMASK, STREAM_XVAL, STREAM_Y, STREAM_X0 = 0, 1, 2, 3, 4


def philox4x32_10(idx: np.ndarray, stream: int, seed: int):
    """counter = (idx_lo, idx_hi, stream, 0), key = (seed_lo, seed_hi).  Returns 4 uint32 arrays."""
    idx = np.asarray(idx, dtype=np.uint64)
    c0 = idx & _MASK
    c1 = idx >> np.uint64(32)
    c2 = np.full_like(c0, np.uint64(stream))
    c3 = np.zeros_like(c0)
    k0 = seed & 0xFFFFFFFF
    k1 = (seed >> 32) & 0xFFFFFFFF
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)), lo1, (hi0 ^ c3 ^ np.uint64(k1)), lo0
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def ih8_normal(idx, stream, seed):
    """Unit-variance Irwin-Hall(8) variate per counter (exact integer sum, one fp64 multiply)."""
    o = philox4x32_10(idx, stream, seed)
    s = np.zeros(np.shape(o[0]), dtype=np.uint64)
    for w in o:
        s = s + (w & np.uint64(0xFFFF)) + (w >> np.uint64(16))
    return (s.astype(np.float64) - IH8_MEAN) * IH8_INV_STD


def uniform53(idx, stream, seed):
    """Uniform in [0,1) with 53 bits: ((o0<<32 | o1) >> 11) * 2^-53."""
    o0, o1, _, _ = philox4x32_10(idx, stream, seed)
    v = ((o0 << np.uint64(32)) | o1) >> np.uint64(11)
    return v.astype(np.float64) * (1.0 / 9007199254740992.0)


def make_A(n: int, m: int, seed: int = 1234, row0: int = 0, n_total: int | None = None, density: float = 1.0):
    """Rows [row0, row0+n) of the n_total x m matrix, column-major semantics: counter = i + j*n_total.
    A_ij = ih8 / sqrt(m); with density<1 entries are kept when uniform53(stream A, word pair 2/3) < density."""
    n_total = n if n_total is None else n_total
    i = np.arange(row0, row0 + n, dtype=np.uint64)[:, None]
    j = np.arange(m, dtype=np.uint64)[None, :]
    idx = i + j * np.uint64(n_total)
    A = ih8_normal(idx, STREAM_A, seed) * (1.0 / np.sqrt(float(m)))
    if density < 1.0:
        _, _, o2, o3 = philox4x32_10(idx, STREAM_A, seed + 7)
        u = (((o2 << np.uint64(32)) | o3) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
        A = np.where(u < density, A, 0.0)
    return np.asfortranarray(A)


def make_x_true(m: int, seed: int = 1235, frac: float = 0.05, sigma: float = 3.0):
    j = np.arange(m, dtype=np.uint64)
    keep = uniform53(j, STREAM_XMASK, seed) < frac
    return np.where(keep, sigma * ih8_normal(j, STREAM_XVAL, seed), 0.0)


def make_x0(m: int, seed: int = 1237):
    return ih8_normal(np.arange(m, dtype=np.uint64), STREAM_X0, seed)


def make_labels_logistic(z: np.ndarray, seed: int = 1236, row0: int = 0):
    """y_i = +1 if u_i < 1/(1+exp(-z_i)) else -1 (z = A x_true computed by the caller)."""
    i = np.arange(row0, row0 + z.shape[0], dtype=np.uint64)
    u = uniform53(i, STREAM_Y, seed)
    return np.where(u < 1.0 / (1.0 + np.exp(-z)), 1.0, -1.0)


def make_targets_ls(z: np.ndarray, seed: int = 1236, row0: int = 0, noise: float = 0.1):
    i = np.arange(row0, row0 + z.shape[0], dtype=np.uint64)
    return z + noise * ih8_normal(i, STREAM_Y, seed)
