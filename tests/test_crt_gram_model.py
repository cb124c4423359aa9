"""CPU model of the emulated-fp64 Gram's arithmetic (csrc/kernels_i8gram.cuh) in exact Python integers:
norm-equalised fixed point, residues modulo the library's moduli, Gram of residues, CRT.  It pins the two claims the
GPU path rests on, independently of the hardware: (1) with columns scaled to a common 2-norm T every entry of X'X is
below T^2 (Cauchy-Schwarz), so the symmetric CRT range P/2 > T^2 of the shortest sufficient moduli prefix recovers the
integer Gram exactly; (2) the only error is the rounding of X: 0.4/T * sqrt(w_max / mean w) of the diagonal scale per entry (standard
deviation), a few times that for the worst entry, whatever n is."""
import math

import numpy as np
import pytest

MODS = [256, 255, 253, 251, 247, 241, 239, 233, 229, 227, 223, 217, 211, 199, 197]  # csrc/i8_moduli.inc
LOG2P = [79.240952, 87.041852, 94.803403, 102.524503, 110.161127, 117.783179]        # prefixes of 10..15 moduli


def T_of(k, rows):
    """Largest column norm the prefix of k moduli holds (scs_lib.cu i8_setup)."""
    return (2.0 ** ((LOG2P[k - 10] - 1.0) / 2.0) - 0.5 * math.sqrt(rows)) * (1.0 - 1e-6)


def pick(bits, rows, rmax):
    T_cap = 2.0 ** 50 / rmax
    T_req = min(2.0 ** bits, T_cap)
    k = next((k for k in range(10, 16) if T_of(k, rows) >= T_req), 15)
    return k, min(T_of(k, rows), T_cap)


def sym_res(v, p):
    r = v % p
    return r - p if 2 * r >= p else r  # in [-p/2, p/2): fits int8 for p <= 256


def crt(res, mods):
    P = math.prod(mods)
    x = 0
    for r, p in zip(res, mods):
        Mi = P // p
        x += Mi * ((r * pow(Mi, -1, p)) % p)
    x %= P
    return x - P if x > P // 2 else x


def test_moduli_table_matches_the_products():
    for k in range(10, 16):
        assert abs(math.log2(math.prod(MODS[:k])) - LOG2P[k - 10]) < 1e-5
    for i, a in enumerate(MODS):
        for b in MODS[:i]:
            assert math.gcd(a, b) == 1


@pytest.mark.parametrize("n,m,bits", [(300, 6, 46), (1000, 5, 30), (64, 4, 48), (2000, 4, 46)])
def test_integer_gram_is_recovered_exactly_and_error_is_0p4_over_T(n, m, bits):
    rng = np.random.default_rng(n + bits)
    A = rng.standard_normal((n, m)) * rng.uniform(0.01, 100.0, m)  # columns of very different scale
    A[:, 0] *= rng.random(n) < 0.05                                   # a spiky column
    w = rng.uniform(0.0, 0.25, n)
    rows = -(-n // 128) * 128
    norms = np.linalg.norm(A, axis=0)
    rmax = (np.abs(A).max(axis=0) / norms).max()
    k, T = pick(bits, rows, rmax)
    assert T >= min(2.0 ** bits, 2.0 ** 50 / rmax)
    mods = MODS[:k]
    P = math.prod(mods)
    scale = T / (math.sqrt(w.max()) * norms)
    X = np.rint((np.sqrt(w)[:, None] * A) * scale)                    # exact integers below 2^50 in fp64
    assert np.abs(X).max() <= 2.0 ** 50
    Xi = [[int(v) for v in col] for col in X.T]
    # (1) range: every entry of the exact integer Gram is inside the symmetric CRT range
    R = [[sum(a * b for a, b in zip(Xi[j], Xi[l])) for l in range(m)] for j in range(m)]
    assert max(abs(v) for row in R for v in row) <= int(T * (1 + 1e-6) + 0.5 * math.sqrt(rows)) ** 2 < P // 2
    # residue planes (int8), Gram of residues accumulated per modulus, reduced, CRT
    for j in range(m):
        for l in range(j + 1):
            res = []
            for p in mods:
                xj = [sym_res(v, p) for v in Xi[j]]
                xl = [sym_res(v, p) for v in Xi[l]]
                assert all(-128 <= v <= 127 for v in xj)
                res.append(sum(a * b for a, b in zip(xj, xl)) % p)
            assert crt(res, mods) == R[j][l]
    # (2) accuracy against the fp64 Gram, relative to the diagonal scale
    G = np.array(R, dtype=np.float64) / np.outer(scale, scale)
    Gref = A.T @ (w[:, None] * A)
    d = np.sqrt(np.diag(Gref))
    err = np.max(np.abs(G - Gref) / np.outer(d, d))
    wbar = np.diag(Gref) / norms ** 2  # weighted mean of w per column
    assert err <= 2.0 / T * math.sqrt(w.max() / wbar.min()) + 1e-15, (err, 1 / T)


@pytest.mark.parametrize("n,m,frac_neg", [(300, 5, 0.2), (500, 4, 0.9), (200, 3, 1.0)])
def test_signed_weights_minority_compaction(n, m, frac_neg):
    """Weights of both signs (k_residues<.., true> / k_crt): X is built from sqrt(|w|); with Xc = the rows of
    the minority sign, sum_i w_i a_i a_i' = s_main * X'X + s_extra * Xc'Xc with (s_main, s_extra) = (+1, -2) when the
    negative rows are the minority and (-1, +2) otherwise.  The combination is taken on the residues (chunk sums weighted by
    the signs before the CRT), and the signed integer Gram stays inside the CRT range by the same Cauchy-Schwarz bound."""
    rng = np.random.default_rng(n)
    A = rng.standard_normal((n, m)) * rng.uniform(0.1, 10.0, m)
    w = rng.uniform(0.01, 0.25, n) * np.where(rng.random(n) < frac_neg, -1.0, 1.0)
    rows = -(-n // 128) * 128
    norms = np.linalg.norm(A, axis=0)
    k, T = pick(46, rows, (np.abs(A).max(axis=0) / norms).max())
    mods = MODS[:k]
    P = math.prod(mods)
    scale = T / (math.sqrt(np.abs(w).max()) * norms)
    X = np.rint((np.sqrt(np.abs(w))[:, None] * A) * scale)
    Xi = [[int(v) for v in col] for col in X.T]
    neg = w < 0
    minor_neg = neg.sum() <= n - neg.sum()
    sel = neg if minor_neg else ~neg
    s_main, s_extra = (1, -2) if minor_neg else (-1, 2)
    for j in range(m):
        for l in range(j + 1):
            exact = sum((-1 if ng else 1) * a * b for ng, a, b in zip(neg, Xi[j], Xi[l]))
            assert abs(exact) < P // 2
            res = []
            for p in mods:
                xj = [sym_res(v, p) for v in Xi[j]]
                xl = [sym_res(v, p) for v in Xi[l]]
                full = sym_res(sum(a * b for a, b in zip(xj, xl)), p)          # partial residue of the main SYRK
                comp = sym_res(sum(a * b for a, b, s in zip(xj, xl, sel) if s), p)  # ... of the compacted rows
                res.append((s_main * full + s_extra * comp) % p)
            assert crt(res, mods) == exact
    G = np.array([[sum((-1 if ng else 1) * a * b for ng, a, b in zip(neg, Xi[j], Xi[l])) for l in range(m)]
                  for j in range(m)], dtype=np.float64) / np.outer(scale, scale)
    Gref = A.T @ (w[:, None] * A)
    d = np.sqrt(np.diag(A.T @ (np.abs(w)[:, None] * A)))
    assert np.max(np.abs(G - Gref) / np.outer(d, d)) <= 2.0 / T * math.sqrt(np.abs(w).max() / np.abs(w).min()) + 1e-15
