"""GPU parity, kernel by kernel, through the C ABI against the numpy oracle on the same seeded inputs.
Tolerances (fp64): 1e-12 relative on streaming reductions / Gram entries (different summation order only),
1e-13 on elementwise formulas, bit-exact for the synthetic generator and integer-like outputs."""
import numpy as np
import pytest

from oracle import scs_oracle as O
from oracle import synth

pytestmark = pytest.mark.gpu


def relerr(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def logistic_problem(n, m, seed=3):
    A = synth.make_A(n, m, seed=seed)
    y = synth.make_labels_logistic(A @ synth.make_x_true(m, seed=seed + 1, frac=0.3), seed=seed + 2)
    return A, y, synth.make_x0(m, seed=seed + 3) * 0.5


# ragged shapes: single row, odd n, one row short of / past a 512-row CTA block, m not a multiple of anything
SHAPES = [(1, 1), (5, 2), (100, 50), (511, 7), (513, 130), (1000, 33), (4096, 256), (3001, 257)]


@pytest.mark.parametrize("n,m", SHAPES)
@pytest.mark.parametrize("loss", ["logistic", "logistic_consistent", "ls"])
@pytest.mark.parametrize("stream", ["two_pass", "fused"])
def test_forward_adjoint(scs, n, m, loss, stream):
    A, y, x = logistic_problem(n, m)
    if loss == "ls":
        Lo, Lg = O.LeastSquaresLoss(float(n)), scs.LeastSquaresLoss(float(n))
        y = synth.make_targets_ls(A @ synth.make_x_true(m), seed=9)
    else:
        mode = "consistent" if loss.endswith("consistent") else "literal"
        Lo, Lg = O.LogisticLoss(1 / n, mode), scs.LogisticLoss(1 / n, mode)
    p = scs.Problem(A, y, x, Lg, 0.1)
    p.set_stream_mode(stream)
    z = A @ x
    for wk in ("newton", "ggn"):
        fv, g, zg, rg, wg = p.loss_eval(x, weights=wk, want_rows=True)
        assert p.stream_path() == stream
        r, w = (Lo.grad_weights(z, y), Lo.hess_weights(z, y)) if wk == "newton" else Lo.ggn_weights(z, y)
        assert relerr(zg, z) <= 1e-13
        assert abs(fv - Lo.f(A, y, x)) <= 1e-13 * abs(Lo.f(A, y, x))
        np.testing.assert_allclose(rg, r, rtol=1e-11, atol=1e-300)
        np.testing.assert_allclose(wg, w, rtol=1e-11, atol=1e-300)
        assert relerr(g, A.T @ r) <= 1e-12
    p.close()


# the single-pass cluster kernel across its geometries (1..16 CTAs of 256 columns per cluster); panel counts that
# are odd, smaller than the cluster count, and several per cluster
@pytest.mark.parametrize("n,m", [(16, 300), (47, 1024), (6000, 777), (5000, 2048), (700, 2300), (2100, 4096),
                                 (300, 3000), (20000, 512), (33, 4000)])
def test_fused_gradient_geometries(scs, n, m):
    A, y, x = logistic_problem(n, m)
    Lo, Lg = O.LogisticLoss(1 / n, "consistent"), scs.LogisticLoss(1 / n, "consistent")
    p = scs.Problem(A, y, x, Lg, 0.1)
    p.set_stream_mode("fused")
    z = A @ x
    fv, g, zg, rg, wg = p.loss_eval(x, weights="ggn", want_rows=True)
    assert p.stream_path() == "fused"
    r, w = Lo.ggn_weights(z, y)
    assert relerr(zg, z) <= 1e-13
    assert abs(fv - Lo.f(A, y, x)) <= 1e-13 * abs(Lo.f(A, y, x))
    np.testing.assert_allclose(rg, r, rtol=1e-11, atol=1e-300)
    np.testing.assert_allclose(wg, w, rtol=1e-11, atol=1e-300)
    assert relerr(g, A.T @ r) <= 1e-12
    # the two-pass kernels agree and a second fused call reproduces the first bit for bit
    fv2, g2, *_ = p.loss_eval(0.5 * x, weights="ggn")
    fv3, g3, *_ = p.loss_eval(x, weights="ggn")
    assert fv3 == fv and np.array_equal(g3, g)
    p.set_stream_mode("two_pass")
    fv4, g4, *_ = p.loss_eval(0.5 * x, weights="ggn")
    assert abs(fv4 - fv2) <= 1e-13 * abs(fv2) and relerr(g2, g4) <= 1e-12
    p.close()


@pytest.mark.parametrize("n,m", [(64, 4100), (700, 6000), (1030, 8192), (333, 9000)])
@pytest.mark.parametrize("loss", ["logistic_consistent", "ls"])
def test_fused_gradient_wider_than_one_cluster(scs, n, m, loss):
    """m > 16 x 256 columns: the cluster kernel covers the last 4096 columns and receives z of the others from a k_forward
    pass (z_in), their share of g comes from a k_adjoint pass: 1.5 reads of A.  Same loss, z, r, w, g as the oracle and as
    the two-pass kernels, window masking included."""
    A = synth.make_A(n, m, seed=5)
    x = synth.make_x0(m, seed=6) * 0.2
    if loss == "ls":
        y = synth.make_targets_ls(A @ synth.make_x_true(m, seed=7), seed=8)
        Lo, Lg = O.LeastSquaresLoss(float(n)), scs.LeastSquaresLoss(float(n))
    else:
        y = synth.make_labels_logistic(A @ synth.make_x_true(m, seed=7, frac=0.1), seed=8)
        Lo, Lg = O.LogisticLoss(1 / n, "consistent"), scs.LogisticLoss(1 / n, "consistent")
    p = scs.Problem(A, y, x, Lg, 0.1)
    for lo, hi in ((0, n), (n // 4 + 1, n - 3)):
        p.set_active_rows(lo, hi)
        out = {}
        for mode in ("fused", "two_pass"):
            p.set_stream_mode(mode)
            p.loss_eval(0.3 * x, weights="ggn")  # evict the cached pass
            out[mode] = p.loss_eval(x, weights="ggn", want_rows=True)
            assert p.stream_path() == mode
        fv, g, zg, rg, wg = out["fused"]
        As, ys = A[lo:hi], y[lo:hi]
        z = As @ x
        r, w = Lo.ggn_weights(z, ys)
        assert relerr(zg[lo:hi], z) <= 1e-13 and np.all(zg[hi:] == 0)
        assert abs(fv - Lo.f(As, ys, x)) <= 1e-13 * abs(Lo.f(As, ys, x))
        np.testing.assert_allclose(rg[lo:hi], r, rtol=1e-11, atol=1e-300)
        assert relerr(g, As.T @ r) <= 1e-12
        fv2, g2, *_ = out["two_pass"]
        assert abs(fv - fv2) <= 1e-13 * abs(fv2) and relerr(g, g2) <= 1e-12
    p.close()


@pytest.mark.parametrize("n,m", [(5, 2), (100, 50), (513, 130), (2048, 128), (3001, 257), (1024, 384)])
@pytest.mark.parametrize("wk", ["newton", "ggn"])
def test_gram(scs, n, m, wk):
    A, y, x = logistic_problem(n, m)
    p = scs.Problem(A, y, x, scs.LogisticLoss(1 / n), 0.1)  # literal labels: GGN weights are partly negative
    Lo = O.LogisticLoss(1 / n)
    z = A @ x
    w = Lo.hess_weights(z, y) if wk == "newton" else Lo.ggn_weights(z, y)[1]
    G = p.gram(x, weights=wk)
    Gref = A.T @ (w[:, None] * A)
    assert np.array_equal(G, G.T)  # exactly symmetric by construction
    scale = np.sqrt(np.outer(np.diag(A.T @ (np.abs(w)[:, None] * A)), np.diag(A.T @ (np.abs(w)[:, None] * A))))
    assert np.max(np.abs(G - Gref) / np.maximum(scale, 1e-300)) <= 1e-13
    p.close()


def test_quadform_loss(scs):
    rng = np.random.default_rng(5)
    A = rng.standard_normal((37, 37))
    y, x = rng.standard_normal(37), rng.standard_normal(37)
    p = scs.Problem(A, y, x, scs.QuadFormLoss(), 0.1)
    Lo = O.QuadFormLoss()
    fv, g, *_ = p.loss_eval(x)
    assert abs(fv - Lo.f(A, y, x)) <= 1e-13 * abs(Lo.f(A, y, x))
    assert relerr(g, Lo.grad(A, y, x)) <= 1e-13
    assert relerr(p.gram(x), Lo.hess(A, y, x)) <= 1e-15
    p.close()


@pytest.mark.parametrize("m", [1, 2, 5, 63, 64, 65, 128, 129, 200, 777, 1472])
def test_linear_solve_spd_and_indefinite(scs, m):
    rng = np.random.default_rng(m)
    B = rng.standard_normal((m + 3, m))
    M = B.T @ B + 0.5 * np.eye(m)
    b = rng.standard_normal(m)
    ctx = scs.default_context()
    d, fb = ctx.linear_solve(M, b)
    assert not fb
    assert relerr(d, np.linalg.solve(M, b)) <= 1e-11 * np.linalg.cond(M)
    if m >= 2:  # symmetric indefinite -> Cholesky flags it and the pivoted LU runs on the device
        Mi = M - 2.0 * np.diag(np.arange(m) % 2) * np.trace(M) / m
        d2, fb2 = ctx.linear_solve(Mi, b)
        assert fb2
        assert relerr(d2, np.linalg.solve(Mi, b)) <= 1e-11 * np.linalg.cond(Mi)


def test_linear_solve_legacy_sequence(scs, monkeypatch):
    """SCS_SOLVE_LEGACY=1 (read when a context is created) selects the k_panel / k_syrk_update / k_bwd_all sequence the
    look-ahead Cholesky replaced; both must solve the same systems."""
    monkeypatch.setenv("SCS_SOLVE_LEGACY", "1")
    ctx = scs.Context(0)
    monkeypatch.delenv("SCS_SOLVE_LEGACY")
    for m in (65, 200, 777):
        rng = np.random.default_rng(m)
        B = rng.standard_normal((m + 3, m))
        M = B.T @ B + 0.5 * np.eye(m)
        b = rng.standard_normal(m)
        d_old, fb = ctx.linear_solve(M, b)
        d_new, fb2 = scs.default_context().linear_solve(M, b)
        assert not fb and not fb2
        ref = np.linalg.solve(M, b)
        assert relerr(d_old, ref) <= 1e-11 * np.linalg.cond(M) and relerr(d_new, ref) <= 1e-11 * np.linalg.cond(M)
    ctx.close()


@pytest.mark.parametrize("pair", ["0", "1"])
def test_linear_solve_paired_trailing_update(scs, monkeypatch, pair):
    """SCS_SOLVE_PAIR forces / forbids the two-panels-at-a-time trailing update of the look-ahead Cholesky (by default
    chosen by size: m > 4096).  Orders with an even and an odd number of 64-blocks, ragged last blocks, one to four
    blocks; the indefinite case still reaches the LU fallback."""
    monkeypatch.setenv("SCS_SOLVE_PAIR", pair)
    ctx = scs.Context(0)
    monkeypatch.delenv("SCS_SOLVE_PAIR")
    for m in (64, 65, 128, 129, 192, 193, 256, 300, 777, 1024, 1472):
        rng = np.random.default_rng(m)
        B = rng.standard_normal((m + 3, m))
        M = B.T @ B + 0.5 * np.eye(m)
        b = rng.standard_normal(m)
        d, fb = ctx.linear_solve(M, b)
        assert not fb
        assert relerr(d, np.linalg.solve(M, b)) <= 1e-11 * np.linalg.cond(M), m
    Mi = M - 2.0 * np.diag(np.arange(m) % 2) * np.trace(M) / m
    d2, fb2 = ctx.linear_solve(Mi, b)
    assert fb2 and relerr(d2, np.linalg.solve(Mi, b)) <= 1e-11 * np.linalg.cond(Mi)
    ctx.close()


def _mk(scs, m, reg, lam, **kw):
    A = synth.make_A(8, m)
    return scs.Problem(A, np.ones(8), np.zeros(m), scs.LeastSquaresLoss(8.0), lam, **kw)


@pytest.mark.parametrize("m", [1, 50, 1025, 5000])
def test_smoothers(scs, m):
    rng = np.random.default_rng(m)
    x = rng.standard_normal(m) * 1.5
    x[:: max(m // 7, 1)] = 0.0
    lb, ub = -0.5, 0.8
    ind = None
    table = [
        (scs.PHuberSmootherL1L2(0.7), O.PHuberSmootherL1L2(0.7), "l1"),
        (scs.PHuberSmootherIndBox(lb, ub, 0.6), O.PHuberSmootherIndBox(lb, ub, 0.6), "indbox"),
        (scs.ExponentialSmootherIndBox(lb, ub, 0.6), O.ExponentialSmootherIndBox(lb, ub, 0.6), "indbox"),
        (scs.LogExpSmootherIndBox(lb, np.inf, 0.9), O.LogExpSmootherIndBox(lb, np.inf, 0.9), "indbox"),
        (scs.LogExpSmootherIndBox(lb, ub, 0.2), O.LogExpSmootherIndBox(lb, ub, 0.2), "indbox"),
        (scs.OsBaSmootherL1L2(0.7), O.OsBaSmootherL1L2(0.7), "l1"),
    ]
    for hg, ho, reg in table:
        p = _mk(scs, m, reg, 0.1, C_set=(lb, ub))
        p.configure(None, reg, hg)
        gr, hr = p.smoother_eval(x)
        with np.errstate(all="ignore"):
            go, ho_ = ho.grad(None, x), ho.hess(None, x)
        np.testing.assert_allclose(gr, go, rtol=2e-13, atol=1e-300, equal_nan=True, err_msg=type(hg).__name__)
        np.testing.assert_allclose(hr, ho_, rtol=2e-13, atol=1e-300, equal_nan=True, err_msg=type(hg).__name__)
        p.close()
    # vector bounds for the box smoother
    lbv, ubv = -np.abs(rng.standard_normal(m)) - 0.1, np.abs(rng.standard_normal(m)) + 0.1
    p = _mk(scs, m, "indbox", 0.1, C_set=(lbv, ubv))
    p.configure(None, "indbox", scs.PHuberSmootherIndBox(lbv, ubv, 0.6))
    gr, hr = p.smoother_eval(x)
    ho = O.PHuberSmootherIndBox(lbv, ubv, 0.6)
    np.testing.assert_allclose(gr, ho.grad(None, x), rtol=2e-13)
    np.testing.assert_allclose(hr, ho.hess(None, x), rtol=2e-13)
    p.close()


def _groups(m, gsz, w=1):
    starts = list(range(1, m + 1, gsz))
    return np.array([starts, [min(s + gsz - 1, m) for s in starts], [w + (i % 3) for i in range(len(starts))]])


@pytest.mark.parametrize("m,gsz", [(6, 2), (256, 64), (1000, 37), (4100, 64)])
def test_group_lasso_pieces(scs, m, gsz):
    rng = np.random.default_rng(m)
    ind = _groups(m, gsz)
    perm = rng.permutation(m) + 1
    Po, Pg = O.GroupStructure(m, perm, ind), scs.get_P(m, perm, ind)
    lam = [1e-3, 0.3]
    A = synth.make_A(8, m)
    mo = O.Problem(A, np.ones(8), np.zeros(m), O.LeastSquaresLoss(8.0), lam, P=Po)
    p = scs.Problem(A, np.ones(8), np.zeros(m), scs.LeastSquaresLoss(8.0), lam, P=Pg)
    x = rng.standard_normal(m)
    x[: gsz] *= 1e-3  # a group the prox will zero
    for hg, ho in ((scs.PHuberSmootherGL(0.05, p), O.PHuberSmootherGL(0.05, mo)),
                   (scs.OsBaSmootherGL(0.5, p), O.OsBaSmootherGL(0.5, mo))):
        p.configure(None, "gl", hg)
        gr, hr = p.smoother_eval(np.abs(x) + 0.1 if "OsBa" in type(hg).__name__ else x)
        xx = np.abs(x) + 0.1 if "OsBa" in type(hg).__name__ else x
        with np.errstate(all="ignore"):
            np.testing.assert_allclose(gr, ho.grad(Po, xx), rtol=1e-12, equal_nan=True)
            np.testing.assert_allclose(hr, ho.hess(Po, xx), rtol=1e-11, equal_nan=True)
    assert abs(p.reg_value(x) - O.get_reg(mo, x, "gl")) <= 1e-13 * abs(O.get_reg(mo, x, "gl"))
    hrv = np.abs(rng.standard_normal(m)) + 0.2
    out = p.prox(x, hrv, 0.8)
    ref = O.prox_step(mo, "gl", x, 1 / hrv, lam[0], 0.8)
    np.testing.assert_allclose(out, ref, rtol=1e-12, atol=1e-300)
    assert np.array_equal(out != 0, ref != 0)
    assert np.any(out == 0)
    p.close()


@pytest.mark.parametrize("reg", ["l1", "l2", "indbox"])
@pytest.mark.parametrize("m", [1, 50, 3000])
def test_prox_and_reg(scs, reg, m):
    rng = np.random.default_rng(m + len(reg))
    lbv, ubv = -np.abs(rng.standard_normal(m)) - 0.05, np.abs(rng.standard_normal(m)) + 0.05
    A = synth.make_A(8, m)
    mo = O.Problem(A, np.ones(8), np.zeros(m), O.LeastSquaresLoss(8.0), 0.37, C_set=(lbv, ubv))
    p = scs.Problem(A, np.ones(8), np.zeros(m), scs.LeastSquaresLoss(8.0), 0.37, C_set=(lbv, ubv))
    p.configure(None, reg, scs.PHuberSmootherL1L2(1.0))
    u = rng.standard_normal(m) * 2
    u[::5] = 0.0
    hr = np.abs(rng.standard_normal(m)) + 0.1
    out = p.prox(u, hr, 0.5)
    with np.errstate(all="ignore"):
        ref = O.prox_step(mo, reg, u, 1 / hr, 0.37, 0.5)
    assert np.array_equal(out, ref)  # same operations in the same order: bit-exact
    v = p.reg_value(u)
    vo = O.get_reg(mo, u, reg)
    assert (np.isinf(v) and np.isinf(vo)) or abs(v - vo) <= 1e-13 * max(abs(vo), 1e-300)
    inside = np.clip(u, lbv, ubv)
    assert p.reg_value(inside) == O.get_reg(mo, inside, reg) or reg != "indbox"
    p.close()


def test_synthetic_generator_bits(scs):
    """The device Philox twin must reproduce oracle/synth.py bit for bit (A), and labels / targets exactly."""
    n_total, m, row0, nl = 3000, 70, 517, 1501
    for loss, dens in ((scs.LogisticLoss(1 / n_total), 1.0), (scs.LeastSquaresLoss(float(n_total)), 1.0),
                       (scs.LogisticLoss(1 / n_total), 0.05)):
        p = scs.Problem.synthetic(n_total, m, loss, 0.1, row0=row0, n_local=nl, seed=1234, density=dens)
        Ag, yg = p.read_rows(0, nl)
        A = synth.make_A(nl, m, seed=1234, row0=row0, n_total=n_total, density=dens)
        assert np.array_equal(Ag, A)
        z = A @ synth.make_x_true(m, seed=1235)
        if isinstance(loss, scs.LogisticLoss):
            y = synth.make_labels_logistic(z, seed=1236, row0=row0)
            assert np.mean(yg != y) <= 1e-3  # a label may flip only if |u - sigma(z)| is at rounding level
            assert set(np.unique(yg)) <= {-1.0, 1.0}
        else:
            np.testing.assert_allclose(yg, synth.make_targets_ls(z, seed=1236, row0=row0), rtol=1e-12, atol=1e-14)
        p.close()


# ---- emulated-fp64 Gram on tcgen05 int8 (kernels_i8gram.cuh) ------------------------------------------------
@pytest.mark.parametrize("n,m", [(5, 2), (100, 50), (513, 130), (3001, 257), (4096, 512), (70001, 300), (131072 + 77, 640)])
@pytest.mark.parametrize("bits", [None, 30, 48])
def test_gram_i8_matches_fp64(scs, n, m, bits):
    """Forced int8/CRT path vs the fp64 oracle Gram.  The integer Gram is exact; the only error is the fixed-point
    quantisation of sqrt(w)*A (columns scaled to a common 2-norm T >= 2^bits, default 46), so entries agree to ~1e-14 of
    the diagonal scale.  n > 65536 exercises several K chunks, m not a multiple of 128/256 the ragged tiles; the
    requested bits select the moduli prefix (10..15 moduli)."""
    A, y, x = logistic_problem(n, m)
    p = scs.Problem(A, y, x, scs.LogisticLoss(1 / n, "consistent"), 0.1)
    p.set_gram_mode("i8")
    if bits:
        p.set_gram_bits(bits)
    tol = 2e-12 if bits != 30 else 2e-9
    Lo = O.LogisticLoss(1 / n, "consistent")
    z = A @ x
    for wk in ("newton", "ggn"):
        w = Lo.hess_weights(z, y) if wk == "newton" else Lo.ggn_weights(z, y)[1]
        G = p.gram(x, weights=wk)
        assert p.gram_path() == "i8"
        Gref = A.T @ (w[:, None] * A)
        assert np.array_equal(G, G.T)
        d = np.sqrt(np.diag(Gref))
        assert np.max(np.abs(G - Gref) / np.outer(d, d)) <= tol
        nmod, kept = p.gram_info()
        assert kept >= (bits or 46) and 10 <= nmod <= 15
        # shortest sufficient prefix (scs_lib.cu i8_setup): column norms are scaled to T >= 2^bits, entries < 2^50
        ldx = -(-(-(-n // 16) * 16) // 128) * 128  # rows padded to 16, then to the 128-row K block
        r = np.max(np.abs(A), axis=0) / np.linalg.norm(A, axis=0)
        T_cap = 2.0 ** 50 / r.max()
        T_req = min(2.0 ** (bits or 46), T_cap)
        T_of = [(2.0 ** ((l2 - 1) / 2) - 0.5 * np.sqrt(ldx)) * (1 - 1e-6)
                for l2 in (79.240952, 87.041852, 94.803403, 102.524503, 110.161127, 117.783179)]
        want = next((10 + k for k, T in enumerate(T_of) if T >= T_req), 15)
        near = any(abs(T / T_req - 1) < 1e-9 for T in T_of)  # on a boundary the device's norms may round the other way
        assert nmod == want or near
    with pytest.raises(scs.ScsError):
        p.set_gram_bits(44)  # planes already laid out
    # weights of both signs (literal +-1 labels, GGN: the reference's documented pair): still the int8 path — the rows of
    # the minority sign are compacted and subtracted twice inside the CRT
    p2 = scs.Problem(A, y, x, scs.LogisticLoss(1 / n, "literal"), 0.1)
    p2.set_gram_mode("i8")
    if bits:
        p2.set_gram_bits(bits)
    G2 = p2.gram(x, weights="ggn")
    w2 = O.LogisticLoss(1 / n).ggn_weights(z, y)[1]
    assert p2.gram_path() == "i8"
    crow, minor_neg = p2.gram_signed()
    nneg = int(np.sum(w2 < 0))
    assert crow == (min(nneg, _ceil(n, 16) - nneg) if nneg else -1)  # the minority sign is the compacted one
    G2ref = A.T @ (w2[:, None] * A)
    d2 = np.sqrt(np.diag(A.T @ (np.abs(w2)[:, None] * A)))  # the fixed-point image is built from |w|
    assert np.array_equal(G2, G2.T)
    assert np.max(np.abs(G2 - G2ref) / np.outer(d2, d2)) <= tol
    p.close()
    p2.close()


def _ceil(v, a):
    return -(-v // a) * a


@pytest.mark.parametrize("n,m", [(300, 40), (4096, 512), (70001, 300), (140000, 256)])
def test_gram_i8_signed_weights(scs, n, m):
    """Signed emulated Gram in both orientations: a few negative rows (negative rows compacted, G = X'X - 2 Xn'Xn), mostly
    negative rows (the non-negative rows are compacted instead, G = -X'X + 2 Xp'Xp), all rows negative (least squares with a
    negative denominator), and a mini-batch window.  Reference: A' diag(w) A in fp64 from the oracle's weights."""
    rng = np.random.default_rng(n + m)
    A = np.abs(synth.make_A(n, m, seed=11)) * np.where(rng.random((n, 1)) < 0.5, 1.0, 0.2)
    x = -np.abs(synth.make_x0(m, seed=12)) * 0.4  # z < 0 on every row: yhat < 1/2
    Lo = O.LogisticLoss(1 / n)  # literal labels
    for frac_neg_label, want_minor_neg in ((0.15, True), (0.97, False)):
        y = np.where(rng.random(n) < frac_neg_label, -1.0, 1.0)
        p = scs.Problem(A, y, x, scs.LogisticLoss(1 / n, "literal"), 0.1)
        p.set_gram_mode("i8")
        z = A @ x
        w = Lo.ggn_weights(z, y)[1]
        nneg = int(np.sum(w < 0))
        assert 0 < nneg < n
        for lo, hi in ((0, n), (n // 3 + 5, n - 7)):
            p.set_active_rows(lo, hi)
            G = p.gram(x, weights="ggn")
            assert p.gram_path() == "i8"
            ww = np.zeros(n)
            ww[lo:hi] = w[lo:hi]
            Gref = A.T @ (ww[:, None] * A)
            d = np.sqrt(np.diag(A.T @ (np.abs(ww)[:, None] * A)))
            assert np.max(np.abs(G - Gref) / np.outer(d, d)) <= 2e-12
            crow, minor_neg = p.gram_signed()
            if (lo, hi) == (0, n):
                assert minor_neg == want_minor_neg == (nneg <= _ceil(n, 16) - nneg)
                assert crow == (nneg if minor_neg else _ceil(n, 16) - nneg)
        # the DMMA kernel on the same weights agrees (it never takes sqrt(w))
        p.set_active_rows(0, n)
        p.set_gram_mode("dmma")
        Gd = p.gram(x, weights="ggn")
        assert p.gram_path() == "dmma"
        d = np.sqrt(np.diag(A.T @ (np.abs(w)[:, None] * A)))
        assert np.max(np.abs(Gd - A.T @ (w[:, None] * A)) / np.outer(d, d)) <= 2e-12
        p.close()
    # all rows negative: least squares with a negative denominator (w = 1/p < 0 everywhere)
    yls = rng.standard_normal(n)
    p = scs.Problem(A, yls, x, scs.LeastSquaresLoss(-float(n)), 0.1)
    p.set_gram_mode("i8")
    G = p.gram(x)
    assert p.gram_path() == "i8" and p.gram_signed()[1] is False
    Gref = -(A.T @ A) / n
    d = np.sqrt(np.diag(-Gref))
    assert np.max(np.abs(G - Gref) / np.outer(d, d)) <= 2e-12
    p.close()


def test_gram_i8_nonfinite_weights_propagate(scs):
    """A non-finite weight must poison the Gram like it does in fp64 arithmetic — never a silently finite matrix."""
    n, m = 2000, 64
    A, y, x = logistic_problem(n, m)
    y = y.copy()
    y[17] = np.inf  # exp(-y z) -> weight NaN on that row
    for mode in ("literal", "consistent"):
        p = scs.Problem(A, y, x, scs.LogisticLoss(1 / n, mode), 0.1)
        p.set_gram_mode("i8")
        with np.errstate(all="ignore"):
            G = p.gram(x, weights="ggn")
        assert not np.all(np.isfinite(G))
        p.close()
    # weights that are non-negative by construction take the path without any read-back: the poison flag travels on the
    # device (k_wstat_part -> k_crt)
    A, y, x = logistic_problem(n, m)
    p = scs.Problem(A, y, x, scs.LogisticLoss(1 / n, "consistent"), 0.1)
    p.set_gram_mode("i8")
    xb = x.copy()
    xb[3] = np.nan
    with np.errstate(all="ignore"):
        G = p.gram(xb, weights="ggn")
    assert p.gram_path() == "i8" and np.all(np.isnan(G))
    G = p.gram(x, weights="ggn")  # and the next call is clean again
    assert np.all(np.isfinite(G))
    p.close()


def test_gram_i8_vs_dmma_bitwise_reproducible(scs):
    n, m = 40000, 384
    A, y, x = logistic_problem(n, m)
    p = scs.Problem(A, y, x, scs.LogisticLoss(1 / n, "consistent"), 0.1)
    p.set_gram_mode("i8")
    G1 = p.gram(x, weights="ggn")
    G2 = p.gram(x * 1.0000001, weights="ggn")
    G3 = p.gram(x, weights="ggn")
    assert np.array_equal(G1, G3) and not np.array_equal(G1, G2)
    p.set_gram_mode("dmma")
    Gd = p.gram(x, weights="ggn")
    d = np.sqrt(np.diag(Gd))
    assert np.max(np.abs(G1 - Gd) / np.outer(d, d)) <= 2e-12
    p.close()


def test_gram_i8_two_cta_variant():
    """The opt-in cta_group::2 SYRK (SCS_I8_2CTA=1, read once per process) gives bit-identical integer Grams."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, numpy as np; sys.path[:0] = [%r, %r]\n"
        "import scs_b200 as S\nfrom oracle import synth\n"
        "n, m = 70001, 300\n"
        "p = S.Problem.synthetic(n, m, S.LogisticLoss(1 / n, 'consistent'), 1e-3)\n"
        "p.set_gram_mode('i8'); x = synth.make_x0(m) * 0.3\n"
        "G = p.gram(x, weights='ggn'); assert p.gram_path() == 'i8'\n"
        "np.save(sys.argv[1], G)\n" % (root, os.path.join(root, "selfconcordantsmoothoptimization.jl_b200")))
    outs = []
    for flag in ("0", "1"):
        out = os.path.join(root, "gpurun_out", f"_g2cta_{flag}.npy")
        os.makedirs(os.path.dirname(out), exist_ok=True)
        env = dict(os.environ, SCS_I8_2CTA=flag)
        r = subprocess.run([sys.executable, "-c", code, out], env=env, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(np.load(out))
        os.remove(out)
    assert np.array_equal(outs[0], outs[1])
