"""Mini-batch mode on the GPU (iterate.jl:122-145,204-255) against the oracle's restatement of the same loop:
the objective history is taken over all rows, step! over one batch at a time.  Same bar as the full-batch path:
1e-10 relative on x and the objective history, identical support, same epoch / history counts."""
import numpy as np
import pytest

import cases
from oracle import scs_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-10


def relerr(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def hist_err(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    assert a.shape == b.shape
    fin = np.isfinite(b)
    assert np.array_equal(np.isfinite(a), fin)
    return float(np.max(np.abs(a[fin] - b[fin]) / np.maximum(np.abs(b[fin]), 1e-300))) if fin.any() else 0.0


def run_pair(scs, name, device_loop, stream="auto", **bkw):
    mo, modelo, reg, ho, kw = cases.build(name, O)
    kw = dict(kw, max_epoch=min(kw["max_epoch"], 5))
    okw = dict(bkw)
    okw.pop("shuffle_batch", None)
    so = O.iterate(mo, modelo, reg, ho, shuffle_batch=bkw.get("shuffle_batch", False), **okw, **kw)
    mg, modelg, reg, hg, _ = cases.build(name, scs)
    modelg.set_stream_mode(stream)
    sg = scs.iterate(mg, modelg, reg, hg, verbose=0, device_loop=device_loop,
                     **dict(bkw, shuffle_batch=bkw.get("shuffle_batch", False)), **kw)
    return so, sg, modelg, reg


# batch sizes that do not divide n, are not multiples of 16/128, and leave a short last batch
@pytest.mark.parametrize("name,bs", [("c3_logreg_lqn_l1", 1000), ("c2_logreg_ggn_l1", 1400), ("c5_ls_n_indbox", 700),
                                     ("c4_ls_ggn_gl", 1024), ("c1_readme_logreg_n", 64), ("c3b_logreg_lqn_l2_bb", 2048)])
@pytest.mark.parametrize("device_loop", [False, True])
def test_minibatch_parity(scs, name, bs, device_loop):
    so, sg, modelg, reg = run_pair(scs, name, device_loop, batch_size=bs)
    assert sg.epochs == so.epochs and len(sg.obj) == len(so.obj)
    assert relerr(sg.x, so.x) <= TOL, relerr(sg.x, so.x)
    assert hist_err(sg.obj, so.obj) <= TOL
    assert hist_err(sg.fval, so.fval) <= TOL
    if reg in ("l1", "gl"):
        assert np.array_equal(sg.x != 0, so.x != 0)
    modelg.close()


@pytest.mark.parametrize("stream", ["two_pass", "fused"])
def test_minibatch_shuffled_and_truncated(scs, stream):
    """shuffle_batch (explicit permutation, the same for both sides) + local_max_iter (only the first batches step)."""
    name = "c3_logreg_lqn_l1"
    n = cases.data(name)[0].shape[0]
    perm = np.random.default_rng(5).permutation(n)
    so, sg, modelg, reg = run_pair(scs, name, True, stream=stream, batch_size=777, shuffle_batch=True, perm=perm,
                                   local_max_iter=3)
    assert modelg.stream_path() == stream
    assert sg.epochs == so.epochs and len(sg.obj) == len(so.obj)
    assert relerr(sg.x, so.x) <= TOL
    assert hist_err(sg.obj, so.obj) <= TOL
    assert np.array_equal(sg.x != 0, so.x != 0)
    modelg.close()


def test_slice_samples(scs):
    """slice_samples=true (utils.jl:14-16): the loader subset is 1:iend with iend = 1 (iterate.jl:122-145), so every
    step! sees the FIRST row only — for ProxNSCORE a 1-row Newton system per step."""
    from test_oracle_reference_fixtures import A1, Y1, X01
    po = O.Problem(A1, Y1, X01, O.LogisticLoss(1 / 5), 1)
    pg = scs.Problem(A1, Y1, X01, scs.LogisticLoss(1 / 5), 1)
    for method in ("ProxNSCORE", "ProxLQNSCORE"):
        so = O.iterate(getattr(O, method)(), po, "l1", O.PHuberSmootherL1L2(1), slice_samples=True, max_epoch=4)
        sg = scs.iterate(getattr(scs, method)(), pg, "l1", scs.PHuberSmootherL1L2(1), slice_samples=True, max_epoch=4,
                         verbose=0)
        assert sg.epochs == so.epochs and len(sg.obj) == len(so.obj)
        assert relerr(sg.x, so.x) <= 1e-9 or np.linalg.norm(so.x) < 1e-12
        assert hist_err(sg.obj, so.obj) <= 1e-9
    pg.close()


def test_small_batches_take_the_ggn_wide_branch(scs):
    """Mini-batches of fewer rows than variables: every step! is the underdetermined GGN branch (n_batch+1 <= m)."""
    so, sg, modelg, reg = run_pair(scs, "c2_logreg_ggn_l1", True, batch_size=200, local_max_iter=4)
    assert sg.epochs == so.epochs and len(sg.obj) == len(so.obj)
    assert relerr(sg.x, so.x) <= 1e-9, relerr(sg.x, so.x)
    assert hist_err(sg.obj, so.obj) <= 1e-9
    modelg.close()


def test_active_rows_components(scs):
    """scs_set_active_rows on the component entry points: loss, gradient, row outputs and both Gram kernels see
    exactly the rows of the window (unaligned bounds), and nothing leaks between consecutive windows."""
    from oracle import synth
    n, m = 5000, 300
    A = synth.make_A(n, m, seed=3)
    y = synth.make_labels_logistic(A @ synth.make_x_true(m, seed=4, frac=0.3), seed=5)
    x = synth.make_x0(m, seed=6) * 0.5
    Lo, Lg = O.LogisticLoss(1 / n, "consistent"), scs.LogisticLoss(1 / n, "consistent")
    p = scs.Problem(A, y, x, Lg, 0.1)
    for stream in ("two_pass", "fused"):
        p.set_stream_mode(stream)
        for lo, hi in [(0, n), (1, 4999), (130, 131), (1000, 3333), (4097, 5000), (0, 17), (2500, 2500)]:
            p.set_active_rows(lo, hi)
            fv, g, zg, rg, wg = p.loss_eval(x, weights="ggn", want_rows=True)
            As, ys = A[lo:hi], y[lo:hi]
            z = As @ x
            r, w = Lo.ggn_weights(z, ys)
            fref = Lo.f(As, ys, x) if hi > lo else 0.0
            assert abs(fv - fref) <= 1e-13 * max(abs(fref), 1e-300)
            assert relerr(g, As.T @ r) <= 1e-12 or hi == lo
            # the kernels sweep the 128-row-aligned superset of the window and must mask what is not in it (rows
            # further out are never read, whatever an earlier window left there)
            alo, ahi = lo // 128 * 128, min(-(-hi // 128) * 128, n)
            if hi == lo:
                alo = ahi = lo  # an empty window sweeps nothing
            assert np.all(rg[alo:lo] == 0) and np.all(rg[hi:ahi] == 0) and np.all(wg[alo:lo] == 0) and np.all(wg[hi:ahi] == 0)
            np.testing.assert_allclose(rg[lo:hi], r, rtol=1e-11, atol=1e-300)
            for mode in ("dmma", "i8"):
                p.set_gram_mode(mode)
                G = p.gram(x, weights="ggn")
                Gref = As.T @ (w[:, None] * As)
                d = np.sqrt(np.maximum(np.diag(Gref), 1e-300))
                # the int8 path quantises against the column maxima of the WHOLE shard (40 bits below them): relative
                # to a short window's diagonal that is 2^-40 * O(10) / sqrt(rows)
                tol = 2e-12 if mode == "dmma" else max(2e-12, 5e-11 / np.sqrt(max(hi - lo, 1)))
                assert np.max(np.abs(G - Gref) / np.outer(d, d)) <= tol or hi - lo < 2
    p.close()


@pytest.mark.parametrize("device_loop", [False, True])
@pytest.mark.parametrize("batch_size", [None, 1500])
def test_held_out_data_history(scs, device_loop, batch_size):
    """Problem(...; Atest, ytest): Solution.fvaltest holds f(Atest, ytest, x) for every recorded state
    (iterate.jl:169-176, utils.jl:55-57), full batch and mini-batch, host loop and in-library loop."""
    name = "c3_logreg_lqn_l1"
    A, y, x0 = cases.data(name)
    At, yt = A[:777] * 1.5, y[:777]
    mo, modelo, reg, ho, kw = cases.build(name, O, Atest=At, ytest=yt)
    kw = dict(kw, max_epoch=6)
    so = O.iterate(mo, modelo, reg, ho, batch_size=batch_size, **kw)
    mg, modelg, reg, hg, _ = cases.build(name, scs, Atest=At, ytest=yt)
    sg = scs.iterate(mg, modelg, reg, hg, verbose=0, device_loop=device_loop, batch_size=batch_size, shuffle_batch=False,
                     **kw)
    assert len(so.fvaltest) == len(so.obj) and len(sg.fvaltest) == len(sg.obj) == len(so.obj)
    assert hist_err(sg.fvaltest, so.fvaltest) <= TOL
    assert hist_err(sg.obj, so.obj) <= TOL and relerr(sg.x, so.x) <= TOL
    modelg.close()
