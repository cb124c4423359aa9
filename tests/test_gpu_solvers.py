"""End-to-end GPU parity through the reference-shaped host API (Problem / iterate / Solution).

The bar (BASELINE.json north_star): relative error <= 1e-10 on x and on the objective history, identical l1
support, same number of epochs / history entries as the oracle restatement of optim_loop!."""
import numpy as np
import pytest

import cases
from oracle import scs_oracle as O
from test_oracle_golden import load_golden
from test_oracle_reference_fixtures import A1, Y1, X01, A2, Y2, X02, XS2

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


# ---- the reference's own tests, through the GPU path (test/test_algs.jl) -----------------------------
@pytest.mark.parametrize("method", ["ProxNSCORE", "ProxGGNSCORE", "ProxLQNSCORE"])
@pytest.mark.parametrize("reg", ["l1", "l2"])
def test_algs_regression_l1_l2(scs, method, reg):  # test_algs.jl:15-52
    model = scs.Problem(A1, Y1, X01, scs.LogisticLoss(1 / 5), 1)
    sol = scs.iterate(getattr(scs, method)(), model, reg, scs.PHuberSmootherL1L2(1), verbose=0)
    assert np.allclose(model.x, np.zeros(2))
    assert sol.epochs + 1 >= 1
    assert sol.rel[-1] <= 1e-6
    assert sol.objrel[-1] <= 1e-6
    so = O.iterate(getattr(O, method)(), O.Problem(A1, Y1, X01, O.LogisticLoss(1 / 5), 1), reg, O.PHuberSmootherL1L2(1))
    assert sol.epochs == so.epochs and len(sol.obj) == len(so.obj)
    assert np.array_equal(sol.x != 0, so.x != 0)
    model.close()


@pytest.mark.parametrize("which", ["phuber", "exp"])
def test_algs_indbox(scs, which):  # test_algs.jl:94-108
    model = scs.Problem(A2, Y2, X02, scs.QuadFormLoss(), 1.0e-4, C_set=(-1.0, 1.0), sol=XS2)
    mo = O.Problem(A2, Y2, X02, O.QuadFormLoss(), 1.0e-4, C_set=(-1.0, 1.0), sol=XS2)
    if which == "phuber":
        hg, ho, al = scs.PHuberSmootherIndBox(-1.0, 1.0, 0.6), O.PHuberSmootherIndBox(-1.0, 1.0, 0.6), 0.8
    else:
        hg, ho, al = scs.ExponentialSmootherIndBox(-1.0, 1.0, 0.6), O.ExponentialSmootherIndBox(-1.0, 1.0, 0.6), 1.0
    sol = scs.iterate(scs.ProxNSCORE(), model, "indbox", hg, alpha=al, verbose=0)
    assert sol.epochs + 1 >= 1
    assert sol.rel[-1] <= 1e-3
    assert sol.objrel[-1] <= 1e-3
    so = O.iterate(O.ProxNSCORE(), mo, "indbox", ho, alpha=al)
    assert sol.epochs == so.epochs
    assert relerr(sol.x, so.x) <= TOL
    model.close()


def test_smooth_constants(scs):  # test_smooth.jl
    assert scs.PHuberSmootherL1L2(1).Mh == 2.0 and scs.PHuberSmootherL1L2(1).nu == 2.6
    assert scs.PHuberSmootherIndBox(-1.0, 1.0, 1).Mh == 2.0 and scs.PHuberSmootherIndBox(-1.0, 1.0, 1).nu == 2.6
    assert scs.OsBaSmootherL1L2(1).Mh == 2 * np.sqrt(2) and scs.OsBaSmootherL1L2(1).nu == 3.0


# ---- the five BASELINE configs at reduced n, vs the live oracle AND the committed golden vectors ------
@pytest.mark.parametrize("name", cases.CASES)
@pytest.mark.parametrize("device_loop", [False, True])
def test_config_parity(scs, name, device_loop):
    mo, modelo, reg, ho, kw = cases.build(name, O)
    so = O.iterate(mo, modelo, reg, ho, **kw)
    mg, modelg, reg, hg, kw = cases.build(name, scs)
    sg = scs.iterate(mg, modelg, reg, hg, verbose=0, device_loop=device_loop, **kw)
    assert modelg.stream_path() == "fused"  # auto: one read of A per objective + gradient
    g = load_golden(name)
    tol = TOL if name != "c2b_logreg_ggn_literal" else 1e-7  # indefinite Gram, diverging iterates: ill-conditioned
    assert sg.epochs == so.epochs == g["epochs"]
    assert len(sg.obj) == len(so.obj)
    assert relerr(sg.x, so.x) <= tol, relerr(sg.x, so.x)
    assert relerr(sg.x, g["x"]) <= tol
    assert hist_err(sg.obj, so.obj) <= tol
    assert hist_err(sg.obj, g["obj"]) <= tol
    assert hist_err(sg.fval, so.fval) <= tol
    assert hist_err(sg.rel, so.rel) <= 1e-8
    pg = [np.nan if v is None else v for v in sg.pri_res_norm]
    po = [np.nan if v is None else v for v in so.pri_res_norm]
    assert np.isnan(pg[0]) and np.isnan(po[0])
    np.testing.assert_allclose(pg[1:], po[1:], rtol=1e-7, atol=1e-14)
    if reg in ("l1", "gl"):
        assert [int(i) for i in np.nonzero(sg.x)[0]] == g["support"]
    modelg.close()


@pytest.mark.parametrize("name", cases.CASES)
def test_config_parity_two_pass_stream(scs, name):
    """The default (auto) runs the single-pass fused gradient kernel; the k_forward + k_adjoint pair must meet the
    same bar."""
    mo, modelo, reg, ho, kw = cases.build(name, O)
    so = O.iterate(mo, modelo, reg, ho, **kw)
    mg, modelg, reg, hg, kw = cases.build(name, scs)
    modelg.set_stream_mode("two_pass")
    sg = scs.iterate(mg, modelg, reg, hg, verbose=0, device_loop=True, **kw)
    assert modelg.stream_path() == "two_pass"
    tol = TOL if name != "c2b_logreg_ggn_literal" else 1e-7
    assert sg.epochs == so.epochs
    assert relerr(sg.x, so.x) <= tol
    assert hist_err(sg.obj, so.obj) <= tol
    modelg.close()


def test_first_step_matches_golden(scs):
    """One step! through scs_step (return_dx=True) against the golden first iterate of every config."""
    for name in cases.CASES:
        g = load_golden(name)
        mg, modelg, reg, hg, kw = cases.build(name, scs)
        mg.set_name()
        if kw.get("alpha") is not None:
            modelg.L = 1 / kw["alpha"]
        modelg.configure(mg, reg, hg)
        x0 = modelg.x0
        xn, dx, pri = modelg.step(x0, x0.copy(), 1, return_dx=True)
        assert relerr(xn, g["x_after_step1"]) <= TOL, name
        assert abs(pri - np.linalg.norm(xn - x0)) <= 1e-12 * max(pri, 1e-300)
        assert np.all(np.isfinite(dx))
        modelg.close()


@pytest.mark.parametrize("method", ["ProxNSCORE", "ProxGGNSCORE", "ProxLQNSCORE"])
def test_linesearch_and_noprox(scs, method):
    """ss_type=3 (Armijo, utils.jl:27-35) and use_prox=false."""
    A, y, x0 = cases.data("c3_logreg_lqn_l1")
    n = A.shape[0]
    for kwm in (dict(ss_type=3), dict(use_prox=False)):
        mo = getattr(O, method)(**kwm)
        mg = getattr(scs, method)(**kwm)
        po = O.Problem(A, y, x0, O.LogisticLoss(1 / n, "consistent"), 1e-2)
        pg = scs.Problem(A, y, x0, scs.LogisticLoss(1 / n, "consistent"), 1e-2)
        so = O.iterate(mo, po, "l1", O.PHuberSmootherL1L2(1.0), max_epoch=5, alpha=0.9)
        sg = scs.iterate(mg, pg, "l1", scs.PHuberSmootherL1L2(1.0), max_epoch=5, alpha=0.9, verbose=0)
        assert sg.epochs == so.epochs
        assert relerr(sg.x, so.x) <= TOL, (method, kwm, relerr(sg.x, so.x))
        assert hist_err(sg.obj, so.obj) <= TOL
        pg.close()


def test_error_behaviour(scs):
    A, y, x0 = cases.data("c1_readme_logreg_n")
    p = scs.Problem(A, y, x0, scs.LogisticLoss(1 / 50), 0.1)
    with pytest.raises(scs.ScsError, match="reg_name not valid"):  # prox-operators.jl:78
        scs.iterate(scs.ProxNSCORE(), p, "l3", scs.PHuberSmootherL1L2(1.0), verbose=0)
    with pytest.raises(scs.ScsError, match="exactly two entries"):  # regularizers.jl:21-23
        scs.iterate(scs.ProxNSCORE(), p, "gl", scs.PHuberSmootherL1L2(1.0), verbose=0)
    with pytest.raises(scs.ScsError, match="ss_type in"):  # prox-N-SCORE.jl:89
        scs.iterate(scs.ProxNSCORE(ss_type=7), p, "l1", scs.PHuberSmootherL1L2(1.0), verbose=0)
    with pytest.raises(scs.ScsError, match="positive"):  # smoothing.jl:15
        scs.iterate(scs.ProxNSCORE(), p, "l1", scs.PHuberSmootherL1L2(-1.0), verbose=0)
    with pytest.raises(scs.UnsupportedError):  # ss_type=2 is broken upstream for N/GGN (quirk 5)
        scs.iterate(scs.ProxGGNSCORE(ss_type=2), p, "l1", scs.PHuberSmootherL1L2(1.0), max_epoch=3, verbose=0)
    with pytest.raises(scs.ScsError, match="bounds"):  # prox-reg-utils.jl:154
        p.C_set = (-1.0, 1.0)
        scs.iterate(scs.ProxNSCORE(), p, "indbox", scs.PHuberSmootherIndBox(np.zeros(3), np.ones(3), 1.0), verbose=0)
    with pytest.raises(scs.UnsupportedError):  # metric callbacks need x on the host every epoch: host loop only
        scs.iterate(scs.ProxNSCORE(), p, "l1", scs.PHuberSmootherL1L2(1.0), metrics={"acc": lambda m, x: 0.0}, verbose=0,
                    device_loop=True)
    with pytest.raises(scs.ScsError):  # a mini-batch window must stay inside the shard
        p.set_active_rows(0, A.shape[0] + 1)
    p.close()


@pytest.mark.parametrize("name", ["c2_logreg_ggn_l1", "c4_ls_ggn_gl", "c5_ls_n_indbox", "c1_readme_logreg_n"])
def test_config_parity_i8_gram(scs, name):
    """Same 1e-10 bar with the Gram built by the tcgen05 int8 / CRT kernel (forced; auto picks it only for big shards)."""
    mo, modelo, reg, ho, kw = cases.build(name, O)
    so = O.iterate(mo, modelo, reg, ho, **kw)
    mg, modelg, reg, hg, kw = cases.build(name, scs)
    modelg.set_gram_mode("i8")
    sg = scs.iterate(mg, modelg, reg, hg, verbose=0, device_loop=True, **kw)
    assert modelg.gram_path() == "i8"
    assert sg.epochs == so.epochs
    assert relerr(sg.x, so.x) <= TOL, relerr(sg.x, so.x)
    assert hist_err(sg.obj, so.obj) <= TOL
    if reg in ("l1", "gl"):
        assert np.array_equal(sg.x != 0, so.x != 0)
    modelg.close()


# ---- ProxGGNSCORE underdetermined branch: n + 1 <= m (prox-GGN-SCORE.jl:124-127) -------------------------------
@pytest.mark.parametrize("n,m,loss", [(10, 50, "logistic"), (40, 100, "consistent"), (63, 64, "ls"), (1, 2, "logistic"),
                                      (200, 777, "consistent"), (129, 300, "ls")])
@pytest.mark.parametrize("ss_type", [1, 3])
def test_ggn_wide_branch(scs, n, m, loss, ss_type):
    from oracle import synth
    A = synth.make_A(n, m, seed=11)
    x0 = synth.make_x0(m, seed=12) * 0.3
    if loss == "ls":
        y = synth.make_targets_ls(A @ synth.make_x_true(m, seed=13), seed=14)
        Lo, Lg = O.LeastSquaresLoss(float(n)), scs.LeastSquaresLoss(float(n))
    else:
        y = synth.make_labels_logistic(A @ synth.make_x_true(m, seed=13, frac=0.3), seed=14)
        mode = "consistent" if loss == "consistent" else "literal"
        Lo, Lg = O.LogisticLoss(1 / n, mode), scs.LogisticLoss(1 / n, mode)
    so = O.iterate(O.ProxGGNSCORE(ss_type=ss_type), O.Problem(A, y, x0, Lo, 1e-2), "l1", O.PHuberSmootherL1L2(1.0),
                   max_epoch=5, alpha=0.9)
    for device_loop in (False, True):
        pg = scs.Problem(A, y, x0, Lg, 1e-2)
        sg = scs.iterate(scs.ProxGGNSCORE(ss_type=ss_type), pg, "l1", scs.PHuberSmootherL1L2(1.0), max_epoch=5, alpha=0.9,
                         verbose=0, device_loop=device_loop)
        assert sg.epochs == so.epochs and len(sg.obj) == len(so.obj)
        assert relerr(sg.x, so.x) <= 1e-9, relerr(sg.x, so.x)  # a general n x n LU sits in the middle of every step
        assert hist_err(sg.obj, so.obj) <= 1e-9
        assert np.array_equal(sg.x != 0, so.x != 0)
        pg.close()


def _sparse_problem(n, m, density, seed, loss="logistic"):
    import scipy.sparse as sp
    rng = np.random.default_rng(seed)
    A = sp.random(n, m, density=density, format="csc", random_state=rng, data_rvs=rng.standard_normal)
    A = A + sp.csc_matrix((np.full(m, 0.3), (rng.integers(0, n, m), np.arange(m))), shape=(n, m))  # no empty column
    Ad = A.toarray()
    xt = np.where(rng.random(m) < 0.3, rng.standard_normal(m) * 2, 0.0)
    if loss == "ls":
        y = Ad @ xt + 0.1 * rng.standard_normal(n)
    else:
        y = np.where(rng.random(n) < 1 / (1 + np.exp(-(Ad @ xt))), 1.0, -1.0)
    return sp.csc_matrix(A), Ad, y, rng.standard_normal(m) * 0.3


def test_sparse_csc_input(scs):
    """README.md:100-125: A = sprandn(n, m, 0.01).  A scipy CSC / CSR matrix goes over the wire as colptr / rowval /
    nzval (scs_problem_create_csc); expanded on the device (storage="dense") it must give bit for bit what the dense
    upload gives, kept sparse (storage="sparse": k_sp_forward / k_sp_adjoint / k_sp_gram) the same 1e-10 bar holds."""
    import scipy.sparse as sp
    A, y, x0 = cases.data("c1_readme_logreg_n")  # 1 %-dense look-alike of the README example
    rng = np.random.default_rng(3)
    A = A + (rng.random(A.shape) < 0.05) * rng.standard_normal(A.shape)  # a few more entries per column
    so = O.iterate(O.ProxNSCORE(), O.Problem(A, y, x0, O.LogisticLoss(1 / 50), 1e-1), "l1", O.PHuberSmootherL1L2(1.0),
                   max_epoch=30, x_tol=1e-6, f_tol=1e-6)
    sols = []
    for mat, storage in ((A, "auto"), (sp.csc_matrix(A), "dense"), (sp.csr_matrix(A), "dense"), (sp.csc_matrix(A), "sparse")):
        p = scs.Problem(mat, y, x0, scs.LogisticLoss(1 / 50), 1e-1, storage=storage)
        assert p.is_sparse()[0] == (storage == "sparse")
        sols.append(scs.iterate(scs.ProxNSCORE(), p, "l1", scs.PHuberSmootherL1L2(1.0), max_epoch=30, x_tol=1e-6,
                                f_tol=1e-6, verbose=0))
        p.close()
    for sg in sols:
        assert sg.epochs == so.epochs and relerr(sg.x, so.x) <= TOL and hist_err(sg.obj, so.obj) <= TOL
        assert np.array_equal(sg.x != 0, so.x != 0)
    assert np.array_equal(sols[0].x, sols[1].x) and np.array_equal(sols[0].x, sols[2].x)  # same resident matrix
    with pytest.raises(scs.ScsError):  # malformed structure is an argument error, not a crash
        bad = sp.csc_matrix(A)
        bad.indices = bad.indices.copy()
        bad.indices[0] = A.shape[0] + 7
        scs.Problem(bad, y, x0, scs.LogisticLoss(1 / 50), 1e-1)


@pytest.mark.parametrize("method,reg,loss", [("ProxNSCORE", "l1", "logistic"), ("ProxGGNSCORE", "l1", "consistent"),
                                             ("ProxLQNSCORE", "l2", "logistic"), ("ProxGGNSCORE", "indbox", "ls")])
@pytest.mark.parametrize("batch_size", [None, 1300])
def test_sparse_compute_path(scs, method, reg, loss, batch_size):
    """The shard stays sparse on the device (1 % density): every pass and the Gram run on the CSR / CSC copies; full
    batch and mini-batches, host loop and in-library loop, against the oracle on the densified matrix."""
    n, m = 6000, 400
    As, Ad, y, x0 = _sparse_problem(n, m, 0.01, seed=21, loss="ls" if loss == "ls" else "logistic")
    if loss == "ls":
        Lo, Lg = O.LeastSquaresLoss(float(n)), scs.LeastSquaresLoss(float(n))
    else:
        mode = "consistent" if loss == "consistent" else "literal"
        Lo, Lg = O.LogisticLoss(1 / n, mode), scs.LogisticLoss(1 / n, mode)
    kwp = dict(C_set=(-0.5, 0.5)) if reg == "indbox" else {}
    ho = O.PHuberSmootherIndBox(-0.5, 0.5, 0.6) if reg == "indbox" else O.PHuberSmootherL1L2(1.0)
    kw = dict(max_epoch=6, alpha=0.9, batch_size=batch_size)
    so = O.iterate(getattr(O, method)(), O.Problem(Ad, y, x0, Lo, 1e-3, **kwp), reg, ho, **kw)
    for device_loop in (False, True):
        pg = scs.Problem(As, y, x0, Lg, 1e-3, storage="sparse", **kwp)
        assert pg.is_sparse() == (True, As.nnz)
        hg = scs.PHuberSmootherIndBox(-0.5, 0.5, 0.6) if reg == "indbox" else scs.PHuberSmootherL1L2(1.0)
        sg = scs.iterate(getattr(scs, method)(), pg, reg, hg, verbose=0, device_loop=device_loop, shuffle_batch=False, **kw)
        assert sg.epochs == so.epochs and len(sg.obj) == len(so.obj)
        assert relerr(sg.x, so.x) <= TOL, relerr(sg.x, so.x)
        assert hist_err(sg.obj, so.obj) <= TOL
        if method != "ProxLQNSCORE":
            assert pg.gram_path() == "sparse"
        pg.close()


@pytest.mark.parametrize("n,m,loss,batch_size", [(60, 200, "consistent", None), (150, 400, "ls", None),
                                                 (3000, 300, "literal", 128)])
def test_sparse_shard_takes_the_ggn_wide_branch(scs, n, m, loss, batch_size):
    """ProxGGNSCORE underdetermined branch (rows + 1 <= m, prox-GGN-SCORE.jl:124-127) on a shard that is resident in
    sparse form: the batch is expanded from the CSR copy into the replicated dense buffer (k_wide_gather_csr), the rest is
    the dense algorithm.  Whole problem wide, and mini-batches smaller than m of a tall problem."""
    As, Ad, y, x0 = _sparse_problem(n, m, 0.05, seed=33, loss="ls" if loss == "ls" else "logistic")
    if loss == "ls":
        Lo, Lg = O.LeastSquaresLoss(float(n)), scs.LeastSquaresLoss(float(n))
    else:
        Lo, Lg = O.LogisticLoss(1 / n, loss), scs.LogisticLoss(1 / n, loss)
    kw = dict(max_epoch=4, alpha=0.9, batch_size=batch_size)
    so = O.iterate(O.ProxGGNSCORE(), O.Problem(Ad, y, x0, Lo, 1e-2), "l1", O.PHuberSmootherL1L2(1.0), **kw)
    for device_loop in (False, True):
        pg = scs.Problem(As, y, x0, Lg, 1e-2, storage="sparse")
        assert pg.is_sparse()[0]
        sg = scs.iterate(scs.ProxGGNSCORE(), pg, "l1", scs.PHuberSmootherL1L2(1.0), verbose=0, device_loop=device_loop,
                         shuffle_batch=False, **kw)
        assert sg.epochs == so.epochs and len(sg.obj) == len(so.obj)
        assert relerr(sg.x, so.x) <= 1e-9, relerr(sg.x, so.x)
        assert hist_err(sg.obj, so.obj) <= 1e-9
        pg.close()


def test_device_side_sparsify_of_a_synthetic_shard(scs):
    """Problem.synthetic(density=..., storage="sparse"): the generator fills the dense layout in HBM, scs_problem_sparsify
    drops the zeros on the device (CSR + CSC copies, k_nnz_* / k_scan_* / k_fill_*).  The sparse shard must give the same
    passes, Gram and iterates as the dense shard generated from the same seed."""
    n, m = 20_000, 300
    L = scs.LogisticLoss(1 / n, "consistent")
    x0 = np.random.default_rng(5).standard_normal(m) * 0.3
    pd = scs.Problem.synthetic(n, m, L, 1e-3, x0=x0, density=0.02, storage="dense")
    ps = scs.Problem.synthetic(n, m, L, 1e-3, x0=x0, density=0.02, storage="sparse")
    sp_flag, nnz = ps.is_sparse()
    Ad, yd = pd.read_rows(0, n)
    assert sp_flag and nnz == int(np.count_nonzero(Ad)) and 0.015 * n * m < nnz < 0.025 * n * m
    fd, gd, zd, rd, wd = pd.loss_eval(x0, weights="ggn", want_rows=True)
    fs, gs, zs, rs, ws = ps.loss_eval(x0, weights="ggn", want_rows=True)
    assert abs(fs - fd) <= 1e-13 * abs(fd) and relerr(gs, gd) <= 1e-12
    np.testing.assert_allclose(zs, zd, rtol=1e-12, atol=1e-15)
    Gd, Gs = pd.gram(x0, weights="ggn"), ps.gram(x0, weights="ggn")
    assert ps.gram_path() == "sparse"
    d = np.sqrt(np.diag(Gd))
    assert np.max(np.abs(Gs - Gd) / np.outer(d, d)) <= 1e-12
    sd = scs.iterate(scs.ProxGGNSCORE(), pd, "l1", scs.PHuberSmootherL1L2(1.0), max_epoch=5, alpha=1, verbose=0, device_loop=True)
    ss = scs.iterate(scs.ProxGGNSCORE(), ps, "l1", scs.PHuberSmootherL1L2(1.0), max_epoch=5, alpha=1, verbose=0, device_loop=True)
    assert relerr(ss.x, sd.x) <= TOL and hist_err(ss.obj, sd.obj) <= TOL and np.array_equal(ss.x != 0, sd.x != 0)
    with pytest.raises(scs.ScsError):
        scs._capi.check(scs._capi.lib().scs_problem_sparsify(pd._h))  # only right after creation
    pd.close()
    ps.close()


def test_sparse_components_and_auto_storage(scs):
    import scipy.sparse as sp
    n, m = 3001, 257
    As, Ad, y, x = _sparse_problem(n, m, 0.02, seed=5)
    p = scs.Problem(As, y, x, scs.LogisticLoss(1 / n, "consistent"), 0.1)  # 2 % stored: auto keeps it sparse
    assert p.is_sparse()[0]
    Lo = O.LogisticLoss(1 / n, "consistent")
    z = Ad @ x
    for wk in ("newton", "ggn"):
        fv, g, zg, rg, wg = p.loss_eval(x, weights=wk, want_rows=True)
        r, w = (Lo.grad_weights(z, y), Lo.hess_weights(z, y)) if wk == "newton" else Lo.ggn_weights(z, y)
        assert relerr(zg, z) <= 1e-13 and abs(fv - Lo.f(Ad, y, x)) <= 1e-13 * abs(Lo.f(Ad, y, x))
        np.testing.assert_allclose(rg, r, rtol=1e-11, atol=1e-300)
        assert relerr(g, Ad.T @ r) <= 1e-12
        G = p.gram(x, weights=wk)
        Gref = Ad.T @ (w[:, None] * Ad)
        assert np.array_equal(G, G.T)
        d = np.sqrt(np.diag(Gref))
        assert np.max(np.abs(G - Gref) / np.outer(d, d)) <= 1e-13
        assert np.array_equal(G, p.gram(x, weights=wk))  # bit-reproducible
    p.close()
    dense_ish = sp.csc_matrix(np.where(np.random.default_rng(1).random((300, 40)) < 0.5, 1.0, 0.0))
    q = scs.Problem(dense_ish, np.ones(300), np.zeros(40), scs.LeastSquaresLoss(300.0), 0.1)  # 50 % stored: expanded
    assert not q.is_sparse()[0]
    q.close()
