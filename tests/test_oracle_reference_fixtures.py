"""Pins the oracle to every assertion of the reference's own tests: test/test_algs.jl and test/test_smooth.jl
(fixtures are the literals of those files).  These are the only known answers the reference ships for this path."""
import numpy as np
import pytest

from oracle import scs_oracle as O

# test/test_algs.jl:2-11
A1 = np.array([[-0.560501, 0.0], [0.0, 1.85278], [-0.0192918, -0.827763], [0.128064, 0.110096], [0.0, -0.251176]])
Y1 = np.array([-1, -1, -1, 1, -1.0])
X01 = np.array([0.5908446386657102, 0.7667970365022592])
# test/test_algs.jl:82-90
A2 = np.array([[1.53976, 0.201833, 0.433995, 0.156497, 0.180124], [0.201833, 2.37257, -0.0594941, -0.671533, 0.0739676],
               [0.433995, -0.0594941, 3.15025, 0.808797, 0.954656], [0.156497, -0.671533, 0.808797, 2.74361, 0.5621],
               [0.180124, 0.0739676, 0.954656, 0.5621, 1.76141]])
Y2 = np.array([0.8673472019512456, -0.9017438158568171, -0.4944787535042339, -0.9029142938652416, 0.8644013132535154])
X02 = np.array([-2.07754990163271, -2.311005948690538, -0.25157276401631606, -0.8858618022602884, 1.3116613046047525])
XS2 = np.array([-0.7139006111210786, 0.642716661564418, 0.3684773651494535, 0.5890487798472874, -0.8324174178513779])


@pytest.mark.parametrize("method", [O.ProxNSCORE, O.ProxGGNSCORE, O.ProxLQNSCORE])
@pytest.mark.parametrize("reg", ["l1", "l2"])
def test_algs_regression_l1_l2(method, reg):  # test_algs.jl:15-52 (TOL = 1e-6)
    model = O.Problem(A1, Y1, X01, O.LogisticLoss(1 / 5), 1)
    sol = O.iterate(method(), model, reg, O.PHuberSmootherL1L2(1))
    assert np.allclose(model.x, np.zeros(2))
    assert sol.epochs + 1 >= 1
    assert sol.rel[-1] <= 1e-6
    assert sol.objrel[-1] <= 1e-6


def test_algs_indbox_phuber():  # test_algs.jl:94-100 (TOL = 1e-3)
    model = O.Problem(A2, Y2, X02, O.QuadFormLoss(), 1.0e-4, C_set=(-1.0, 1.0), sol=XS2)
    sol = O.iterate(O.ProxNSCORE(), model, "indbox", O.PHuberSmootherIndBox(-1.0, 1.0, 0.6), alpha=0.8)
    assert sol.epochs + 1 >= 1
    assert sol.rel[-1] <= 1e-3
    assert sol.objrel[-1] <= 1e-3


def test_algs_indbox_exp():  # test_algs.jl:102-108
    model = O.Problem(A2, Y2, X02, O.QuadFormLoss(), 1.0e-4, C_set=(-1.0, 1.0), sol=XS2)
    sol = O.iterate(O.ProxNSCORE(), model, "indbox", O.ExponentialSmootherIndBox(-1.0, 1.0, 0.6), alpha=1.0)
    assert sol.rel[-1] <= 1e-3
    assert sol.objrel[-1] <= 1e-3


def test_smooth_constants():  # test_smooth.jl:5-21
    h = O.PHuberSmootherL1L2(1)
    assert h.Mh == 2.0 and h.nu == 2.6
    h = O.PHuberSmootherIndBox(-1.0, 1.0, 1)
    assert h.Mh == 2.0 and h.nu == 2.6
    h = O.OsBaSmootherL1L2(1)
    assert h.Mh == 2 * np.sqrt(2) and h.nu == 3.0


def test_smoother_derivatives_against_mpmath():
    """Independent pin of the elementwise formulas: grad/hess must be the 1st/2nd derivative of the reference's
    own value functions (phuber-smooth.jl:28-30, ostrovskii-bach-smooth.jl:28-30), evaluated in 50-digit mpmath."""
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 50
    mu = 0.7

    def ph(x):
        return (mu**2 - mu * mp.sqrt(mu**2 + x**2) + x**2) / mp.sqrt(mu**2 + x**2)

    def osba(x):
        s = mp.sqrt(mu**2 + 4 * x**2)
        return s / 2 - mp.mpf(mu) / 2 + mu * mp.log((2 * x - s + mu) / x) / 2 - mp.log(2) * mu + mu * mp.log((s - mu + 2 * x) / x) / 2

    xs = np.array([-3.1, -0.4, 0.05, 0.9, 2.5])
    g, h = O.huber_grad(xs, mu), O.huber_hess(xs, mu)
    og, oh = O.osba_smooth_grad_l1(xs, mu), O.osba_smooth_hess_l1(xs, mu)
    for i, x in enumerate(xs):
        assert abs(float(mp.diff(ph, mp.mpf(x))) - g[i]) < 1e-12
        assert abs(float(mp.diff(ph, mp.mpf(x), 2)) - h[i]) < 1e-10
        if x > 0:  # the OsBa value has logs of ratios that are real for x > 0
            assert abs(float(mp.diff(osba, mp.mpf(x))) - og[i]) < 1e-10
            assert abs(float(mp.diff(osba, mp.mpf(x), 2)) - oh[i]) < 1e-9


def test_loss_derivatives_finite_difference():
    """r, w of the built-in losses are d f/dz and d2 f/dz2 of the README closures; the GGN pair reproduces J'res."""
    rng = np.random.default_rng(0)
    A = rng.standard_normal((40, 6))
    y = np.where(rng.random(40) < 0.5, 1.0, -1.0)
    x = rng.standard_normal(6) * 0.3
    for L in (O.LogisticLoss(1 / 40), O.LeastSquaresLoss(40.0)):
        z = A @ x
        g = A.T @ L.grad_weights(z, y)
        H = A.T @ (L.hess_weights(z, y)[:, None] * A)
        eps = 1e-6
        for j in range(6):
            e = np.zeros(6)
            e[j] = eps
            fd = (L.f(A, y, x + e) - L.f(A, y, x - e)) / (2 * eps)
            assert abs(fd - g[j]) < 1e-8
            gp = A.T @ L.grad_weights(A @ (x + e), y)
            gm = A.T @ L.grad_weights(A @ (x - e), y)
            assert np.allclose((gp - gm) / (2 * eps), H[:, j], atol=1e-7)
    # consistent-label cross-entropy == logistic loss, so the GGN gradient equals the Newton gradient
    Lc = O.LogisticLoss(1 / 40, "consistent")
    z = A @ x
    r_ggn, w_ggn = Lc.ggn_weights(z, y)
    assert np.allclose(A.T @ r_ggn, A.T @ Lc.grad_weights(z, y), atol=1e-14)
    assert np.all(w_ggn >= 0)
    # literal +-1 labels: some Gram weights are negative
    assert np.any(O.LogisticLoss(1 / 40).ggn_weights(z, y)[1] < 0)


def test_ggn_wide_branch_matches_tall_formula_when_square_system_is_consistent():
    """prox-GGN-SCORE.jl:124-127 (n+1 <= m): the oracle restates it; sanity: it returns a finite step."""
    rng = np.random.default_rng(1)
    A = rng.standard_normal((3, 8))
    y = np.array([1.0, -1.0, 1.0])
    model = O.Problem(A, y, rng.standard_normal(8), O.LeastSquaresLoss(3.0), 0.1)
    sol = O.iterate(O.ProxGGNSCORE(), model, "l1", O.PHuberSmootherL1L2(1.0), max_epoch=3)
    assert np.all(np.isfinite(sol.x))


def test_history_quirks():
    """iterate.jl:202,219-231: at epoch == max_epoch the pre-step stats are pushed twice; pri_res_norm[1] is nothing."""
    model = O.Problem(A1, Y1, X01, O.LogisticLoss(1 / 5), 1)
    sol = O.iterate(O.ProxNSCORE(), model, "l1", O.PHuberSmootherL1L2(1), max_epoch=2)
    assert sol.pri_res_norm[0] is None
    assert len(sol.obj) == 3 and sol.obj[1] == sol.obj[2]
    assert sol.epochs == 2
