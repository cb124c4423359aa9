"""CPU oracle for the proximal-SCORE hot path (TEST INFRASTRUCTURE, NOT PRODUCT).

A numpy/fp64 restatement of SelfConcordantSmoothOptimization.jl v0.1.8, following the
reference file by file.  It exists to check the CUDA path; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs
may import it.  Nothing under ``selfconcordantsmoothoptimization.jl_b200/`` imports it.

PARITY PINNING: the reference cannot be executed here (no ``julia`` binary in this image, no
network).  The oracle is pinned against every assertion of the reference's own tests
(``test/test_algs.jl``, ``test/test_smooth.jl``) in ``tests/test_oracle_reference_fixtures.py``.
Those tests assert only 1e-6 / 1e-3 end-to-end tolerances; the reference ships no golden
iterates.  At the 1e-10 level parity is therefore **unpinned**: it rests on this restatement.

ForwardDiff (the reference's default derivative engine, not vendored) is replaced by the
analytic derivatives ForwardDiff's dual-number rules produce for the README / test closures;
LinearAlgebra's ``\\`` (LU) and ``qr(...) \\`` are replaced by ``numpy.linalg.solve`` (LAPACK LU).

Citations are ``file:line`` relative to ``/root/reference``.  Convention: n = rows (samples),
m = columns (variables).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

EPS = float(np.finfo(np.float64).eps)  # Julia eps()
L_INF_CACHE = -1e32  # src/utils/prox-reg-utils.jl:6
U_INF_CACHE = 1e32  # src/utils/prox-reg-utils.jl:7


# --------------------------------------------------------------------------------------
# Built-in losses.  The reference has none: the formulas are the README / test closures.
# --------------------------------------------------------------------------------------
class LogisticLoss:
    """f(A,y,x) = scale*sum(log(1+exp(-y.*(A*x))))          README.md:113, test/test_algs.jl:9
    f(y,yhat)  = -scale*sum(y*log(yhat)+(1-y)*log(1-yhat))   README.md:135-137, test_algs.jl:10
    out_fn     = 1/(1+exp(-A*x))                             README.md:139, test_algs.jl:11

    label_mode "literal": the cross-entropy is fed y as given (the README feeds it +-1, SURVEY
    quirk 8); "consistent": the cross-entropy sees (y+1)/2.
    """

    kind = "logistic"

    def __init__(self, scale: float, label_mode: str = "literal"):
        assert label_mode in ("literal", "consistent")
        self.scale = float(scale)
        self.label_mode = label_mode

    # objective f(A,y,x)
    def f(self, A, y, x):
        z = A @ x
        return self.scale * np.sum(np.log(1.0 + np.exp(-y * z)))

    def fz(self, z, y):
        return self.scale * np.sum(np.log(1.0 + np.exp(-y * z)))

    # d f / d z_i and d2 f / d z_i^2 of the (A,y,x) form (what ForwardDiff differentiates)
    def grad_weights(self, z, y):
        e = np.exp(-y * z)
        return self.scale * (-y) * (e / (1.0 + e))

    def hess_weights(self, z, y):
        e = np.exp(-y * z)
        return self.scale * (y * y) * (e / ((1.0 + e) * (1.0 + e)))

    def _ce_labels(self, y):
        return y if self.label_mode == "literal" else (y + 1.0) / 2.0

    # GGN pieces (prox-GGN-SCORE.jl:44-56): J = diag(s)A, residual = df/dyhat, Q = diag(q)
    def ggn_parts(self, z, y):
        yc = self._ce_labels(y)
        e = np.exp(-z)
        yhat = 1.0 / (1.0 + e)
        s = (yhat / (1.0 + e)) * e  # d yhat / d z as ForwardDiff's x/Dual rule builds it
        om = 1.0 - yhat
        residual = -self.scale * (yc / yhat - (1.0 - yc) / om)
        q = self.scale * (yc / (yhat * yhat) + (1.0 - yc) / (om * om))
        return s, residual, q

    # (r, w) with J'res = A'r, J'QJ = A'diag(w)A
    def ggn_weights(self, z, y):
        s, residual, q = self.ggn_parts(z, y)
        return s * residual, (s * s) * q

    def has_out_fn(self):
        return True


class LeastSquaresLoss:
    """f(A,y,x) = 0.5*sum((A*x-y).^2)/denom   README.md:212-214;  f(y,yhat), out_fn=A*x README.md:233-239"""

    kind = "leastsquares"

    def __init__(self, denom: float):
        self.denom = float(denom)

    def f(self, A, y, x):
        z = A @ x
        return 0.5 * np.sum((z - y) ** 2) / self.denom

    def fz(self, z, y):
        return 0.5 * np.sum((z - y) ** 2) / self.denom

    def grad_weights(self, z, y):
        return (z - y) / self.denom

    def hess_weights(self, z, y):
        return np.full_like(z, 1.0 / self.denom)

    def ggn_parts(self, z, y):
        return np.ones_like(z), (z - y) / self.denom, np.full_like(z, 1.0 / self.denom)

    def ggn_weights(self, z, y):
        return (z - y) / self.denom, np.full_like(z, 1.0 / self.denom)

    def has_out_fn(self):
        return True


class QuadFormLoss:
    """f(A,y,x) = 1/2*(x'*(A*x)) + y'*x   test/test_algs.jl:90 (A square).  Newton / L-BFGS only."""

    kind = "quadform"

    def f(self, A, y, x):
        return 0.5 * (x @ (A @ x)) + y @ x

    def grad(self, A, y, x):
        return 0.5 * ((A @ x) + (A.T @ x)) + y

    def hess(self, A, y, x):
        return 0.5 * (A + A.T)

    def has_out_fn(self):
        return False


# --------------------------------------------------------------------------------------
# Group structure  (src/utils/prox-reg-utils.jl:9-62, 84-142)
# --------------------------------------------------------------------------------------
class GroupStructure:
    """get_P(n, G, ind): ind is 3 x grpNUM (1-based start, end, integer weight); G a permutation.

    Cmat (prox-reg-utils.jl:121-142) is the SV x n sparse matrix with C[k, k] = weight of the
    group covering position k, i.e. diag(weights) for groups that tile 1..n contiguously.
    """

    def __init__(self, n: int, G, ind):
        ind = np.asarray(ind, dtype=np.int64)
        assert ind.shape[0] == 3
        self.n = int(n)
        self.ind = ind
        self.grpNUM = ind.shape[1]
        self.grpSIZES = ind[1] - ind[0] + 1
        self.ntotal = int(self.grpSIZES.sum())
        self.G = np.asarray(G, dtype=np.int64)  # 1-based
        cw = np.zeros(self.n)
        for g in range(self.grpNUM):
            cw[ind[0, g] - 1 : ind[1, g]] = ind[2, g]
        self.cdiag = cw  # Cmat == diag(cdiag) for contiguous tiling groups

    def Cmat_times(self, v):
        return self.cdiag * v

    def matrix_times(self, x):  # P.matrix*x, prox-reg-utils.jl:31
        return x[self.G - 1]

    def twonorm(self, z, g_start, g_end):  # prox-reg-utils.jl:112-119 (sequential sum)
        nrm2 = 0.0
        for i in range(g_start - 1, g_end):
            nrm2 += z[i] ** 2
        return math.sqrt(nrm2)

    def ProxL2(self, x, lam, h):  # prox-reg-utils.jl:84-99
        Px = np.empty_like(x)
        for j in range(self.grpNUM):
            bg = lam * self.ind[2, j]
            gs, ge = int(self.ind[0, j]), int(self.ind[1, j])
            nrm = self.twonorm(x, gs, ge)
            with np.errstate(divide="ignore", invalid="ignore"):
                for k in range(gs - 1, ge):
                    Px[k] = x[k] * max(1.0 - np.float64(bg) / (h[k] * np.float64(nrm)), 0.0)
        return Px

    def fz(self, z):  # prox-reg-utils.jl:101-110
        s = 0.0
        for j in range(self.grpNUM):
            gs, ge = int(self.ind[0, j]), int(self.ind[1, j])
            s += self.ind[2, j] * self.twonorm(z, gs, ge)
        return s


def bounds_sanity_check(n, lb, ub):  # prox-reg-utils.jl:144-159
    lb = np.atleast_1d(np.asarray(lb, dtype=np.float64))
    ub = np.atleast_1d(np.asarray(ub, dtype=np.float64))
    if lb.size == 1 and ub.size == 1:
        a = np.repeat(lb[0], n)
        b = np.repeat(ub[0], n)
    elif lb.size == n and ub.size == n:
        a, b = lb.copy(), ub.copy()
    else:
        raise ValueError("Lengths of the bounds do not match that of the variable.")
    a[a == -np.inf] = L_INF_CACHE
    b[b == np.inf] = U_INF_CACHE
    return a, b


# --------------------------------------------------------------------------------------
# Smoothers (src/regularizers/*.jl): each has mu, Mh, nu, grad(Cmat,x), hess(Cmat,x) (diagonal)
# --------------------------------------------------------------------------------------
def pseudo_huber(x, mu):  # phuber-smooth.jl:28-30
    return (mu**2 - mu * np.sqrt(mu**2 + x**2) + x**2) * (mu**2 + x**2) ** (-1 / 2)


def huber_grad(x, mu):  # phuber-smooth.jl:31-33
    return x * (mu**2 + x**2) ** -(1 / 2)


def huber_hess(x, mu):  # phuber-smooth.jl:34-36
    return mu**2 * (mu**2 + x**2) ** -(3 / 2)


class PHuberSmootherL1L2:  # phuber-smooth.jl:6-27
    name = "phuber_l1l2"

    def __init__(self, mu):
        self.mu, self.Mh, self.nu = float(mu), 2.0, 2.6

    def grad(self, Cmat, x):
        return huber_grad(x, self.mu)

    def hess(self, Cmat, x):
        return huber_hess(x, self.mu)


class PHuberSmootherIndBox:  # phuber-smooth.jl:38-114
    name = "phuber_indbox"

    def __init__(self, lb, ub, mu):
        self.mu, self.Mh, self.nu = float(mu), 2.0, 2.6
        self.lb, self.ub = lb, ub

    def grad(self, Cmat, x):  # :83-98, including the `-x[i] < a[i]` test
        mu = self.mu
        n = x.shape[0]
        a, b = bounds_sanity_check(n, self.lb, self.ub)
        g = np.empty(n)
        for i in range(n):
            if -x[i] < a[i]:
                g[i] = (a[i] ** 2 - 2 * x[i] * a[i] + mu**2 + x[i] ** 2) ** (-1 / 2) * (-x[i] + a[i])
            elif x[i] == a[i] or x[i] < b[i]:
                g[i] = EPS
            else:
                g[i] = (b[i] ** 2 - 2 * b[i] * x[i] + mu**2 + x[i] ** 2) ** (-1 / 2) * (b[i] - x[i])
        return g

    def hess(self, Cmat, x):  # :99-114
        mu = self.mu
        n = x.shape[0]
        a, b = bounds_sanity_check(n, self.lb, self.ub)
        h = np.empty(n)
        for i in range(n):
            if x[i] <= a[i]:
                h[i] = mu**2 * (a[i] ** 2 - 2 * a[i] * x[i] + mu**2 + x[i] ** 2) ** (-3 / 2)
            elif a[i] < x[i] < b[i]:
                h[i] = EPS
            elif x[i] >= b[i]:
                h[i] = mu**2 * (b[i] ** 2 - 2 * b[i] * x[i] + mu**2 + x[i] ** 2) ** (-3 / 2)
            else:  # NaN input: Julia leaves the `undef` slot untouched
                h[i] = np.nan
        return h


class PHuberSmootherGL:  # phuber-smooth.jl:116-164 (lambda1/lambda2 unused there, quirk 4)
    name = "phuber_gl"

    def __init__(self, mu, model):
        self.mu, self.Mh, self.nu = float(mu), 2.0, 2.6
        self.P = model.P

    def grad(self, Cmat, x):  # :150-155
        g1 = pseudo_huber(x, self.mu)
        Dg = huber_grad(x, self.mu)
        return huber_grad(Cmat.Cmat_times(g1), self.mu) * Dg

    def hess(self, Cmat, x):  # :156-164
        g1 = pseudo_huber(x, self.mu)
        Dg = huber_grad(x, self.mu)
        DDg = huber_hess(x, self.mu)
        c = Cmat.Cmat_times(g1)
        return huber_hess(c, self.mu) * np.dot(Dg, Dg) + huber_grad(c, self.mu) * DDg


class ExponentialSmootherIndBox:  # exponential-smooth.jl:28-50 (ignores ub, quirk 9)
    name = "exp_indbox"

    def __init__(self, lb, ub, mu):
        self.mu, self.Mh, self.nu = float(mu), 1.0, 2.0
        self.lb, self.ub = lb, ub

    def grad(self, Cmat, x):
        a, _ = bounds_sanity_check(x.shape[0], self.lb, self.ub)
        return -np.exp((-x + a) / self.mu)

    def hess(self, Cmat, x):
        a, _ = bounds_sanity_check(x.shape[0], self.lb, self.ub)
        return 1 / self.mu * np.exp((-x + a) / self.mu)


class LogExpSmootherIndBox:  # log-exp-smooth.jl:28-61
    name = "logexp_indbox"

    def __init__(self, lb, ub, mu):
        self.mu, self.Mh, self.nu = float(mu), 1.0, 2.0
        self.lb, self.ub = lb, ub

    def grad(self, Cmat, x):  # :47-54
        mu = self.mu
        a, b = bounds_sanity_check(x.shape[0], self.lb, self.ub)
        with np.errstate(divide="ignore", invalid="ignore"):
            t1 = np.where(x <= a + mu, (x - a - 2 * mu) / mu, np.where(x >= b - mu, (x - b + 2 * mu) / mu, 0.0))
            t2 = np.where(x < a, mu / (a - x), np.where(x > b, -mu / (b - x), 0.0))
        return t1 + t2

    def hess(self, Cmat, x):  # :56-63
        mu = self.mu
        a, b = bounds_sanity_check(x.shape[0], self.lb, self.ub)
        with np.errstate(divide="ignore", invalid="ignore"):
            t1 = np.where(x <= a + mu, 1 / mu, np.where(x >= b - mu, 1 / mu, 0.0))
            t2 = np.where(x < a, mu / (a - x) ** 2, np.where(x > b, mu / (b - x) ** 2, 0.0))
        return t1 + t2


def osba_smooth_l1(x, mu):  # ostrovskii-bach-smooth.jl:28-30
    with np.errstate(divide="ignore", invalid="ignore"):
        sq = np.sqrt(mu**2 + 4 * x**2)
        return (sq / 2 - mu / 2 + mu * np.log((2 * x - sq + mu) / x) / 2 - math.log(2) * mu
                + mu * np.log((sq - mu + 2 * x) / x) / 2)


def osba_smooth_grad_l1(x, mu):  # :31-33
    with np.errstate(divide="ignore", invalid="ignore"):
        sq = np.sqrt(mu**2 + 4 * x**2)
        return ((-(mu**3) + mu**2 * sq - 4 * x**2 * mu + 2 * x**2 * sq) * (mu * sq + mu**2 + 4 * x**2)
                / (4 * mu**2 * x**3 + 16 * x**5))


def osba_smooth_hess_l1(x, mu):  # :34-36
    with np.errstate(divide="ignore", invalid="ignore"):
        sq = np.sqrt(mu**2 + 4 * x**2)
        return (sq - mu) * mu / x**2 * (mu**2 + 4 * x**2) ** (-1 / 2) / 2


class OsBaSmootherL1L2:  # ostrovskii-bach-smooth.jl:6-27
    name = "osba_l1l2"

    def __init__(self, mu):
        self.mu, self.Mh, self.nu = float(mu), 2 * math.sqrt(2), 3.0

    def grad(self, Cmat, x):
        return osba_smooth_grad_l1(x, self.mu)

    def hess(self, Cmat, x):
        return osba_smooth_hess_l1(x, self.mu)


class OsBaSmootherGL:  # ostrovskii-bach-smooth.jl:38-85
    name = "osba_gl"

    def __init__(self, mu, model):
        self.mu, self.Mh, self.nu = float(mu), 2 * math.sqrt(2), 3.0
        self.P = model.P

    def grad(self, Cmat, x):
        g1 = osba_smooth_l1(x, self.mu)
        Dg = osba_smooth_grad_l1(x, self.mu)
        return osba_smooth_grad_l1(Cmat.Cmat_times(g1), self.mu) * Dg

    def hess(self, Cmat, x):
        g1 = osba_smooth_l1(x, self.mu)
        Dg = osba_smooth_grad_l1(x, self.mu)
        DDg = osba_smooth_hess_l1(x, self.mu)
        c = Cmat.Cmat_times(g1)
        return osba_smooth_hess_l1(c, self.mu) * np.dot(Dg, Dg) + osba_smooth_grad_l1(c, self.mu) * DDg


def get_Mg(Mh, nu, mu, n):  # smoothing.jl:12-26
    if Mh < 0:
        raise ValueError("Mh must be nonnegative.")
    elif mu <= 0:
        raise ValueError("μ must be positive.")
    if 0 < nu <= 3:
        return float(n) ** ((3 - nu) / 2) * mu ** (nu / 2 - 2) * Mh
    elif nu > 3:
        return mu ** (4 - 3 * nu / 2) * Mh
    raise ValueError("ν must be positive.")


# --------------------------------------------------------------------------------------
# Problem (src/problems.jl:21-40,61-81)
# --------------------------------------------------------------------------------------
@dataclass
class Problem:
    A: np.ndarray
    y: np.ndarray
    x0: np.ndarray
    f: object  # a built-in loss object
    lam: object  # scalar or (lam1, lam2)
    L: Optional[float] = None
    sol: Optional[np.ndarray] = None
    C_set: Optional[tuple] = None
    P: Optional[GroupStructure] = None
    Atest: Optional[np.ndarray] = None  # problems.jl:28-29: held-out data, only ever used for ftest(x) (iterate.jl:169-176)
    ytest: Optional[np.ndarray] = None

    def __post_init__(self):
        self.A = np.asarray(self.A, dtype=np.float64)
        self.y = np.asarray(self.y, dtype=np.float64)
        self.x0 = np.asarray(self.x0, dtype=np.float64)
        self.x = np.zeros_like(self.x0) if self.sol is None else np.asarray(self.sol, dtype=np.float64)

    def lam_scalar(self):  # prox-*-SCORE.jl: `length(model.λ) > 1 ? model.λ[1] : model.λ`
        return float(self.lam[0]) if np.ndim(self.lam) > 0 and len(self.lam) > 1 else float(np.ravel(self.lam)[0])


def _bounds_of(model):  # prox-operators.jl:36-45 / regularizers.jl:9-17 (Vector / Tuple form)
    return model.C_set[0], model.C_set[1]


def get_reg(model, x, reg_name):  # regularizers.jl:4-31
    if reg_name == "l1":
        return _lam_reg(model) * np.sum(np.abs(x))
    elif reg_name == "l2":
        return _lam_reg(model) * np.sum(np.abs(x) ** 2)
    elif reg_name == "indbox":
        lb, ub = _bounds_of(model)
        return np.inf if (np.any(x < lb) or np.any(x > ub)) else 0.0  # :33-39
    elif reg_name == "gl":
        if np.ndim(model.lam) == 0 or len(model.lam) != 2:
            raise ValueError("Please provide a Tuple or Vector with exactly two entries for λ, e.g. [λ1, λ2]")
        P = model.P
        Px = P.matrix_times(x)
        lam1, lam2 = float(model.lam[0]), float(model.lam[1])
        return lam2 * P.fz(Px) + lam1 * np.sum(np.abs(x))
    raise ValueError("reg_name not valid.")


def _lam_reg(model):
    # regularizers.jl:6,8 multiplies by model.λ itself; for l1/l2 λ is a scalar
    return float(np.ravel(model.lam)[0])


# --------------------------------------------------------------------------------------
# Scaled proximal operators (src/prox/prox-operators.jl)
# --------------------------------------------------------------------------------------
def prox_step(model, reg_name, x, h_scale, lam, alpha):
    if reg_name == "l1":  # :8-12
        t = alpha * lam / h_scale
        return np.sign(x) * np.maximum(np.abs(x) - t, 0)
    elif reg_name == "l2":  # :21-25
        t = alpha * lam / h_scale
        with np.errstate(divide="ignore", invalid="ignore"):
            return x * np.maximum(1 - t / np.abs(x) ** 2, 0)
    elif reg_name == "indbox":  # :34-46
        lb, ub = _bounds_of(model)
        return np.minimum(np.maximum(x, lb), ub)
    elif reg_name == "gl":  # :55-66
        P = model.P
        lam1, lam2 = float(model.lam[0]), float(model.lam[1])
        t = lam1 / h_scale
        utmp = np.sign(x) * np.maximum(np.abs(x) - t, 0)
        return P.ProxL2(utmp, alpha * lam2, h_scale)
    raise ValueError("reg_name not valid.")


# --------------------------------------------------------------------------------------
# Step-size helpers (src/utils/utils.jl:27-48)
# --------------------------------------------------------------------------------------
def linesearch(x, d, f, grad_f):  # :27-35 (re-evaluates f(x) and grad_f(x) each trial, as written)
    alpha, rho, c = 1.0, 0.5, 1e-4
    while f(x + alpha * d) > f(x) + c * alpha * np.dot(grad_f(x), d):
        alpha = rho * alpha
    return alpha


def inv_BB_step(x, x_prev, gradx, gradx_prev):  # :43-48
    delta = x - x_prev
    gamma = gradx - gradx_prev
    return np.dot(gamma, gamma) / np.dot(delta, gamma)


# --------------------------------------------------------------------------------------
# Methods
# --------------------------------------------------------------------------------------
def _model_grad_f(model, x):
    """gradient(f, x) for the built-in closures (ForwardDiff path, prox-N-SCORE.jl:56-64)."""
    L = model.f
    if L.kind == "quadform":
        return L.grad(model.A, model.y, x)
    z = model.A @ x
    return model.A.T @ L.grad_weights(z, model.y)


def _model_hess_f(model, x):
    L = model.f
    if L.kind == "quadform":
        return L.hess(model.A, model.y, x)
    z = model.A @ x
    w = L.hess_weights(z, model.y)
    return model.A.T @ (w[:, None] * model.A)


def _step_size(method, model, it, x, x_prev, d, hmu, Cmat, lam, reg_name, grad_q_x, is_lqn):
    """prox-N-SCORE.jl:73-90, prox-GGN-SCORE.jl:70-87, prox-L-BFGS-SCORE.jl:108-125."""
    obj = lambda v: model.f.f(model.A, model.y, v) + get_reg(model, v, reg_name)
    grad_q = lambda v: _model_grad_f(model, v) + lam * hmu.grad(Cmat, v)
    st = method.ss_type
    if st == 1 and model.L is not None:
        return min(1 / model.L, 1.0)
    elif st == 1 and model.L is None:
        return 0.5
    elif st == 2 or (is_lqn and model.L is None):
        if not is_lqn:
            # prox-N-SCORE.jl:81-83 / prox-GGN-SCORE.jl:78-80 reference an undefined `∇f` and call
            # hμ.grad with one argument -> the reference throws
            if it == 1:
                return 1
            raise NameError("UndefVarError: ∇f not defined (reference ss_type=2 for N/GGN)")
        if it == 1:
            return 1
        grad_q_prev = _model_grad_f(model, x_prev) + lam * hmu.grad(Cmat, x_prev)
        return inv_BB_step(x, x_prev, grad_q_x, grad_q_prev)
    elif st == 3:
        return linesearch(x, d, obj, grad_q)
    raise ValueError("Please, choose ss_type in [1, 2, 3].")


def _tail(method, model, reg_name, hmu, x, d, lam, lgr, Hr_diag, step_size):
    """Common damping + prox tail: prox-N-SCORE.jl:92-112, GGN :89-104, LQN :127-146."""
    Hdiag_inv = 1 / Hr_diag
    Mg = get_Mg(hmu.Mh, hmu.nu, hmu.mu, x.shape[0])
    eta = np.sqrt(np.dot(lgr, Hdiag_inv * lgr))
    alpha = step_size / (1 + Mg * eta)
    safe_alpha = min(1, alpha)
    dx = safe_alpha * d
    if method.use_prox:
        x_new = prox_step(model, reg_name, x + dx, Hdiag_inv, lam, step_size)
        delta = x_new - x
    else:
        x_new = x + dx
        delta = dx
    return x_new, dx, delta


@dataclass
class ProxNSCORE:  # prox-N-SCORE.jl:6-22
    ss_type: int = 1
    use_prox: bool = True
    name: str = "prox-newtonscore"

    def init(self, x):
        pass

    def step(self, model, reg_name, hmu, x, x_prev, Cmat, it):  # :34-119
        lam = model.lam_scalar()
        gr = hmu.grad(Cmat, x)
        lgr = lam * gr
        Hr_diag = hmu.hess(Cmat, x)
        H = _model_hess_f(model, x)
        gq = _model_grad_f(model, x) + lgr
        sol = np.linalg.solve(H + lam * np.diag(Hr_diag), gq)  # :70 generic `\` -> LU
        d = -sol
        ss = _step_size(self, model, it, x, x_prev, d, hmu, Cmat, lam, reg_name, gq, False)
        x_new, dx, delta = _tail(self, model, reg_name, hmu, x, d, lam, lgr, Hr_diag, ss)
        return x_new, float(np.linalg.norm(delta))


def ggn_score_step(model, z, gr, Hr_diag, lam):
    """prox-GGN-SCORE.jl:114-135.  Tall case (n+1 > m) in collapsed form: the
    augmented row/column of Q is zero and the augmented residual entry is 1, so
    JQJ = A'diag(w)A + λ·diag(Hr), Je = A'r + λ·gr.  Wide case (:124-127) restated literally."""
    A = model.A
    n_rows, m = A.shape
    if n_rows + 1 <= m:
        s, res, q = model.f.ggn_parts(z, model.y)
        J = s[:, None] * A
        Jt = np.hstack([J.T, (lam * gr)[:, None]])  # :121
        residual = np.concatenate([res, [1.0]])  # :122
        Q = np.zeros((n_rows + 1, n_rows + 1))  # :123
        Q[:n_rows, :n_rows] = np.diag(q)
        H_inv = 1 / Hr_diag
        A_ = Q @ (Jt.T * H_inv[None, :]) @ Jt  # :125
        B = np.linalg.solve(np.eye(n_rows + 1) + A_, residual)  # :126
        d = H_inv * (Jt @ B)  # :127
        return -d
    r, w = model.f.ggn_weights(z, model.y)
    JQJ = A.T @ (w[:, None] * A) + lam * np.diag(Hr_diag)  # :129
    Je = A.T @ r + lam * gr  # :130
    return -np.linalg.solve(JQJ, Je)  # :131,134


@dataclass
class ProxGGNSCORE:  # prox-GGN-SCORE.jl:6-22
    ss_type: int = 1
    use_prox: bool = True
    name: str = "prox-ggnscore"

    def init(self, x):
        pass

    def step(self, model, reg_name, hmu, x, x_prev, Cmat, it):  # :34-112
        if not model.f.has_out_fn():
            raise TypeError("ProxGGNSCORE needs out_fn (MethodError in the reference)")
        lam = model.lam_scalar()
        gr = hmu.grad(Cmat, x)
        lgr = lam * gr
        Hr_diag = hmu.hess(Cmat, x)
        z = model.A @ x
        d = ggn_score_step(model, z, gr, Hr_diag, lam)
        gq = None
        if self.ss_type != 1:
            gq = _model_grad_f(model, x) + lgr
        ss = _step_size(self, model, it, x, x_prev, d, hmu, Cmat, lam, reg_name, gq, False)
        x_new, dx, delta = _tail(self, model, reg_name, hmu, x, d, lam, lgr, Hr_diag, ss)
        return x_new, float(np.linalg.norm(delta))


@dataclass
class ProxLQNSCORE:  # prox-L-BFGS-SCORE.jl:6-30
    ss_type: int = 1
    use_prox: bool = True
    m: int = 10
    name: str = "prox-lbfgsscore"
    s_list: list = field(default_factory=list)
    y_list: list = field(default_factory=list)
    H0: float = 1.0

    def init(self, x):  # :31-36
        self.s_list, self.y_list, self.H0 = [], [], 1.0

    def two_loop_recursion(self, grad):  # :47-68
        q = grad.copy()
        alpha, rho = [], []
        for s, y in zip(reversed(self.s_list), reversed(self.y_list)):
            rho_i = 1.0 / np.dot(y, s)
            alpha_i = rho_i * np.dot(s, q)
            q = q - alpha_i * y
            alpha.append(alpha_i)
            rho.append(rho_i)
        r = self.H0 * q
        k = len(self.s_list)
        for i in range(k):
            s, y = self.s_list[i], self.y_list[i]
            rho_i = rho[k - 1 - i]
            alpha_i = alpha[k - 1 - i]
            beta = rho_i * np.dot(y, r)
            r = r + s * (alpha_i - beta)
        return -r

    def step(self, model, reg_name, hmu, x, x_prev, Cmat, it):  # :69-169
        lam = model.lam_scalar()
        gr = hmu.grad(Cmat, x)
        lgr = lam * gr
        Hr_diag = hmu.hess(Cmat, x)
        gq = _model_grad_f(model, x) + lgr
        if it == 1 or len(self.s_list) == 0:
            d = -gq
        else:
            d = self.two_loop_recursion(gq)
        ss = _step_size(self, model, it, x, x_prev, d, hmu, Cmat, lam, reg_name, gq, True)
        x_new, dx, delta = _tail(self, model, reg_name, hmu, x, d, lam, lgr, Hr_diag, ss)
        gq_new = _model_grad_f(model, x_new) + lam * hmu.grad(Cmat, x_new)
        gamma = gq_new - gq
        if np.dot(delta, gamma) > 1e-10:  # :154
            if len(self.s_list) == self.m:
                self.s_list.pop(0)
                self.y_list.pop(0)
            self.s_list.append(delta)
            self.y_list.append(gamma)
            self.H0 = float(np.dot(gamma, delta) / np.dot(gamma, gamma))
        return x_new, float(np.linalg.norm(delta))


# --------------------------------------------------------------------------------------
# Driver (src/algorithms/iterate.jl:56-76, 100-266): full batch, mini-batches, slice_samples
# --------------------------------------------------------------------------------------
@dataclass
class Solution:  # iterate.jl:3-32
    x: np.ndarray
    obj: list
    fval: list
    pri_res_norm: list
    rel: list
    objrel: list
    epochs: int
    iterates: list  # oracle extra: x after each step (not in the reference's Solution)
    fvaltest: list = field(default_factory=list)  # f(Atest, ytest, x) at every recorded state (utils.jl:55-57)


def _norm(v):
    return float(np.linalg.norm(v))


def make_batches(n, batch_size=None, slice_samples=False, shuffle_batch=False, local_max_iter=None, perm=None):
    """Row index sets of the batches optim_loop! iterates over (iterate.jl:122-145, utils.jl:14-25).

    MLUtils.DataLoader(batchsize=b, shuffle, partial=true) over n observations: ceil(n/b) consecutive batches of the
    (once) shuffled order, the last one possibly short.  `collect(loader)` runs the loader ONCE before the epoch
    loop, so the same batches are reused in every epoch.  slice_samples yields to batch_size when both are given
    (:127-130); on its own it steps on the first row only (see below).  local_max_iter keeps only the first min(floor(local_max_iter), max_iter) batches
    (:124,:127,:145).  The shuffle is Julia's RNG upstream; here the permutation is an explicit input (`perm`) so
    that the oracle and the GPU path see the same batches."""
    if batch_size is not None and slice_samples:
        slice_samples = False
    # max_iter / iend are fixed BEFORE slice_samples sets opt.batch_size = 1 (:124-127 vs :136-138): without a
    # batch_size max_iter stays 1, so `get_loader_subset(data, 1:iend)` (:145) keeps ONE entry — the whole data for the
    # full-batch loader, only the FIRST ROW for slice_samples (the other rows never take part in a step).
    max_iter = -(-n // int(batch_size)) if batch_size is not None else 1
    iend = max_iter
    if local_max_iter is not None and int(np.floor(local_max_iter)) > 0:
        iend = min(int(np.floor(local_max_iter)), max_iter)
    if slice_samples:
        batch_size = 1
        shuffle_batch = False
    if batch_size is None:
        batch_size = n
        shuffle_batch = False
    batch_size = int(batch_size)
    order = np.arange(n)
    if shuffle_batch:
        if perm is None:
            raise ValueError("shuffle_batch needs an explicit permutation (perm=) in this restatement")
        order = np.asarray(perm, dtype=np.int64)
        assert sorted(order.tolist()) == list(range(n))
    return [order[i * batch_size:min((i + 1) * batch_size, n)] for i in range(iend)]


def iterate(method, model, reg_name, hmu, alpha=None, max_epoch=1000, x_tol=1e-10, f_tol=1e-10, batch_size=None,
            slice_samples=False, shuffle_batch=False, local_max_iter=None, perm=None):
    import copy
    if local_max_iter is not None:  # iterate!: Options(max_epoch = local_max_iter !== nothing ? 1 : max_epoch) (:58-70)
        max_epoch = 1
    if alpha is not None:  # iterate.jl:113-115
        model.L = 1 / alpha
    n = model.A.shape[0]
    batches = make_batches(n, batch_size, slice_samples, shuffle_batch, local_max_iter, perm)
    full = len(batches) == 1 and len(batches[0]) == n and not shuffle_batch
    views = []
    for idx in batches:  # As = Matrix(As'), ys = vec(ys') (:206-207): the batch rows as their own problem data
        if full:
            views.append(model)
        else:
            bm = copy.copy(model)
            bm.A = np.ascontiguousarray(model.A[idx])
            bm.y = model.y[idx]
            views.append(bm)
    iend = len(batches)
    f = lambda v: model.f.f(model.A, model.y, v)
    objs, fvals, pris, rels, frels, iterates = [], [], [], [], [], []
    epochs = 0
    x_star = model.x
    pri_res_norm = None
    with np.errstate(all="ignore"):
        obj_star = f(x_star) + get_reg(model, x_star, reg_name)  # :179
    x = model.x0.copy()
    x_prev = x.copy()
    method.init(x)  # :183
    Cmat = model.P if reg_name == "gl" else None

    def rel_err(v):  # :192-197
        if reg_name == "gl":
            return float(np.mean((x_star - v) ** 2))
        return max(_norm(v - x_star) / max(_norm(x_star), 1), x_tol)

    def frel(obj):  # :200  (np.maximum propagates NaN like Julia's max)
        with np.errstate(all="ignore"):
            return float(np.maximum(np.abs(obj - obj_star) / np.abs(obj_star), f_tol))

    def push(obj, fval, pri, rel, fr):  # utils.jl:106-113
        objs.append(obj), fvals.append(fval), pris.append(pri), rels.append(rel), frels.append(fr)

    test_model = model.Atest is not None and model.ytest is not None  # iterate.jl:169
    fvaltests = []

    def stats(v):
        with np.errstate(all="ignore"):
            fval = float(f(v))
            obj = fval + float(get_reg(model, v, reg_name))
            if test_model:  # show_stat! pushes ftest(x) next to every recorded state (utils.jl:55-57)
                fvaltests.append(float(model.f.f(np.asarray(model.Atest, float), np.asarray(model.ytest, float), v)))
        return obj, fval, rel_err(v), frel(obj)

    for epoch_t in range(1, max_epoch + 1):  # :185
        obj, fval, rel_error, f_rel_error = stats(x)
        push(obj, fval, pri_res_norm, rel_error, f_rel_error)  # :202
        for i, bm in enumerate(views, start=1):  # :204
            if epoch_t == max_epoch and i == iend:  # :219-231: stats of the current x once more (quirk 2)
                obj, fval, rel_error, f_rel_error = stats(x)
                push(obj, fval, pri_res_norm, rel_error, f_rel_error)
            x_new, pri_res_norm = method.step(bm, reg_name, hmu, x, x_prev, Cmat, epoch_t)  # :233
            iterates.append(x_new.copy())
            if _norm(x_new - x) < x_tol * max(_norm(x), 1) or f_rel_error <= f_tol or pri_res_norm < x_tol:  # :234
                if epoch_t != max_epoch:  # :235-247 (one extra entry for x_new; refreshes f_rel_error)
                    obj, fval, rel_error, f_rel_error = stats(x_new)
                    push(obj, fval, pri_res_norm, rel_error, f_rel_error)
                x_prev = x.copy()  # :248-251
                x = x_new
                epochs += 1
                break
            x_prev = x.copy()  # :253-254
            x = x_new
        if _norm(x - x_prev) < x_tol * max(_norm(x_prev), 1) or f_rel_error <= f_tol or pri_res_norm < x_tol:  # :257
            break
        epochs += 1  # :261
    return Solution(x, objs, fvals, pris, rels, frels, epochs, iterates, fvaltests)
