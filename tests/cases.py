"""Small-shape versions of the five BASELINE.json configs, shared by the golden generator, the oracle tests
and the GPU parity tests.  `build(name, lib)` returns (method, model, reg_name, hmu, iterate_kwargs) built from
either the oracle module or the scs_b200 host mirror — the two expose the same constructor names on purpose.
Data comes from oracle/synth.py (seeded Philox), so every consumer sees identical bits.
"""
from __future__ import annotations

import numpy as np

from oracle import synth

CASES = ("c1_readme_logreg_n", "c2_logreg_ggn_l1", "c3_logreg_lqn_l1", "c4_ls_ggn_gl", "c5_ls_n_indbox",
         "c2b_logreg_ggn_literal", "c3b_logreg_lqn_l2_bb", "c5b_ls_lqn_logexp")


def _logistic_data(n, m, density=1.0, zero_prob=False):
    A = synth.make_A(n, m, seed=1234, density=density)
    xt = synth.make_x_true(m, seed=1235)
    z = np.zeros(n) if zero_prob else A @ xt
    y = synth.make_labels_logistic(z, seed=1236)
    return A, y, synth.make_x0(m, seed=1237)


def _ls_data(n, m):
    A = synth.make_A(n, m, seed=1234)
    xt = synth.make_x_true(m, seed=1235)
    y = synth.make_targets_ls(A @ xt, seed=1236)
    return A, y, synth.make_x0(m, seed=1237)


def data(name):
    if name == "c1_readme_logreg_n":
        return _logistic_data(100, 50, density=0.01, zero_prob=True)
    if name in ("c2_logreg_ggn_l1", "c2b_logreg_ggn_literal"):
        return _logistic_data(4096, 256)
    if name in ("c3_logreg_lqn_l1", "c3b_logreg_lqn_l2_bb"):
        return _logistic_data(4096, 128)
    if name == "c4_ls_ggn_gl":
        return _ls_data(2048, 256)
    if name in ("c5_ls_n_indbox", "c5b_ls_lqn_logexp"):
        return _ls_data(2048, 128)
    raise KeyError(name)


def build(name, lib, A=None, y=None, x0=None, **problem_kw):
    """lib: `oracle.scs_oracle` or `scs_b200`.  Returns (method, model, reg_name, hmu, kwargs for iterate)."""
    if A is None:
        A, y, x0 = data(name)
    n, m = A.shape
    gp = getattr(lib, "GroupStructure", None) or lib.get_P
    if name == "c1_readme_logreg_n":  # README.md:95-125
        model = lib.Problem(A, y, x0, lib.LogisticLoss(1 / m), 1e-1, **problem_kw)
        return lib.ProxNSCORE(), model, "l1", lib.PHuberSmootherL1L2(1.0), dict(max_epoch=100, x_tol=1e-6, f_tol=1e-6)
    if name == "c2_logreg_ggn_l1":
        model = lib.Problem(A, y, x0, lib.LogisticLoss(1 / n, "consistent"), 1e-2, **problem_kw)
        return lib.ProxGGNSCORE(), model, "l1", lib.PHuberSmootherL1L2(1.0), dict(max_epoch=12, alpha=1)
    if name == "c2b_logreg_ggn_literal":  # the README's label-inconsistent pair (negative Gram weights)
        model = lib.Problem(A, y, x0, lib.LogisticLoss(1 / n, "literal"), 1e-3, **problem_kw)
        return lib.ProxGGNSCORE(), model, "l1", lib.PHuberSmootherL1L2(1.0), dict(max_epoch=6, alpha=1)
    if name == "c3_logreg_lqn_l1":
        model = lib.Problem(A, y, x0, lib.LogisticLoss(1 / n), 1e-2, **problem_kw)
        return lib.ProxLQNSCORE(m=10), model, "l1", lib.PHuberSmootherL1L2(1.0), dict(max_epoch=30, alpha=1)
    if name == "c3b_logreg_lqn_l2_bb":
        model = lib.Problem(A, y, x0, lib.LogisticLoss(1 / n), 1e-3, **problem_kw)
        return lib.ProxLQNSCORE(ss_type=2, m=5), model, "l2", lib.PHuberSmootherL1L2(0.5), dict(max_epoch=12)
    if name == "c4_ls_ggn_gl":
        gsz = 64
        ng = m // gsz
        ind = np.array([[g * gsz + 1 for g in range(ng)], [(g + 1) * gsz for g in range(ng)], [1] * ng])
        P = gp(m, np.arange(1, m + 1), ind)
        model = lib.Problem(A, y, x0, lib.LeastSquaresLoss(n), [1e-8, 1e-2], P=P, **problem_kw)
        return lib.ProxGGNSCORE(), model, "gl", lib.PHuberSmootherGL(1e-2, model), dict(max_epoch=10, alpha=1)
    if name == "c5_ls_n_indbox":
        model = lib.Problem(A, y, x0, lib.LeastSquaresLoss(n), 1e-4, C_set=(-0.5, 0.5), **problem_kw)
        return lib.ProxNSCORE(), model, "indbox", lib.PHuberSmootherIndBox(-0.5, 0.5, 0.6), dict(max_epoch=15, alpha=0.8)
    if name == "c5b_ls_lqn_logexp":
        model = lib.Problem(A, y, x0, lib.LeastSquaresLoss(n), 1.0, C_set=(-0.5, np.inf), **problem_kw)
        return lib.ProxLQNSCORE(m=10), model, "indbox", lib.LogExpSmootherIndBox(-0.5, np.inf, 10.0), dict(max_epoch=15, alpha=1)
    raise KeyError(name)
