"""Host-side mirror of the reference's operator interface for the proximal-SCORE hot path.

Same names, argument meaning and error behaviour as SelfConcordantSmoothOptimization.jl, so the parity tests
read like the reference's own tests (test/test_algs.jl):

    model = Problem(A, y, x0, LogisticLoss(1/5), 1)                       # src/problems.jl:61-81
    sol   = iterate(ProxNSCORE(), model, "l1", PHuberSmootherL1L2(1))     # src/algorithms/iterate.jl:56-76

Everything numeric runs in libscs_b200.so on the GPU; this file only holds the epoch loop bookkeeping of
optim_loop! (iterate.jl:100-266: histories, stopping rules) — the part the survey keeps on the host — and
argument marshalling.  julia/SCSB200.jl mirrors it 1:1 over the same C symbols.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from . import _capi as K
from ._capi import ScsError, UnsupportedError


# ---- contexts ---------------------------------------------------------------------------------
class Context:
    """One per (process, GPU).  world > 1: every rank passes the same 128-byte unique id."""

    def __init__(self, device: int = 0, rank: int = 0, world: int = 1, unique_id: Optional[bytes] = None):
        L = K.lib()
        h = C.c_void_p()
        buf = None
        if world > 1:
            if unique_id is None or len(unique_id) != 128:
                raise ValueError("world > 1 needs the 128-byte unique id from Context.unique_id()")
            buf = C.create_string_buffer(unique_id, 128)
        K.check(L.scs_ctx_create(device, rank, world, buf, C.byref(h)))
        self._h, self.device, self.rank, self.world = h, device, rank, world

    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        K.check(K.lib().scs_comm_unique_id(buf))
        return buf.raw

    def sync(self):
        K.check(K.lib().scs_ctx_sync(self._h))

    def stream(self) -> int:
        s = C.c_uint64()
        K.check(K.lib().scs_ctx_stream(self._h, C.byref(s)))
        return s.value

    def launches(self, reset=False) -> int:
        v = C.c_int64()
        K.check(K.lib().scs_get_counters(self._h, C.byref(v), int(reset)))
        return v.value

    def set_profiling(self, on: bool):
        K.check(K.lib().scs_set_profiling(self._h, int(on)))

    def stage_ms(self, reset=False):
        ms = np.zeros(len(K.STAGES))
        calls = np.zeros(len(K.STAGES), dtype=np.int64)
        K.check(K.lib().scs_get_stage_ms(self._h, K.dptr(ms), K.iptr(calls), int(reset)))
        return {n: (float(ms[i]), int(calls[i])) for i, n in enumerate(K.STAGES)}

    def measure_i8_peak(self, seconds=1.5):
        """(burst, sustained) int8 tensor-pipe TOP/s of this device with the library's own UMMA loop (no memory traffic)."""
        a, b = C.c_double(), C.c_double()
        K.check(K.lib().scs_measure_i8_peak(self._h, float(seconds), C.byref(a), C.byref(b)))
        return a.value, b.value

    def i8_pipe_probe(self, tma_mode):
        """TOP/s of k_i8syrk's main loop on an L2-resident operand (tuning aid; tma_mode 0 / 1 / 2, see scs_b200.h)."""
        a = C.c_double()
        K.check(K.lib().scs_i8_pipe_probe(self._h, int(tma_mode), C.byref(a)))
        return a.value

    def linear_solve(self, M, b):
        """(H + λHr) \\ ∇q on the device: Cholesky, pivoted-LU fallback.  Returns (d, used_fallback)."""
        M = np.asfortranarray(np.asarray(M, dtype=np.float64))
        b = K.vec(b, M.shape[0])
        d = np.empty_like(b)
        fb = C.c_int()
        K.check(K.lib().scs_linear_solve(self._h, K.dptr(M), K.dptr(b), M.shape[0], K.dptr(d), C.byref(fb)))
        return d, bool(fb.value)

    def close(self):
        if getattr(self, "_h", None):
            K.lib().scs_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


# ---- built-in losses (the README / test closures) -------------------------------
@dataclass
class LogisticLoss:
    """f(A,y,x)=scale*sum(log(1+exp(-y.*(A*x)))), f(y,ŷ) cross-entropy, out_fn=σ(Ax)  (README.md:113,135-139)."""
    scale: float
    label_mode: str = "literal"  # "literal": CE sees y as given (README feeds ±1); "consistent": (y+1)/2
    kind = K.LOSS_LOGISTIC

    def param(self):
        return float(self.scale)

    def label_code(self):
        if self.label_mode not in ("literal", "consistent"):
            raise ValueError("label_mode must be 'literal' or 'consistent'")
        return K.LABELS_LITERAL if self.label_mode == "literal" else K.LABELS_CONSISTENT


@dataclass
class LeastSquaresLoss:
    """f(A,y,x)=0.5*sum((A*x-y).^2)/denom, out_fn=A*x  (README.md:212-214,233-239)."""
    denom: float
    kind = K.LOSS_LEASTSQUARES

    def param(self):
        return float(self.denom)

    def label_code(self):
        return 0


@dataclass
class QuadFormLoss:
    """f(A,y,x)=1/2*(x'*(A*x))+y'*x  (test/test_algs.jl:90).  ProxNSCORE / ProxLQNSCORE only."""
    kind = K.LOSS_QUADFORM

    def param(self):
        return 0.0

    def label_code(self):
        return 0


_BUILTIN = (LogisticLoss, LeastSquaresLoss, QuadFormLoss)


class get_P:
    """Group structure for "gl" (src/utils/prox-reg-utils.jl:9-62): ind = 3 x grpNUM (1-based start,end,weight)."""

    def __init__(self, n, G, ind):
        self.ind = np.asfortranarray(np.asarray(ind, dtype=np.int64))
        if self.ind.ndim != 2 or self.ind.shape[0] != 3:
            raise ValueError("ind must be 3 x grpNUM")
        self.n = int(n)
        self.G = np.ascontiguousarray(np.asarray(G, dtype=np.int64))
        self.grpNUM = self.ind.shape[1]


class Problem:
    """Problem(A, y, x0, f, λ; L, sol, C_set, P, out_fn, ...) — src/problems.jl:61-81.

    A is copied to the GPU once here (replacing the per-iteration `Matrix(As')` of iterate.jl:206-207).
    `f` must be one of the built-in loss objects: an arbitrary callable would need ForwardDiff on the CPU and is
    rejected explicitly.  Row-sharded use: each rank passes its own rows and a Context with world > 1.
    """

    def __init__(self, A, y, x0, f, lam, *, L=None, sol=None, C_set=None, P=None, out_fn=None, grad_fx=None,
                 hess_fx=None, jac_yx=None, grad_fy=None, hess_fy=None, Atest=None, ytest=None, name=None, ctx=None,
                 storage="auto"):
        if not isinstance(f, _BUILTIN):
            raise UnsupportedError(K.SCS_UNSUPPORTED,
                                   "arbitrary user f (ForwardDiff-only path) is not supported on the GPU: pass "
                                   "LogisticLoss / LeastSquaresLoss / QuadFormLoss")
        if any(v is not None for v in (out_fn, grad_fx, hess_fx, jac_yx, grad_fy, hess_fy)):
            raise UnsupportedError(K.SCS_UNSUPPORTED, "user derivative closures cannot run on the GPU; the built-in "
                                                      "losses carry their own out_fn and derivatives")
        self.ctx = ctx or default_context()
        self.f, self.lam, self.L, self.C_set, self.P, self.name = f, lam, L, C_set, P, name
        self.x0 = K.vec(x0)
        self.x = np.zeros_like(self.x0) if sol is None else K.vec(sol, self.x0.shape[0])  # problems.jl:70
        self._h = C.c_void_p()
        self._reg_name = None
        self._host = None  # (A, y) as given: needed again only if iterate!(shuffle_batch=true) reorders the rows
        self.row_order = None  # row permutation the device copy is currently laid out in (None: as given)
        self._storage = storage
        if A is not None and hasattr(A, "tocsc"):  # scipy.sparse (the README builds A with sprandn): CSC over the wire
            Ac = A.tocsc().astype(np.float64)
            Ac.sum_duplicates()
            Ac.sort_indices()
            yv = K.vec(y, Ac.shape[0])
            if Ac.shape[1] != self.x0.shape[0]:
                raise ValueError("x0 length must equal the number of columns of A")
            self.n, self.m = Ac.shape
            cp = np.ascontiguousarray(Ac.indptr, dtype=np.int64)
            rv = np.ascontiguousarray(Ac.indices, dtype=np.int64)
            nz = np.ascontiguousarray(Ac.data, dtype=np.float64)
            self._host = (Ac, yv)
            K.check(K.lib().scs_problem_create_csc(self.ctx._h, K.iptr(cp), K.iptr(rv), K.dptr(nz), 0, self.n, self.m,
                                                   K.dptr(yv), f.kind, f.param(), f.label_code(),
                                                   {"auto": 0, "dense": 1, "sparse": 2}[storage], C.byref(self._h)))
            A = None
        if A is not None:
            A = np.asarray(A, dtype=np.float64)
            if A.ndim != 2:
                raise ValueError("A must be a matrix")
            if not A.flags["F_CONTIGUOUS"]:
                A = np.asfortranarray(A)  # Julia's layout
            yv = K.vec(y, A.shape[0])  # Int / Bool / Float labels all go over the wire as fp64
            if A.shape[1] != self.x0.shape[0]:
                raise ValueError("x0 length must equal the number of columns of A")
            self.n, self.m = A.shape
            self._host = (A, yv)
            K.check(K.lib().scs_problem_create(self.ctx._h, K.dptr(A), self.n, self.m, A.shape[0], K.dptr(yv),
                                               f.kind, f.param(), f.label_code(), C.byref(self._h)))
        # held-out data: a second resident shard, only ever used for ftest(x) (iterate.jl:169-176)
        self._test = None
        if Atest is not None and ytest is not None:
            self._test = Problem(Atest, ytest, self.x0, f, lam, ctx=self.ctx)
        elif (Atest is None) != (ytest is None) and A is not None:
            print("[ Info: Both input (Atest) and target (ytest) data are required for testing the model, but only "
                  "one of these has been provided.\nWill skip testing...")  # iterate.jl:170-171

    def ftest(self, x):
        """model.f(Atest, ytest, x) — iterate.jl:173."""
        return self._test.loss_eval(x, want_grad=False)[0]

    def reorder_rows(self, order):
        """Lay the shard out in the given row order (the data loader's one-time shuffle, utils.jl:18-25): the device
        copy is rebuilt from the ORIGINAL host arrays (every iterate! call shuffles model.A as given, never an earlier
        shuffle), so every mini-batch becomes a contiguous row range.  order=None restores the original order.  The
        order in effect is recorded in `row_order`."""
        if self._host is None:
            raise UnsupportedError(K.SCS_UNSUPPORTED, "a shard generated on the device cannot be shuffled; build the "
                                                      "problem from host arrays or pass shuffle_batch=False")
        A, yv = self._host
        if order is None:
            if self.row_order is None:
                return
            A2, y2 = A, yv
        else:
            order = np.asarray(order, dtype=np.int64)
            y2 = np.ascontiguousarray(yv[order])
            A2 = None
        K.lib().scs_problem_destroy(self._h)
        self._h = C.c_void_p()
        self._reg_name = None
        if hasattr(A, "tocsc"):  # sparse shard: permute the rows of the CSC structure, same storage choice
            if A2 is None:
                A2 = A.tocsr()[order].tocsc()
                A2.sort_indices()
            cp = np.ascontiguousarray(A2.indptr, dtype=np.int64)
            rv = np.ascontiguousarray(A2.indices, dtype=np.int64)
            nz = np.ascontiguousarray(A2.data, dtype=np.float64)
            K.check(K.lib().scs_problem_create_csc(self.ctx._h, K.iptr(cp), K.iptr(rv), K.dptr(nz), 0, self.n, self.m,
                                                   K.dptr(y2), self.f.kind, self.f.param(), self.f.label_code(),
                                                   {"auto": 0, "dense": 1, "sparse": 2}[self._storage],
                                                   C.byref(self._h)))
        else:
            if A2 is None:
                A2 = np.asfortranarray(A[order])
            K.check(K.lib().scs_problem_create(self.ctx._h, K.dptr(A2), self.n, self.m, A2.shape[0], K.dptr(y2),
                                               self.f.kind, self.f.param(), self.f.label_code(), C.byref(self._h)))
        self.row_order = None if order is None else order.copy()
        for name, val in getattr(self, "_modes", {}).items():  # kernel selections survive the rebuild
            getattr(self, name)(val)

    def set_active_rows(self, lo, hi):
        """Rows [lo, hi) of the shard take part in the following passes (one mini-batch); (0, n) = all."""
        K.check(K.lib().scs_set_active_rows(self._h, int(lo), int(hi)))

    def set_batches(self, offsets):
        """Batch table for the in-library loop (scs_solve): local offsets, len = nbatch + 1; None = full batch."""
        if offsets is None:
            K.check(K.lib().scs_set_batches(self._h, 0, None))
            return
        off = np.ascontiguousarray(np.asarray(offsets, dtype=np.int64))
        K.check(K.lib().scs_set_batches(self._h, len(off) - 1, K.iptr(off)))

    @classmethod
    def synthetic(cls, n_total, m, f, lam, *, x0=None, row0=0, n_local=None, seed=1234, density=1.0, ctx=None,
                  storage="dense", **kw):
        """Benchmark-sized shard generated directly in HBM (oracle/synth.py twin)."""
        self = cls(None, None, np.zeros(m) if x0 is None else x0, f, lam, ctx=ctx, **kw)
        n_local = n_total - row0 if n_local is None else n_local
        self.n, self.m = n_local, m
        K.check(K.lib().scs_problem_create_synthetic(self.ctx._h, n_total, row0, n_local, m, f.kind, f.param(),
                                                     f.label_code(), seed, density, C.byref(self._h)))
        if storage == "sparse":  # drop the zeros on the device: CSR + CSC copies, the dense matrix is freed
            K.check(K.lib().scs_problem_sparsify(self._h))
        return self

    # -- helpers -------------------------------------------------------------------------------
    def lam_scalar(self):
        return float(self.lam[0]) if np.ndim(self.lam) > 0 and len(self.lam) > 1 else float(np.ravel(self.lam)[0])

    def read_rows(self, row0, nrows):
        A = np.empty((nrows, self.m), order="F")
        y = np.empty(nrows)
        K.check(K.lib().scs_problem_read_rows(self._h, row0, nrows, K.dptr(A), K.dptr(y)))
        return A, y

    def _set_reg(self, reg_name):
        if reg_name not in K.REG_KINDS:
            raise ScsError(K.SCS_INVALID_ARG, "reg_name not valid.")  # prox-operators.jl:78, regularizers.jl:29
        kind = K.REG_KINDS[reg_name]
        lam1 = lam2 = 0.0
        ind = perm = lb = ub = None
        ng = nlb = nub = 0
        if reg_name == "gl":
            if np.ndim(self.lam) == 0 or len(self.lam) != 2:
                raise ScsError(K.SCS_INVALID_ARG,
                               "Please provide a Tuple or Vector with exactly two entries for λ, e.g. [λ1, λ2]")
            if self.P is None:
                raise ScsError(K.SCS_INVALID_ARG, "reg_name \"gl\" needs P = get_P(n, G, ind)")
            lam1, lam2 = float(self.lam[0]), float(self.lam[1])
            ind, ng = self.P.ind, self.P.grpNUM
            perm = self.P.G if not np.array_equal(self.P.G, np.arange(1, self.m + 1)) else None
        else:
            lam1 = self.lam_scalar()
        if reg_name == "indbox":
            if self.C_set is None:
                raise ScsError(K.SCS_INVALID_ARG, "reg_name \"indbox\" needs C_set = (lb, ub)")
            lb, ub = K.vec(self.C_set[0]), K.vec(self.C_set[1])
            nlb, nub = lb.shape[0], ub.shape[0]
        K.check(K.lib().scs_set_regularizer(self._h, kind, lam1, lam2, K.iptr(ind), ng, K.iptr(perm), K.dptr(lb), nlb,
                                            K.dptr(ub), nub))
        self._reg_name = reg_name

    def _set_smoother(self, h):
        lb = ub = None
        nlb = nub = 0
        if getattr(h, "lb", None) is not None:
            lb, ub = K.vec(h.lb), K.vec(h.ub)
            nlb, nub = lb.shape[0], ub.shape[0]
        K.check(K.lib().scs_set_smoother(self._h, h.kind, float(h.mu), K.dptr(lb), nlb, K.dptr(ub), nub))

    def _set_method(self, method):
        K.check(K.lib().scs_set_method(self._h, method.kind, int(method.ss_type), int(bool(method.use_prox)),
                                       int(getattr(method, "m", 10))))
        K.check(K.lib().scs_set_L(self._h, int(self.L is not None), float(self.L) if self.L is not None else 0.0))

    def configure(self, method, reg_name, hmu):
        self._set_reg(reg_name)
        self._set_smoother(hmu)
        if method is not None:
            self._set_method(method)

    # -- the two call sites of optim_loop! ----------------------------------------------------------
    def objective(self, x):
        """(model.f(A,y,x), get_reg(model,x,reg_name))  — iterate.jl:189-190."""
        fv, rv = C.c_double(), C.c_double()
        K.check(K.lib().scs_objective(self._h, K.dptr(K.vec(x, self.m)), C.byref(fv), C.byref(rv)))
        return fv.value, rv.value

    def step(self, x, x_prev, it, return_dx=False):
        """step!(method, model, reg_name, hμ, As, x, x_prev, ys, Cmat, iter; return_dx) — iterate.jl:233."""
        x = K.vec(x, self.m)
        xp = K.vec(x_prev, self.m) if x_prev is not None else None
        xn = np.empty(self.m)
        dx = np.empty(self.m) if return_dx else None
        pri = C.c_double()
        K.check(K.lib().scs_step(self._h, K.dptr(x), K.dptr(xp), int(it), K.dptr(xn), K.dptr(dx), C.byref(pri)))
        return (xn, dx, pri.value) if return_dx else (xn, pri.value)

    # -- component entry points -----------------------------------------------------------------------
    def loss_eval(self, x, weights="newton", want_grad=True, want_rows=False):
        wk = K.WEIGHTS_GGN if weights == "ggn" else K.WEIGHTS_NEWTON
        fv = C.c_double()
        g = np.empty(self.m) if want_grad else None
        z = r = w = None
        if want_rows:
            z, r, w = np.empty(self.n), np.empty(self.n), np.empty(self.n)
        K.check(K.lib().scs_loss_eval(self._h, K.dptr(K.vec(x, self.m)), wk, C.byref(fv), K.dptr(g), K.dptr(z),
                                      K.dptr(r), K.dptr(w)))
        return fv.value, g, z, r, w

    def gram(self, x, weights="newton"):
        wk = K.WEIGHTS_GGN if weights == "ggn" else K.WEIGHTS_NEWTON
        G = np.empty((self.m, self.m), order="F")
        K.check(K.lib().scs_gram(self._h, K.dptr(K.vec(x, self.m)), wk, K.dptr(G)))
        return G

    def smoother_eval(self, x):
        gr, hr = np.empty(self.m), np.empty(self.m)
        K.check(K.lib().scs_smoother_eval(self._h, K.dptr(K.vec(x, self.m)), K.dptr(gr), K.dptr(hr)))
        return gr, hr

    def prox(self, u, hr, ss):
        out = np.empty(self.m)
        K.check(K.lib().scs_prox(self._h, K.dptr(K.vec(u, self.m)), K.dptr(K.vec(hr, self.m)), float(ss), K.dptr(out)))
        return out

    def set_gram_mode(self, mode):
        """"auto" | "dmma" | "i8": which Gram kernel builds A'diag(w)A (the int8 tcgen05 path needs w >= 0)."""
        K.check(K.lib().scs_set_gram_mode(self._h, {"auto": 0, "dmma": 1, "i8": 2}[mode]))
        self.__dict__.setdefault("_modes", {})["set_gram_mode"] = mode

    def gram_path(self):
        v = C.c_int()
        K.check(K.lib().scs_get_gram_path(self._h, C.byref(v)))
        return {0: None, 1: "dmma", 2: "i8", 3: "sparse"}[v.value]

    def is_sparse(self):
        """(resident in sparse form?, stored entries)"""
        a, b = C.c_int(), C.c_int64()
        K.check(K.lib().scs_problem_is_sparse(self._h, C.byref(a), C.byref(b)))
        return bool(a.value), b.value

    def set_gram_bits(self, bits):
        """log2 of the common column 2-norm T of the fixed-point image the emulated-fp64 Gram works on (24..58, default 46:
        Gram entries are off by ~0.4/T of the diagonal scale; see include/scs_b200.h)."""
        K.check(K.lib().scs_set_gram_bits(self._h, int(bits)))
        self.__dict__.setdefault("_modes", {})["set_gram_bits"] = bits

    def gram_info(self):
        """(moduli used, bits kept) by the emulated-fp64 Gram; (0, 0) before it has run."""
        a, b = C.c_int(), C.c_int()
        K.check(K.lib().scs_get_gram_info(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def gram_signed(self):
        """How the last emulated-fp64 Gram handled the weight signs: (rows compacted into the second plane set, or -1 when
        the weights had one sign; True if the compacted rows were the negative ones)."""
        a, b = C.c_int64(), C.c_int()
        K.check(K.lib().scs_get_gram_signed(self._h, C.byref(a), C.byref(b)))
        return a.value, bool(b.value)

    def set_stream_mode(self, mode):
        """"auto" | "two_pass" | "fused": how objective + gradient at the same x read A (once or twice)."""
        K.check(K.lib().scs_set_stream_mode(self._h, {"auto": 0, "two_pass": 1, "fused": 2}[mode]))
        self.__dict__.setdefault("_modes", {})["set_stream_mode"] = mode

    def stream_path(self):
        v = C.c_int()
        K.check(K.lib().scs_get_stream_path(self._h, C.byref(v)))
        return {0: None, 1: "two_pass", 2: "fused"}[v.value]

    def reg_value(self, x):
        v = C.c_double()
        K.check(K.lib().scs_reg_value(self._h, K.dptr(K.vec(x, self.m)), C.byref(v)))
        return v.value

    def close(self):
        if getattr(self, "_test", None) is not None:
            self._test.close()
            self._test = None
        if getattr(self, "_h", None):
            K.lib().scs_problem_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def ProblemGeneric(*a, **k):
    """Problem(x0, f, λ; ...) without data (src/problems.jl:44-59): arbitrary f(x), ForwardDiff only."""
    raise UnsupportedError(K.SCS_UNSUPPORTED,
                           "ProblemGeneric (arbitrary f(x) differentiated by ForwardDiff) has no GPU path and is "
                           "rejected instead of running on the CPU")


# ---- smoothers (constants pinned by test/test_smooth.jl) ------------------------------------------
class _Smoother:
    lb = ub = None

    def __init__(self, mu, Mh, nu):
        self.mu, self.Mh, self.nu = mu, Mh, nu


class PHuberSmootherL1L2(_Smoother):  # phuber-smooth.jl:6-27
    kind = K.SMOOTH_PHUBER_L1L2

    def __init__(self, mu):
        super().__init__(mu, 2.0, 2.6)


class PHuberSmootherIndBox(_Smoother):  # phuber-smooth.jl:38-64
    kind = K.SMOOTH_PHUBER_INDBOX

    def __init__(self, lb, ub, mu):
        super().__init__(mu, 2.0, 2.6)
        self.lb, self.ub = lb, ub


class PHuberSmootherGL(_Smoother):  # phuber-smooth.jl:116-148 (takes the problem, reads model.P)
    kind = K.SMOOTH_PHUBER_GL

    def __init__(self, mu, model):
        super().__init__(mu, 2.0, 2.6)
        self.model = model


class ExponentialSmootherIndBox(_Smoother):  # exponential-smooth.jl:28-34
    kind = K.SMOOTH_EXP_INDBOX

    def __init__(self, lb, ub, mu):
        super().__init__(mu, 1.0, 2.0)
        self.lb, self.ub = lb, ub


class LogExpSmootherIndBox(_Smoother):  # log-exp-smooth.jl:28-34
    kind = K.SMOOTH_LOGEXP_INDBOX

    def __init__(self, lb, ub, mu):
        super().__init__(mu, 1.0, 2.0)
        self.lb, self.ub = lb, ub


class OsBaSmootherL1L2(_Smoother):  # ostrovskii-bach-smooth.jl:6-27
    kind = K.SMOOTH_OSBA_L1L2

    def __init__(self, mu):
        super().__init__(mu, 2 * np.sqrt(2), 3.0)


class OsBaSmootherGL(_Smoother):  # ostrovskii-bach-smooth.jl:38-70
    kind = K.SMOOTH_OSBA_GL

    def __init__(self, mu, model):
        super().__init__(mu, 2 * np.sqrt(2), 3.0)
        self.model = model


# ---- methods -----------------------------------------------------------------------------------
@dataclass
class ProxNSCORE:  # prox-N-SCORE.jl:6-33
    ss_type: int = 1
    use_prox: bool = True
    name: str = "prox-newtonscore"
    label: str = "Prox-N-SCORE"
    kind = K.METHOD_N

    def set_name(self):
        if not self.use_prox:
            self.name, self.label = "newtonscore", "Newton-SCORE"


@dataclass
class ProxGGNSCORE:  # prox-GGN-SCORE.jl:6-33
    ss_type: int = 1
    use_prox: bool = True
    name: str = "prox-ggnscore"
    label: str = "Prox-GGN-SCORE"
    kind = K.METHOD_GGN

    def set_name(self):
        if not self.use_prox:
            self.name, self.label = "ggnscore", "GGN-SCORE"


@dataclass
class ProxLQNSCORE:  # prox-L-BFGS-SCORE.jl:6-46 (s_list / y_list / H0 live on the device)
    ss_type: int = 1
    use_prox: bool = True
    m: int = 10
    name: str = "prox-lbfgsscore"
    label: str = "Prox-LBFGS-SCORE"
    kind = K.METHOD_LQN

    def set_name(self):
        if not self.use_prox:
            self.name, self.label = "lbfgsscore", "LBFGS-SCORE"


@dataclass
class Solution:  # iterate.jl:3-32
    x: np.ndarray
    obj: list
    fval: list
    pri_res_norm: list
    fvaltest: list
    rel: list
    objrel: list
    metricvals: dict
    times: list
    epochs: int
    model: object


def _norm(v):
    return float(np.linalg.norm(v))


def batch_plan(n, batch_size=None, slice_samples=False, shuffle_batch=False, local_max_iter=None, perm=None):
    """(row order, batch offsets) of optim_loop!'s data loader (iterate.jl:122-145, utils.jl:14-25) for n rows.

    MLUtils.DataLoader(batchsize, shuffle, partial=true): ceil(n/b) consecutive batches of the once-shuffled order,
    collected ONCE before the epoch loop; slice_samples steps on the FIRST row only (the loader subset is 1:iend with
    iend = 1, see below; batch_size wins if both are given); local_max_iter keeps the first
    min(floor(local_max_iter), max_iter) batches (and makes iterate! run a single epoch).  order is None when the rows stay
    where they are.  The shuffle itself is Julia's RNG upstream: here the permutation is an input (perm), or a
    fixed-seed numpy permutation when omitted."""
    if batch_size is not None and slice_samples:
        slice_samples = False  # iterate.jl:127-130
    if batch_size is not None and int(batch_size) < 1:
        raise ScsError(K.SCS_INVALID_ARG, "batch_size must be positive")
    # max_iter and iend are computed before slice_samples sets opt.batch_size = 1 (iterate.jl:124-127 vs :136-138): with
    # no batch_size max_iter is 1, so the loader subset 1:iend (:145) is ONE entry — for slice_samples the first row.
    max_iter = -(-n // int(batch_size)) if batch_size is not None else 1
    iend = max_iter
    if local_max_iter is not None and int(np.floor(local_max_iter)) > 0:
        iend = min(int(np.floor(local_max_iter)), max_iter)
    if slice_samples:
        batch_size, shuffle_batch = 1, False
    if batch_size is None:
        batch_size, shuffle_batch = n, False
    batch_size = int(batch_size)
    order = None
    if shuffle_batch:
        order = np.random.default_rng(1234).permutation(n) if perm is None else np.asarray(perm, dtype=np.int64)
        if order.shape != (n,) or not np.array_equal(np.sort(order), np.arange(n)):
            raise ScsError(K.SCS_INVALID_ARG, "perm must be a permutation of 0..n-1")
    offsets = np.minimum(np.arange(iend + 1, dtype=np.int64) * batch_size, n)
    return order, offsets


def iterate(method, model, reg_name, hmu, *, metrics=None, alpha=None, batch_size=None, slice_samples=False,
            shuffle_batch=True, max_epoch=1000, comm_rounds=100, local_max_iter=None, x_tol=1e-10, f_tol=1e-10,
            verbose=1, device_loop=False, perm=None, batch_offsets=None):
    """iterate!(method, model, reg_name, hμ; ...) — iterate.jl:56-76 → optim_loop! :100-266.

    device_loop=False: the epoch loop runs here and calls scs_objective / scs_step (what the Julia shim does).
    device_loop=True: the same loop runs inside the library (scs_solve), x never leaves HBM.
    Mini-batches: the objective is taken over all rows every epoch, step! over one batch at a time (:204-233).  On one
    rank the batches are laid out here (rows are re-uploaded once in shuffled order when shuffle_batch is set); with
    several ranks the caller shards every batch over the ranks and passes this rank's local offsets (batch_offsets).
    """
    import time
    if local_max_iter is not None:  # Options(max_epoch = local_max_iter !== nothing ? 1 : max_epoch), iterate.jl:58-70
        max_epoch = 1
    if metrics is not None:
        if device_loop:
            raise UnsupportedError(K.SCS_UNSUPPORTED, "metrics callbacks need x on the host every epoch: use the host "
                                                      "loop (device_loop=False)")
        if not all(callable(fn) for fn in metrics.values()):
            raise ScsError(K.SCS_INVALID_ARG, "metrics must map names to callables (model, x) -> value")
    offsets = None
    if batch_offsets is not None:
        offsets = np.asarray(batch_offsets, dtype=np.int64)
    elif batch_size is not None or slice_samples:
        if model.ctx.world > 1:
            raise UnsupportedError(K.SCS_UNSUPPORTED, "with several ranks pass batch_offsets (see batch_shard)")
        order, offsets = batch_plan(model.n, batch_size, slice_samples, shuffle_batch, local_max_iter, perm)
        if order is not None:
            model.reorder_rows(order)
        elif model.row_order is not None:
            model.reorder_rows(None)  # an earlier shuffled call left the device copy permuted
    method.set_name()  # iterate.jl:112
    if alpha is not None:
        model.L = 1 / alpha  # :113-115
    model.configure(method, reg_name, hmu)
    x_star = model.x
    m = model.m
    if device_loop:
        cap = int(max_epoch) + 2
        xo = np.empty(m)
        h = [np.empty(cap) for _ in range(5)]
        nh, ep = C.c_int64(), C.c_int64()
        model.set_batches(offsets)
        test = getattr(model, "_test", None)
        K.check(K.lib().scs_set_test_problem(model._h, test._h if test is not None else None))
        try:
            K.check(K.lib().scs_solve(model._h, K.dptr(model.x0), K.dptr(x_star), int(max_epoch), float(x_tol),
                                      float(f_tol), K.dptr(xo), *[K.dptr(a) for a in h], C.byref(nh), C.byref(ep)))
        finally:
            model.set_batches(None)
        k = nh.value
        ftests = []
        if test is not None:
            buf, cnt = np.empty(cap), C.c_int64()
            K.check(K.lib().scs_get_test_history(model._h, K.dptr(buf), cap, C.byref(cnt)))
            ftests = list(buf[:cnt.value])
        pri = [None if np.isnan(v) else float(v) for v in h[2][:k]]
        return Solution(xo, list(h[0][:k]), list(h[1][:k]), pri, ftests, list(h[3][:k]), list(h[4][:k]), {}, [],
                        ep.value, model)

    objs, fvals, pris, rels, frels, times = [], [], [], [], [], []
    epochs = 0
    pri_res_norm = None
    batched = offsets is not None
    windows = [(int(offsets[i]), int(offsets[i + 1])) for i in range(len(offsets) - 1)] if batched else [(0, model.n)]
    iend = len(windows)

    def objective(v):  # always over the whole data (iterate.jl:168,189)
        if batched:
            model.set_active_rows(0, model.n)
        return model.objective(v)

    fs, rs = objective(x_star)
    with np.errstate(all="ignore"):
        obj_star = fs + rs  # :179
    x = model.x0.copy()
    x_prev = x.copy()
    K.check(K.lib().scs_method_init(model._h))  # init!(method, x) :183
    t0 = time.perf_counter()

    def rel_err(v):  # :192-197
        if reg_name == "gl":
            return float(np.mean((x_star - v) ** 2))
        return max(_norm(v - x_star) / max(_norm(x_star), 1), x_tol)

    def frel(o):  # :200
        with np.errstate(all="ignore"):
            return float(np.maximum(np.abs(np.float64(o) - obj_star) / np.abs(np.float64(obj_star)), f_tol))

    ftests = []
    has_test = getattr(model, "_test", None) is not None
    metric_vals = {name: [] for name in metrics} if metrics is not None else {}
    if verbose > 0 and method.ss_type == 1 and model.L is None:  # iterate.jl:116-120
        print("[ Info: Neither L nor α is set for the problem... Now fixing α = 0.5...")

    def push(o, f, p, r, fr, v=None, tag="epoch", ep=0):  # show_stat! + update_stat! (utils.jl:50-113)
        dt = time.perf_counter() - t0
        objs.append(o), fvals.append(f), pris.append(p), rels.append(r), frels.append(fr)
        times.append(dt)
        ft = None
        if has_test:
            ft = model.ftest(v)
            ftests.append(ft)
        for name, fn in (metrics or {}).items():  # utils.jl:80-83: metrics[name](model, x) at every recorded state
            metric_vals[name].append(fn(model, v))
        if verbose > 1:  # utils.jl:51-78
            print("\n" + "=" * 30 + f"\nOptimizer:\t{method.label}")
            print(f"{tag} = {ep}\nobj = {o}\nfval = {f}\npri_res_norm = {p}")
            if has_test:
                print(f"fvaltest = {ft}")
            print(f"rel_error = {r}\nΔtime = {dt}")
            for name in metric_vals:
                print(f"{name}:\t{metric_vals[name][-1]}")

    try:
        for epoch_t in range(1, int(max_epoch) + 1):
            fval, reg = objective(x)
            obj = fval + reg
            rel_error = rel_err(x)
            f_rel_error = frel(obj)
            push(obj, fval, pri_res_norm, rel_error, f_rel_error, x, "epoch", epoch_t - 1)
            for i, (lo, hi) in enumerate(windows, start=1):  # :204
                if epoch_t == max_epoch and i == iend:  # :219-231
                    fval, reg = objective(x)
                    obj = fval + reg
                    f_rel_error = frel(obj)
                    push(obj, fval, pri_res_norm, rel_err(x), f_rel_error, x, "max_epoch", epoch_t)
                if batched:
                    model.set_active_rows(lo, hi)
                x_new, pri_res_norm = model.step(x, x_prev, epoch_t)  # :233
                if _norm(x_new - x) < x_tol * max(_norm(x), 1) or f_rel_error <= f_tol or pri_res_norm < x_tol:  # :234
                    if epoch_t != max_epoch:
                        fval, reg = objective(x_new)
                        obj = fval + reg
                        f_rel_error = frel(obj)
                        push(obj, fval, pri_res_norm, rel_err(x_new), f_rel_error, x_new, "terminate_epoch", epoch_t)
                    x_prev, x = x.copy(), x_new
                    epochs += 1
                    break
                x_prev, x = x.copy(), x_new
            if _norm(x - x_prev) < x_tol * max(_norm(x_prev), 1) or f_rel_error <= f_tol or pri_res_norm < x_tol:  # :257
                break
            epochs += 1
    finally:
        if batched:
            model.set_active_rows(0, model.n)
    return Solution(x, objs, fvals, pris, ftests, rels, frels, metric_vals, times, epochs, model)


def batch_shard(n, world, rank, batch_size, local_max_iter=None, perm=None):
    """Row ids (global, in upload order) and local batch offsets of one rank when every mini-batch of the global
    problem is split evenly over the ranks: rank r holds rows shard_rows(len(batch), world, r) of each batch, batch
    after batch, followed by its share of the rows no batch uses (local_max_iter), so that the objective still sees
    every row."""
    from .dist import shard_rows
    order, offsets = batch_plan(n, batch_size, False, perm is not None, local_max_iter, perm)
    order = np.arange(n) if order is None else order
    rows, loc = [], [0]
    bounds = list(offsets) + ([n] if offsets[-1] < n else [])
    for i in range(len(bounds) - 1):
        b0, b1 = int(bounds[i]), int(bounds[i + 1])
        r0, cnt = shard_rows(b1 - b0, world, rank)
        rows.append(order[b0 + r0:b0 + r0 + cnt])
        if i < len(offsets) - 1:
            loc.append(loc[-1] + cnt)
    return np.concatenate(rows), np.asarray(loc, dtype=np.int64)
