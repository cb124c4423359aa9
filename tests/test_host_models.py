"""CPU models of two device-side schemes whose correctness argument is arithmetic, not timing (no GPU needed):

* the sparse Gram's fixed-point accumulation (csrc/kernels_sparse.cuh k_sp_gram): every addend is rint(x * 2^61) of a
  value scaled so that all partial sums stay below 1 in magnitude; integer addition is order-independent, and the error
  is at most nnz_col * 2^-62 of the diagonal scale;
* the slice partition of the peer-memory all-reduce (csrc/kernels_i8gram.cuh p2p_slice): the slices of the W ranks tile
  [0, count) exactly once, start on even offsets (128-bit loads), and only the last one may have an odd tail;
plus the host-side pieces of bench.py's reference arm (row sample >= 5 %, explicit BLAS thread count)."""
import itertools
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sparse_gram_fixed_point_is_order_independent_and_accurate():
    rng = np.random.default_rng(3)
    n, m, dens = 400, 12, 0.3
    A = rng.standard_normal((n, m)) * (rng.random((n, m)) < dens) * rng.uniform(0.01, 50.0, m)
    w = rng.uniform(0.01, 0.25, n) * np.where(rng.random(n) < 0.2, -1.0, 1.0)  # signed weights are fine
    cs = 1.0 / np.sqrt(np.abs(w) @ (A * A))  # k_sp_colscale: 1 / sqrt(sum_i |w_i| a_ij^2)
    FIX = 2.0 ** 61
    Gref = A.T @ (w[:, None] * A)
    d = np.sqrt(np.diag(A.T @ (np.abs(w)[:, None] * A)))
    for k in range(m):
        rows = np.nonzero(A[:, k])[0]
        for j in range(k, m):
            add = [int(np.rint((w[i] * A[i, k] * (cs[k] * FIX)) * (A[i, j] * cs[j]))) for i in rows if A[i, j] != 0.0]
            # bounded: every partial sum fits a signed 64-bit accumulator (|sum| <= 2^61 by Cauchy-Schwarz on |.|)
            assert sum(abs(a) for a in add) <= 2 ** 61 * (1 + 1e-9)
            tot = sum(add)
            perm = rng.permutation(len(add))
            assert sum(add[p] for p in perm) == tot  # integers: any interleaving of the warps gives the same bits
            g = tot / (cs[j] * cs[k] * FIX)
            assert abs(g - Gref[j, k]) <= (len(add) * 2.0 ** -62 + 4e-16) * d[j] * d[k]
    # the two-halves accumulator (low word + carry into the high word) is the same 64-bit sum
    vals = [int(v) for v in rng.integers(-2 ** 60, 2 ** 60, 50)]
    lo = hi = 0
    for v in vals:
        u = v & (2 ** 64 - 1)
        l, h = u & 0xFFFFFFFF, u >> 32
        old = lo
        lo = (lo + l) & 0xFFFFFFFF
        hi = (hi + h + (1 if lo < l else 0)) & 0xFFFFFFFF
        assert (old + l >= 2 ** 32) == (lo < l)
    acc = (hi << 32) | lo
    acc = acc - 2 ** 64 if acc >= 2 ** 63 else acc
    assert acc == sum(vals)


def _p2p_slice(count, r, world):
    lo = (count * r // world) & ~1
    hi = count if r == world - 1 else (count * (r + 1) // world) & ~1
    return lo, hi


@pytest.mark.parametrize("count,world", list(itertools.product([1, 2, 7, 8, 1001, 8390656, 33558528], [2, 3, 4, 8, 16])))
def test_p2p_slices_tile_the_packed_triangle(count, world):
    prev = 0
    for r in range(world):
        lo, hi = _p2p_slice(count, r, world)
        assert lo == prev and hi >= lo and lo % 2 == 0
        if r < world - 1:
            assert (hi - lo) % 2 == 0
        prev = hi
    assert prev == count


def test_reference_arm_sample_is_at_least_five_percent():
    sys.path.insert(0, ROOT)
    import bench
    for name, wl in bench.WORKLOADS.items():
        if "density" in wl:
            continue
        n_s = bench.cpu_sample_rows(wl)
        assert n_s >= 0.05 * wl["n"] and n_s <= wl["n"], (name, n_s)
    assert bench.cpu_sample_rows(bench.WORKLOADS["c2"]) % 1024 == 0


def test_blas_thread_count_is_set_explicitly_even_under_torchrun_env():
    """torch.distributed.run exports OMP_NUM_THREADS=1 for nproc > 1; the reference arm must still use all usable cores."""
    code = ("import sys; sys.path.insert(0, %r); import bench, threadpoolctl; ctl, got = bench.blas_threads(bench.host_cores());\n"
            "import numpy as np\n"
            "with ctl:\n"
            "    now = max(int(p['num_threads']) for p in threadpoolctl.threadpool_info())\n"
            "print(got, now, bench.host_cores())" % ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True,
                         env=dict(os.environ, OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1"))
    assert out.returncode == 0, out.stderr[-2000:]
    got, now, cores = map(int, out.stdout.split())
    assert got == now == cores or cores > 64  # OpenBLAS builds cap their pool (64 / 128 threads)


def _chol_schedule(M, pair):
    """numpy model of chol_enqueue's look-ahead schedule (scs_lib.cu): which kernel touches which 64-block when.
    diag(k) applies `nprev` pending panels to its own tile and factors it; trsm(k) solves the panel below; the trailing
    update runs with tile0 = 1 (never the first diagonal tile), either every step (pair = False) or — paired — on the
    next block column only after even steps and on everything with both panels after odd steps."""
    NB = 64
    M = np.tril(M).copy()
    m = M.shape[0]
    nblk = (m + NB - 1) // NB
    blk = lambda i: slice(i * NB, min(m, (i + 1) * NB))
    for k in range(nblk):
        nprev = 0 if k == 0 else (2 if (pair and k % 2 == 0) else 1)
        for h in range(nprev, 0, -1):
            P = M[blk(k), blk(k - h)]
            M[blk(k), blk(k)] -= np.tril(P @ P.T)
        L = np.linalg.cholesky(M[blk(k), blk(k)] + np.tril(M[blk(k), blk(k)], -1).T)
        M[blk(k), blk(k)] = L
        if k == nblk - 1:
            break
        M[(k + 1) * NB:, blk(k)] = np.linalg.solve(L, M[(k + 1) * NB:, blk(k)].T).T
        rb = nblk - 1 - k
        colonly = pair and k % 2 == 0
        kpan = 2 if (pair and k % 2 == 1) else 1
        tiles = [(ti, 0) for ti in range(1, rb)] if colonly else \
            [(ti, tj) for ti in range(rb) for tj in range(ti + 1) if (ti, tj) != (0, 0)]
        for ti, tj in tiles:
            for h in range(kpan - 1, -1, -1):
                Pi = M[blk(k + 1 + ti), blk(k - h)]
                Pj = M[blk(k + 1 + tj), blk(k - h)]
                upd = Pi @ Pj.T
                M[blk(k + 1 + ti), blk(k + 1 + tj)] -= np.tril(upd) if ti == tj else upd
    return M


@pytest.mark.parametrize("m", [64, 65, 128, 129, 192, 200, 256, 321, 448])
@pytest.mark.parametrize("pair", [False, True])
def test_lookahead_cholesky_schedule_applies_every_panel_exactly_once(m, pair):
    rng = np.random.default_rng(m)
    B = rng.standard_normal((m + 5, m))
    G = B.T @ B + np.eye(m)
    L = _chol_schedule(G, pair)
    assert np.allclose(L, np.linalg.cholesky(G), rtol=0, atol=1e-11 * np.abs(G).max())


def test_clock_sampler_parses_power_and_throttle_reasons():
    """bench.ClockSampler: median SM clock and power, the enforced limit, and the set of active throttle reasons from
    nvidia-smi's CSV rows ([N/A] fields and short rows are skipped, not fatal)."""
    sys.path.insert(0, ROOT)
    import bench
    c = bench.ClockSampler(0)
    c.rows = [["1586", "1965", "987.30", "Not Active", "Not Active", "Not Active", "Active", "1000.00"],
              ["1590", "1965", "[N/A]", "Not Active", "Not Active", "Not Active", "Active", "1000.00"],
              ["1605", "1965", "995.10", "Not Active", "Not Active", "Not Active", "Not Active", "1000.00"],
              ["garbage"]]
    out = c.stop()
    assert out["sm_mhz"] == 1590.0 and out["sm_max_mhz"] == 1965.0 and out["samples"] == 3
    assert out["reasons"] == ["sw_power_cap"]
    assert abs(out["power_w"] - 991.2) < 1e-9 and out["power_limit_w"] == 1000.0
