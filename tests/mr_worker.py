"""Worker for tests/test_gpu_multirank.py — launched by torch.distributed.run, one rank per GPU.  Each rank uploads
its contiguous row block; results must match the full-batch oracle within the 1e-10 bar on every rank."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "selfconcordantsmoothoptimization.jl_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402
import scs_b200 as S  # noqa: E402
from oracle import scs_oracle as O  # noqa: E402


def _gram_exchange_checks(ctx, rank, world):
    """The Gram exchange step itself, on a shape that takes the slab-pipelined packed-triangle all-reduce (m >= 2048: four
    slabs; emulated-fp64 Gram forced) and on the plain full-square all-reduce of the DMMA path: the all-reduced matrix on
    every rank against A' diag(w) A of the whole problem in fp64, and bitwise identical across the ranks."""
    from oracle import synth
    n, m = 9000, 2100
    A = synth.make_A(n, m, seed=21)
    y = synth.make_labels_logistic(A @ synth.make_x_true(m, seed=22, frac=0.2), seed=23)
    x = synth.make_x0(m, seed=24) * 0.4
    Lo = O.LogisticLoss(1 / n, "consistent")
    w = Lo.ggn_weights(A @ x, y)[1]
    Gref = A.T @ (w[:, None] * A)
    d = np.sqrt(np.diag(Gref))
    r0, nl = S.shard_rows(n, world, rank)
    model = S.Problem(A[r0:r0 + nl], y[r0:r0 + nl], x, S.LogisticLoss(1 / n, "consistent"), 1e-2, ctx=ctx)
    worst = 0.0
    for mode in ("i8", "dmma"):
        model.set_gram_mode(mode)
        G = model.gram(x, weights="ggn")
        assert model.gram_path() == mode
        err = float(np.max(np.abs(G - Gref) / np.outer(d, d)))
        assert err <= 2e-12, (mode, rank, err)
        assert np.array_equal(G, G.T)
        t = torch.from_numpy(np.ascontiguousarray(G)).cuda()
        lst = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(lst, t)
        assert all(torch.equal(lst[0], u) for u in lst), mode
        worst = max(worst, err)
    model.close()
    # weights of both signs (the literal +-1 cross-entropy pair): every rank compacts its own minority rows, the signed
    # integer Grams are reconstructed per rank and summed by the same exchange step
    wl = O.LogisticLoss(1 / n).ggn_weights(A @ x, y)[1]
    assert np.any(wl < 0)
    Gl = A.T @ (wl[:, None] * A)
    dl = np.sqrt(np.diag(A.T @ (np.abs(wl)[:, None] * A)))
    model = S.Problem(A[r0:r0 + nl], y[r0:r0 + nl], x, S.LogisticLoss(1 / n, "literal"), 1e-2, ctx=ctx)
    model.set_gram_mode("i8")
    G = model.gram(x, weights="ggn")
    assert model.gram_path() == "i8" and model.gram_signed()[0] > 0
    err = float(np.max(np.abs(G - Gl) / np.outer(dl, dl)))
    assert err <= 2e-12, ("signed", rank, err)
    worst = max(worst, err)
    model.close()
    return worst


def run_checks(ctx, rank, world, quick=False):
    """Multi-rank parity against the full-batch oracle; torch.distributed must be initialised (NCCL).  Returns the worst
    relative error seen.  quick: two configurations only (bench.py --selftest)."""
    worst = 0.0
    names = ("c2_logreg_ggn_l1", "c3_logreg_lqn_l1") if quick else ("c2_logreg_ggn_l1", "c3_logreg_lqn_l1", "c4_ls_ggn_gl", "c5_ls_n_indbox")
    worst = _full_batch_checks(ctx, rank, world, names)
    worst = max(worst, _gram_exchange_checks(ctx, rank, world))
    if not quick:
        worst = max(worst, _minibatch_checks(ctx, rank, world))
    return worst


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = S.context_from_env()
    worst = run_checks(ctx, rank, world)
    dist.barrier()
    if rank == 0:
        print(f"MULTIRANK_OK world={world} worst_rel_err={worst:.3e}")
    dist.destroy_process_group()


def _full_batch_checks(ctx, rank, world, names):
    worst = 0.0
    for name in names:
        A, y, x0 = cases.data(name)
        mo, modelo, reg, ho, kw = cases.build(name, O)
        so = O.iterate(mo, modelo, reg, ho, **kw)
        r0, nl = S.shard_rows(A.shape[0], world, rank)
        # loss scale / denominator must stay the full-batch one: build with the full data for the constants,
        # then hand this rank's rows to the GPU problem
        mg, full_model_unused, reg, hg, kw = None, None, reg, None, kw
        n = A.shape[0]
        lib = S
        if name == "c2_logreg_ggn_l1":
            model = lib.Problem(A[r0:r0 + nl], y[r0:r0 + nl], x0, lib.LogisticLoss(1 / n, "consistent"), 1e-2, ctx=ctx)
            mg, hg = lib.ProxGGNSCORE(), lib.PHuberSmootherL1L2(1.0)
        elif name == "c3_logreg_lqn_l1":
            model = lib.Problem(A[r0:r0 + nl], y[r0:r0 + nl], x0, lib.LogisticLoss(1 / n), 1e-2, ctx=ctx)
            mg, hg = lib.ProxLQNSCORE(m=10), lib.PHuberSmootherL1L2(1.0)
        elif name == "c4_ls_ggn_gl":
            m = A.shape[1]
            ng = m // 64
            ind = np.array([[g * 64 + 1 for g in range(ng)], [(g + 1) * 64 for g in range(ng)], [1] * ng])
            model = lib.Problem(A[r0:r0 + nl], y[r0:r0 + nl], x0, lib.LeastSquaresLoss(n), [1e-8, 1e-2],
                                P=lib.get_P(m, np.arange(1, m + 1), ind), ctx=ctx)
            mg, hg = lib.ProxGGNSCORE(), lib.PHuberSmootherGL(1e-2, model)
        else:
            model = lib.Problem(A[r0:r0 + nl], y[r0:r0 + nl], x0, lib.LeastSquaresLoss(n), 1e-4, C_set=(-0.5, 0.5), ctx=ctx)
            mg, hg = lib.ProxNSCORE(), lib.PHuberSmootherIndBox(-0.5, 0.5, 0.6)
        for dl in (False, True):
            sg = S.iterate(mg, model, reg, hg, verbose=0, device_loop=dl, **kw)
            ex = np.linalg.norm(sg.x - so.x) / np.linalg.norm(so.x)
            eo = max(abs(a - b) / abs(b) for a, b in zip(sg.obj, so.obj) if np.isfinite(b))
            assert sg.epochs == so.epochs, (name, sg.epochs, so.epochs)
            assert ex <= 1e-10 and eo <= 1e-10, (name, rank, ex, eo)
            assert np.array_equal(sg.x != 0, so.x != 0)
            worst = max(worst, ex, eo)
        # every rank must hold bitwise the same iterate (replicated x-update on identical all-reduced data)
        t = torch.from_numpy(sg.x.copy()).cuda()
        lst = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(lst, t)
        assert all(torch.equal(lst[0], u) for u in lst), name
        model.close()
    return worst


def _minibatch_checks(ctx, rank, world):
    worst = 0.0
    # ---- mini-batches across ranks: every (shuffled) batch is split over the ranks, each rank uploads its slices
    # batch after batch and steps through its local offsets; held-out rows are sharded the same way
    name = "c3_logreg_lqn_l1"
    A, y, x0 = cases.data(name)
    n = A.shape[0]
    perm = np.random.default_rng(9).permutation(n)
    At, yt = A[:600] * 1.25, y[:600]
    # bs = 100 < m = 128: every step of that run is the GGN underdetermined branch with the batch rows spread over the ranks
    for mname, bs in (("ProxLQNSCORE", 700), ("ProxGGNSCORE", 1100), ("ProxGGNSCORE", 100)):
        loss_o = O.LogisticLoss(1 / n, "consistent")
        loss_g = S.LogisticLoss(1 / n, "consistent")
        so = O.iterate(getattr(O, mname)(), O.Problem(A, y, x0, loss_o, 1e-2, Atest=At, ytest=yt), "l1",
                       O.PHuberSmootherL1L2(1.0), max_epoch=4, alpha=1, batch_size=bs, shuffle_batch=True, perm=perm,
                       local_max_iter=3)
        rows, loc = S.batch_shard(n, world, rank, bs, local_max_iter=3, perm=perm)
        t0, tl = S.shard_rows(600, world, rank)
        for dl in (False, True):
            model = S.Problem(A[rows], y[rows], x0, loss_g, 1e-2, ctx=ctx, Atest=At[t0:t0 + tl], ytest=yt[t0:t0 + tl])
            # local_max_iter also makes iterate! run a single epoch (iterate.jl:58-70); the offsets already hold 3 batches
            sg = S.iterate(getattr(S, mname)(), model, "l1", S.PHuberSmootherL1L2(1.0), max_epoch=4, alpha=1,
                           verbose=0, device_loop=dl, batch_offsets=loc, local_max_iter=3)
            ex = np.linalg.norm(sg.x - so.x) / np.linalg.norm(so.x)
            eo = max(abs(a - b) / abs(b) for a, b in zip(sg.obj, so.obj))
            et = max(abs(a - b) / abs(b) for a, b in zip(sg.fvaltest, so.fvaltest))
            assert sg.epochs == so.epochs and len(sg.obj) == len(so.obj) == len(sg.fvaltest), (mname, dl)
            tol = 1e-9 if bs < A.shape[1] else 1e-10  # the wide branch has a general LU in every step
            assert ex <= tol and eo <= tol and et <= tol, (mname, bs, dl, rank, ex, eo, et)
            worst = max(worst, ex, eo, et)
            model.close()
    return worst


if __name__ == "__main__":
    main()
