"""World-size-2 gloo test of the host-side sharding logic (runs on CPU): contiguous row blocks, the unique-id
broadcast, and the fact that per-shard partial sums of loss / gradient / Gram add up to the full-batch
quantities the oracle computes (the only data-path exchange is an fp64 sum all-reduce)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "selfconcordantsmoothoptimization.jl_b200"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import scs_oracle as O
    from oracle import synth
    from scs_b200.dist import broadcast_unique_id, shard_rows
    uid = broadcast_unique_id(lambda: bytes(range(128)), rank, world)
    n, m = 1001, 17
    r0, nl = shard_rows(n, world, rank)
    A = synth.make_A(nl, m, row0=r0, n_total=n)
    xt = synth.make_x_true(m)
    y = synth.make_labels_logistic(A @ xt, row0=r0)
    x = synth.make_x0(m)
    L = O.LogisticLoss(1 / n)
    z = A @ x
    S = np.sum(np.log(1.0 + np.exp(-y * z)))
    g = A.T @ L.grad_weights(z, y)
    G = A.T @ (L.hess_weights(z, y)[:, None] * A)
    buf = torch.from_numpy(np.concatenate([g, [S], G.ravel()]))
    dist.all_reduce(buf)  # fp64 sum: the exchange step of SURVEY.md §8(e)
    q.put((rank, uid, r0, nl, buf.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_world2_shards_sum_to_full_batch():
    sys.path.insert(0, ROOT)
    from oracle import scs_oracle as O
    from oracle import synth
    world, port = 2, 29533
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in ps]
    res = sorted([q.get(timeout=180) for _ in range(world)], key=lambda t: t[0])
    [p.join(60) for p in ps]
    assert all(p.exitcode == 0 for p in ps)
    assert res[0][1] == res[1][1] == bytes(range(128))
    assert res[0][2] == 0 and res[0][3] + res[1][3] == 1001 and res[1][2] == res[0][3]
    n, m = 1001, 17
    A = synth.make_A(n, m)
    y = synth.make_labels_logistic(A @ synth.make_x_true(m))
    x = synth.make_x0(m)
    L = O.LogisticLoss(1 / n)
    z = A @ x
    full = np.concatenate([A.T @ L.grad_weights(z, y), [np.sum(np.log(1.0 + np.exp(-y * z)))],
                           (A.T @ (L.hess_weights(z, y)[:, None] * A)).ravel()])
    for r in res:
        np.testing.assert_allclose(r[4], full, rtol=1e-12, atol=1e-15)


def test_shard_rows_partition():
    sys.path.insert(0, os.path.join(ROOT, "selfconcordantsmoothoptimization.jl_b200"))
    from scs_b200.dist import shard_rows
    for n in (1, 7, 1000, 1_000_000):
        for w in (1, 2, 4, 8):
            blocks = [shard_rows(n, w, r) for r in range(w)]
            assert blocks[0][0] == 0 and sum(b[1] for b in blocks) == n
            for a, b in zip(blocks, blocks[1:]):
                assert a[0] + a[1] == b[0]
