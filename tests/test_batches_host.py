"""Host-side mini-batch logic (no GPU): the batch tables of the host mirror agree with the oracle's restatement of
optim_loop!'s data loader (iterate.jl:122-145, utils.jl:14-25), and the multi-rank layout covers every row once."""
import numpy as np
import pytest

from oracle import scs_oracle as O


def _api():
    import scs_b200.api as api  # importing the mirror does not load the CUDA library
    return api


@pytest.mark.parametrize("n,bs,lmi,ss", [(100, 10, None, False), (103, 10, None, False), (103, 10, 4, False),
                                          (7, 100, None, False), (9, None, None, True), (50, 7, 2.9, False),
                                          (50, 7, 0, False), (50, 7, 99, False), (20, 5, None, True)])
@pytest.mark.parametrize("shuffle", [False, True])
def test_batch_plan_matches_oracle(n, bs, lmi, ss, shuffle):
    api = _api()
    perm = np.random.default_rng(n).permutation(n) if shuffle else None
    order, off = api.batch_plan(n, bs, ss, shuffle, lmi, perm)
    ob = O.make_batches(n, bs, ss, shuffle, lmi, perm)
    assert len(ob) == len(off) - 1
    rows = np.arange(n) if order is None else order
    for i, idx in enumerate(ob):
        assert np.array_equal(rows[off[i]:off[i + 1]], idx)
    if bs is not None and not ss:  # max_iter = ceil(m / batch_size) (:126), capped by local_max_iter (:127)
        want = -(-n // bs)
        if lmi is not None and int(np.floor(lmi)) > 0:
            want = min(want, int(np.floor(lmi)))
        assert len(ob) == want


@pytest.mark.parametrize("n,bs,world,lmi", [(1000, 64, 2, None), (1003, 100, 4, None), (500, 37, 3, 5), (64, 64, 8, None)])
def test_batch_shard_partitions_rows(n, bs, world, lmi):
    api = _api()
    perm = np.random.default_rng(1).permutation(n)
    seen = []
    per_rank = [api.batch_shard(n, world, r, bs, lmi, perm) for r in range(world)]
    _, goff = api.batch_plan(n, bs, False, True, lmi, perm)
    for rows, loc in per_rank:
        assert loc[0] == 0 and np.all(np.diff(loc) >= 0) and loc[-1] <= len(rows)
        assert len(loc) == len(goff)
        seen.append(rows)
    allrows = np.concatenate(seen)
    assert np.array_equal(np.sort(allrows), np.arange(n))  # every row on exactly one rank
    for i in range(len(goff) - 1):  # batch i, gathered over the ranks, is the global batch i
        got = np.concatenate([rows[loc[i]:loc[i + 1]] for rows, loc in per_rank])
        assert np.array_equal(np.sort(got), np.sort(perm[goff[i]:goff[i + 1]]))


def test_oracle_minibatch_full_batch_equivalence():
    """batch_size = n (one batch, no shuffle) must reproduce the full-batch loop exactly."""
    import cases
    for name in ("c3_logreg_lqn_l1", "c5_ls_n_indbox"):
        m1, mod1, reg, h1, kw = cases.build(name, O)
        s1 = O.iterate(m1, mod1, reg, h1, **kw)
        m2, mod2, reg, h2, kw = cases.build(name, O)
        s2 = O.iterate(m2, mod2, reg, h2, batch_size=mod2.A.shape[0], **kw)
        assert s1.epochs == s2.epochs and s1.obj == s2.obj and np.array_equal(s1.x, s2.x)
