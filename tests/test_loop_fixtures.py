"""Loop-shape fixtures derived BY HAND from src/algorithms/iterate.jl — independent of oracle/scs_oracle.py, so a
misreading shared by the oracle and the host mirror (both written from the same reading of the reference) shows up.

With x_tol = f_tol = 0 nothing terminates early (`pri_res_norm < 0` and `‖Δx‖ < 0` are never true, `f_rel_error <= 0`
needs obj == obj_star exactly), so the counts below follow from the control flow alone:

  * iterate! builds Options(max_epoch = local_max_iter !== nothing ? 1 : max_epoch)            iterate.jl:58-70
  * max_iter = batch_size !== nothing ? ceil(m / batch_size) : 1                               :122,:126
  * iend = (local_max_iter !== nothing && floor(local_max_iter) > 0) ? min(floor(lmi), max_iter) : max_iter   :127
  * batch_size && slice_samples -> slice_samples = false                                       :128-131
  * slice_samples: data = zip(rows, targets); opt.batch_size = 1 — AFTER iend was fixed        :136-138
  * data = collect(loader)[1:iend]                                                             :145, utils.jl:21-25
  * every epoch pushes one history entry (:202); the last batch of the last epoch pushes one more (:219-231);
    epochs += 1 at the end of every epoch that does not break (:261)

=> steps = max_epoch' * iend, history = max_epoch' + 1, epochs = max_epoch', and step i of an epoch sees the rows of
batch i (slice_samples: the first row, every time)."""
import numpy as np
import pytest

from oracle import scs_oracle as O

N, M = 23, 4
# (options) -> (steps, history entries, epochs, rows seen by the steps of one epoch)
FIXTURES = [
    (dict(max_epoch=4), 4, 5, 4, [23]),
    (dict(max_epoch=1), 1, 2, 1, [23]),
    (dict(max_epoch=4, batch_size=10), 12, 5, 4, [10, 10, 3]),
    (dict(max_epoch=4, batch_size=10, local_max_iter=2), 2, 2, 1, [10, 10]),
    (dict(max_epoch=4, batch_size=10, local_max_iter=2.9), 2, 2, 1, [10, 10]),
    (dict(max_epoch=4, batch_size=10, local_max_iter=99), 3, 2, 1, [10, 10, 3]),
    (dict(max_epoch=4, batch_size=10, local_max_iter=0), 3, 2, 1, [10, 10, 3]),  # floor(0) > 0 is false: iend = max_iter
    (dict(max_epoch=4, local_max_iter=5), 1, 2, 1, [23]),  # no batch_size: max_iter = 1
    (dict(max_epoch=4, slice_samples=True), 4, 5, 4, [1]),  # iend = 1: only the first row ever steps
    (dict(max_epoch=4, slice_samples=True, local_max_iter=7), 1, 2, 1, [1]),
    (dict(max_epoch=4, slice_samples=True, batch_size=10), 12, 5, 4, [10, 10, 3]),  # batch_size wins
    (dict(max_epoch=3, batch_size=23), 3, 4, 3, [23]),
    (dict(max_epoch=3, batch_size=100), 3, 4, 3, [23]),
]


def _data():
    rng = np.random.default_rng(7)
    A = rng.standard_normal((N, M))
    y = np.where(rng.random(N) < 0.5, -1.0, 1.0)
    x0 = 0.3 * rng.standard_normal(M)
    return A, y, x0


@pytest.mark.parametrize("opts,steps,hist,epochs,rows", FIXTURES)
def test_oracle_loop_shape(opts, steps, hist, epochs, rows):
    A, y, x0 = _data()
    model = O.Problem(A, y, x0, O.LogisticLoss(1 / N), 1e-2)
    method = O.ProxLQNSCORE()
    seen = []
    real_step = method.step

    def spy(bm, *a, **k):
        seen.append((bm.A.shape[0], bm.A[0].copy()))
        return real_step(bm, *a, **k)

    method.step = spy
    sol = O.iterate(method, model, "l1", O.PHuberSmootherL1L2(1.0), alpha=0.5, x_tol=0.0, f_tol=0.0, **opts)
    assert len(seen) == steps == len(sol.iterates)
    assert len(sol.obj) == len(sol.fval) == len(sol.pri_res_norm) == len(sol.rel) == len(sol.objrel) == hist
    assert sol.epochs == epochs
    assert [s[0] for s in seen] == rows * (steps // len(rows))
    assert sol.pri_res_norm[0] is None  # :178,:202
    if rows == [1]:  # slice_samples: always the FIRST row of A (no shuffle)
        assert all(np.array_equal(s[1], A[0]) for s in seen)
    if not opts.get("shuffle_batch"):
        assert np.array_equal(seen[0][1], A[0])


@pytest.mark.parametrize("opts,steps,hist,epochs,rows", FIXTURES)
def test_host_mirror_batch_plan_shape(opts, steps, hist, epochs, rows):
    """The batch table of the host mirror (no GPU needed): number and size of the batches of one epoch."""
    import scs_b200.api as api
    order, off = api.batch_plan(N, opts.get("batch_size"), opts.get("slice_samples", False), False,
                                opts.get("local_max_iter"))
    assert order is None
    assert list(np.diff(off)) == rows and off[0] == 0


@pytest.mark.gpu
@pytest.mark.parametrize("device_loop", [False, True])
@pytest.mark.parametrize("opts,steps,hist,epochs,rows", FIXTURES)
def test_gpu_loop_shape(scs, opts, steps, hist, epochs, rows, device_loop):
    A, y, x0 = _data()
    model = scs.Problem(A, y, x0, scs.LogisticLoss(1 / N), 1e-2)
    sol = scs.iterate(scs.ProxLQNSCORE(), model, "l1", scs.PHuberSmootherL1L2(1.0), alpha=0.5, x_tol=0.0, f_tol=0.0,
                      verbose=0, shuffle_batch=False, device_loop=device_loop, **opts)
    assert len(sol.obj) == len(sol.fval) == len(sol.pri_res_norm) == len(sol.rel) == len(sol.objrel) == hist
    assert sol.epochs == epochs and sol.pri_res_norm[0] is None
    # the iterate itself against the (now fixture-pinned) oracle
    so = O.iterate(O.ProxLQNSCORE(), O.Problem(A, y, x0, O.LogisticLoss(1 / N), 1e-2), "l1", O.PHuberSmootherL1L2(1.0),
                   alpha=0.5, x_tol=0.0, f_tol=0.0, **opts)
    assert np.linalg.norm(sol.x - so.x) <= 1e-10 * max(np.linalg.norm(so.x), 1e-300)
    model.close()


@pytest.mark.gpu
def test_shuffle_twice_permutes_the_original_rows(scs):
    """Two iterate!(…, shuffle_batch=true) calls on the same Problem both shuffle model.A as given (utils.jl:18-25):
    the second call must not compose its permutation with the first one's."""
    A, y, x0 = _data()
    p1, p2 = np.random.default_rng(1).permutation(N), np.random.default_rng(2).permutation(N)
    kw = dict(alpha=0.5, x_tol=0.0, f_tol=0.0, max_epoch=3, batch_size=5)
    model = scs.Problem(A, y, x0, scs.LogisticLoss(1 / N), 1e-2)
    scs.iterate(scs.ProxLQNSCORE(), model, "l1", scs.PHuberSmootherL1L2(1.0), verbose=0, shuffle_batch=True, perm=p1, **kw)
    assert np.array_equal(model.row_order, p1)
    s2 = scs.iterate(scs.ProxLQNSCORE(), model, "l1", scs.PHuberSmootherL1L2(1.0), verbose=0, shuffle_batch=True, perm=p2, **kw)
    assert np.array_equal(model.row_order, p2)
    so = O.iterate(O.ProxLQNSCORE(), O.Problem(A, y, x0, O.LogisticLoss(1 / N), 1e-2), "l1", O.PHuberSmootherL1L2(1.0),
                   shuffle_batch=True, perm=p2, **kw)
    assert np.linalg.norm(s2.x - so.x) <= 1e-10 * np.linalg.norm(so.x)
    s3 = scs.iterate(scs.ProxLQNSCORE(), model, "l1", scs.PHuberSmootherL1L2(1.0), verbose=0, shuffle_batch=False, **kw)
    assert model.row_order is None  # unshuffled batches see the rows as given again
    so3 = O.iterate(O.ProxLQNSCORE(), O.Problem(A, y, x0, O.LogisticLoss(1 / N), 1e-2), "l1", O.PHuberSmootherL1L2(1.0), **kw)
    assert np.linalg.norm(s3.x - so3.x) <= 1e-10 * np.linalg.norm(so3.x)
    model.close()


@pytest.mark.gpu
def test_metrics_callbacks(scs):
    """metrics::Dict{name => (model, x) -> value}: one value per recorded state (utils.jl:80-83)."""
    A, y, x0 = _data()
    model = scs.Problem(A, y, x0, scs.LogisticLoss(1 / N), 1e-2)
    sol = scs.iterate(scs.ProxLQNSCORE(), model, "l1", scs.PHuberSmootherL1L2(1.0), alpha=0.5, x_tol=0.0, f_tol=0.0,
                      max_epoch=3, verbose=0, metrics={"nnz": lambda mdl, x: int(np.count_nonzero(x)),
                                                       "norm": lambda mdl, x: float(np.linalg.norm(x))})
    assert len(sol.metricvals["nnz"]) == len(sol.metricvals["norm"]) == len(sol.obj) == 4
    assert sol.metricvals["norm"][0] == float(np.linalg.norm(x0))
    with pytest.raises(scs.UnsupportedError):
        scs.iterate(scs.ProxLQNSCORE(), model, "l1", scs.PHuberSmootherL1L2(1.0), device_loop=True, verbose=0,
                    metrics={"nnz": lambda mdl, x: 0})
    model.close()
