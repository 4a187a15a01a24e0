"""GPU parity, BatchKF (batch.go:34-79) and BatchGroundTruth (truth.go), against the CPU oracle, 1e-10."""
import numpy as np
import pytest

import fixtures as fx

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _gpu():
    import gokalman_b200 as gk
    gk.load()
    return gk


@pytest.mark.parametrize("n,m,nf,shared", [(6, 2, 133, False), (4, 2, 7, True), (3, 1, 40, False), (8, 3, 20, False), (7, 1, 9, True)])
def test_batch_solve_matches_oracle(oracle, n, m, nf, shared):
    gk = _gpu()
    rng = np.random.default_rng(31 + n + nf)
    steps = 57
    H = rng.standard_normal((steps, m, n)) if shared else rng.standard_normal((steps, m, n, nf))
    real = rng.standard_normal((steps, m, nf))
    comp = real + 0.1 * rng.standard_normal((steps, m, nf))
    R = np.diag(rng.uniform(0.5, 2.0, m))
    kf = gk.NewBatchKF(steps, gk.NewNoiseless(np.eye(n), R))
    x, P, status = kf.SolveBatch(H, real, comp, n_filters=nf)
    assert np.all(status == 0)
    for f in sorted(set([0, 1, nf // 2, nf - 1])):
        Hf = H if shared else H[:, :, :, f]
        xr, Pr = oracle.batch_solve(R, np.ascontiguousarray(Hf), real[:, :, f], comp[:, :, f])
        assert fx.scaled_err(x[:, f], xr) <= TOL, (f, fx.scaled_err(x[:, f], xr))
        assert fx.scaled_err(P[:, :, f], Pr) <= TOL, (f, fx.scaled_err(P[:, :, f], Pr))


def test_batchkf_go_style_api_and_singular(oracle):
    """NewBatchKF / SetNextMeasurement / Solve one measurement at a time, like batch.go's callers; a rank
    deficient Lambda makes Solve return an error."""
    gk = _gpu()
    rng = np.random.default_rng(5)
    n, m, count = 6, 2, 20
    R = np.diag([1e-2, 1e-3])
    kf = gk.NewBatchKF(count, gk.NewNoiseless(np.eye(n), R))
    Hs, ys, cs = [], [], []
    for k in range(count):
        H, y, c = rng.standard_normal((m, n)), rng.standard_normal(m), rng.standard_normal(m)
        kf.SetNextMeasurement(y, c, np.eye(n), H)
        Hs.append(H); ys.append(y); cs.append(c)
    with pytest.raises(IndexError):
        kf.SetNextMeasurement(ys[0], cs[0], np.eye(n), Hs[0])
    x, P = kf.Solve()
    xr, Pr = oracle.batch_solve(R, np.stack(Hs), np.stack(ys), np.stack(cs))
    assert fx.scaled_err(x, xr) <= TOL and fx.scaled_err(P, Pr) <= TOL
    kf2 = gk.NewBatchKF(2, gk.NewNoiseless(np.eye(n), R))  # 2 measurements x 2 rows < 6 states: singular Lambda
    kf2.SetNextMeasurement(ys[0], cs[0], np.eye(n), np.zeros((m, n)))
    with pytest.raises(gk.GkbError):
        kf2.Solve()


def test_batch_ground_truth_error(oracle):
    gk = _gpu()
    f = fx.jerk3()
    kf, _ = gk.NewVanilla(f["x0"], f["P0"], f["F"], f["G"], f["H"], gk.NewNoiseless(f["Q"], f["R"]))
    est = kf.Update(np.array([0.3]), np.array([0.1]))
    truth = gk.NewBatchGroundTruth([np.array([1.0, 2.0, 3.0])], [np.array([0.25])])
    err = truth.ErrorWithOffset(0, est, np.array([0.5, 0.5, 0.5]))
    assert np.allclose(err.State(), np.asarray(est.State()) + 0.5 - np.array([1.0, 2.0, 3.0]), rtol=0, atol=0)
    assert np.allclose(err.Measurement(), np.asarray(est.Measurement()) - 0.25, rtol=0, atol=0)
    assert np.array_equal(err.Covariance(), est.Covariance())
    zero = truth.Error(-1, est)
    assert not np.any(zero.State()) and not np.any(zero.Measurement())
