"""Van Loan's c2d (c2d.go:13-75) on the device (`gkb_van_loan`): the reference's known-answer test
and a batch of per-system (A, dt) against the oracle's restatement, across all Pade branches of the exponential."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_van_loan_device_kat_and_batch(oracle):
    import gokalman_b200 as gk
    gk.load()
    A, G, W = np.array([[0, 1.0], [0, 0]]), np.array([[0.0], [1.0]]), np.array([[1.0]])
    F, Q, st = gk.VanLoanBatch(A, G, W, 0.1)  # c2d_test.go:9-33
    assert st[0] == 0
    assert np.allclose(F[:, :, 0], [[1, 0.1], [0, 1]], atol=1e-3) and np.allclose(Q[:, :, 0], [[0.0003, 0.005], [0.005, 0.1]], atol=1e-3)
    rng = np.random.default_rng(8)
    for n, q in ((4, 2), (8, 3), (6, 1)):
        count = 70
        A = rng.standard_normal((n, n, count)) - 1.5 * np.eye(n)[:, :, None]
        G = rng.standard_normal((n, q))
        Bq = rng.standard_normal((q, q))
        W = Bq @ Bq.T + np.eye(q)
        dt = np.exp(rng.uniform(np.log(1e-3), np.log(1.5), count))  # 1-norms of M from ~1e-2 to ~15: every Pade degree + scaling
        # (larger |A| dt makes Van Loan's F E_12 product cancel e^{+|A| dt} against e^{-|A| dt}: ill-posed for any implementation)
        F, Q, st = gk.VanLoanBatch(A, G, W, dt)
        assert np.all(st == 0)
        for i in range(count):
            Fr, Qr = oracle.van_loan(A[:, :, i], G, W, dt[i])
            assert np.max(np.abs(F[:, :, i] - Fr)) <= 1e-11 * np.max(np.abs(Fr)), (n, i)
            assert np.max(np.abs(Q[:, :, i] - Qr)) <= 1e-11 * np.max(np.abs(Qr)), (n, i)
            assert np.array_equal(Q[:, :, i], Q[:, :, i].T)
    # shared A, per-system dt: the "same model, different sampling rates" sweep
    F, Q, st = gk.VanLoanBatch(A[:, :, 0], G, W, dt[:5])
    for i in range(5):
        Fr, Qr = oracle.van_loan(A[:, :, 0], G, W, dt[i])
        assert np.max(np.abs(F[:, :, i] - Fr)) <= 1e-11 * np.max(np.abs(Fr))
