"""c2d.go VanLoan (host-side by BASELINE's own scope statement): the reference's known-answer test."""
import numpy as np


def test_van_loan_kat():
    """c2d_test.go:9-33"""
    import gokalman_b200 as gk
    A = np.array([[0, 1.0], [0, 0]])
    G = np.array([[0.0], [1.0]])
    W = np.array([[1.0]])
    F, Q, err = gk.VanLoan(A, G, W, 0.1)
    assert err is None
    assert np.allclose(F, [[1, 0.1], [0, 1]], atol=1e-3)
    assert np.allclose(Q, [[0.0003, 0.005], [0.005, 0.1]], atol=1e-3)
    assert np.array_equal(Q, Q.T)
    _, _, err = gk.VanLoan(np.array([[1, 1.0], [0, 1]]), G, W, 10)
    assert err is not None and "Nyquist" in err


def test_oracle_van_loan_against_scipy_and_kat(oracle):
    """The oracle's restatement of c2d.go (Higham-2005 expm) against scipy.linalg.expm over all five Pade branches
    (norms from 1e-3 to 50) and the reference's known-answer test."""
    import gokalman_b200 as gk
    from scipy.linalg import expm
    rng = np.random.default_rng(4)
    for d, scale in ((4, 1e-3), (6, 0.1), (8, 0.5), (8, 1.5), (12, 4.0), (16, 40.0)):
        M = scale * rng.standard_normal((d, d)) / np.sqrt(d)
        E, Es = oracle.expm(M), expm(M)
        assert np.max(np.abs(E - Es)) <= 1e-12 * max(1.0, np.max(np.abs(Es))), (d, scale)
    A, G, W = np.array([[0, 1.0], [0, 0]]), np.array([[0.0], [1.0]]), np.array([[1.0]])
    F, Q = oracle.van_loan(A, G, W, 0.1)
    assert np.allclose(F, [[1, 0.1], [0, 1]], atol=1e-3) and np.allclose(Q, [[0.0003, 0.005], [0.005, 0.1]], atol=1e-3)
    for n, q, dt in ((4, 2, 0.1), (8, 3, 2.0)):
        A = rng.standard_normal((n, n)) - 1.5 * np.eye(n)
        G = rng.standard_normal((n, q))
        Bq = rng.standard_normal((q, q))
        W = Bq @ Bq.T + np.eye(q)
        F, Q = oracle.van_loan(A, G, W, dt)
        Fh, Qh, _ = gk.VanLoan(A, G, W, dt)  # the host-side scipy implementation
        assert np.max(np.abs(F - Fh)) <= 1e-12 * np.max(np.abs(Fh)) and np.max(np.abs(Q - Qh)) <= 1e-11 * np.max(np.abs(Qh))
