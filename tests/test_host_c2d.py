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
