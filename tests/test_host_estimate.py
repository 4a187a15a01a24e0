"""Host-side pieces of the Go API mirror that need no GPU: VanillaEstimate.IsWithinNσ (vanilla.go:231-246), the
exported helpers of helper.go and BatchGroundTruth (truth.go:10-70)."""
import numpy as np
import pytest


def _est(gk, x, P):
    n = len(x)
    f = {"state": np.asarray(x, dtype=np.float64).reshape(1, n, 1), "covar": np.asarray(P, dtype=np.float64).reshape(1, n * n, 1),
         "meas": np.zeros((1, 1, 1))}
    return gk.Estimate(n, 1, f)


def test_is_within_n_sigma():
    """vanilla.go:231-239: every component inside +- N sqrt(P_ii); IsWithin2σ is N = 2 (242-246)."""
    import gokalman_b200 as gk
    P = np.diag([4.0, 0.25])
    assert _est(gk, [3.9, 0.9], P).IsWithinNσ(2)          # |3.9| <= 2*2, |0.9| <= 2*0.5
    assert not _est(gk, [4.1, 0.0], P).IsWithinNσ(2)      # first component outside
    assert not _est(gk, [0.0, -1.01], P).IsWithinNσ(2)    # negative side
    assert _est(gk, [0.0, -1.0], P).IsWithin2σ()          # the bound itself is inside (the reference tests > and <)
    assert _est(gk, [5.9, 1.4], P).IsWithinNσ(3) and not _est(gk, [6.1, 0.0], P).IsWithinNσ(3)
    off = np.array([[4.0, 1.9], [1.9, 0.25]])             # only the diagonal matters
    assert _est(gk, [3.9, 0.4], off).IsWithinNσ(2)


def test_helpers_mirror_helper_go():
    import gokalman_b200 as gk
    assert np.array_equal(gk.Identity(3), np.eye(3)) and np.array_equal(gk.ScaledIdentity(2, 4.0), 4.0 * np.eye(2))
    assert gk.IsNil(None) and gk.IsNil(np.zeros((2, 2))) and not gk.IsNil(np.array([[0, 1e-300]]))
    assert gk.Sign(1e-13) == 1.0 and gk.Sign(-1e-13) == 1.0 and gk.Sign(-1e-11) == -1.0 and gk.Sign(2.0) == 1.0  # helper.go:133-138
    m = np.array([[1.0, 2.0], [2.0 + 5e-7, 3.0]])          # within 1e-6 absolute: accepted, upper triangle kept
    assert np.array_equal(gk.AsSymDense(m), [[1.0, 2.0], [2.0, 3.0]])
    m2 = np.array([[1.0, 2.0], [2.015, 3.0]])              # 1.5e-2 absolute but 7e-3 relative: accepted (abs OR rel)
    assert np.array_equal(gk.AsSymDense(m2), [[1.0, 2.0], [2.0, 3.0]])
    with pytest.raises(gk.GkbError):                        # beyond both tolerances: helper.go:75 returns an error
        gk.AsSymDense(np.array([[1.0, 2.0], [2.1, 3.0]]))
    with pytest.raises(gk.GkbError):
        gk.AsSymDense(np.ones((2, 3)))


def test_batch_ground_truth_error_offsets():
    """truth.go:24-66: Error(k, est) = est - truth (state and measurement), with an optional offset on the state."""
    import gokalman_b200 as gk
    est = _est(gk, [1.0, 2.0], np.eye(2))
    gt = gk.NewBatchGroundTruth([np.array([0.5, 0.5])], [np.array([0.25])])
    e = gt.Error(0, est)
    assert np.allclose(e.State(), [0.5, 1.5]) and np.allclose(e.Measurement(), [-0.25])
    e2 = gt.ErrorWithOffset(0, est, np.array([1.0, -1.0]))
    assert np.allclose(e2.State(), [1.5, 0.5])
    e3 = gt.Error(-1, est)  # k < 0: zeros (truth.go:29)
    assert not np.any(e3.State()) and not np.any(e3.Measurement())
    with pytest.raises(ValueError):
        gk.NewBatchGroundTruth([np.zeros(3)], [np.zeros(1)]).Error(0, est)


def test_implements_ldkf_nldkf_estimate():
    """kalman_test.go:9-33 (TestImplementsLDKF / TestImplementsNLDKF / TestImplementsEst): the host mirror carries the method
    sets of kalman.go:35-72 under the reference's names."""
    import gokalman_b200 as gk
    ldkf = ("Update", "GetNoise", "GetStateTransition", "GetInputControl", "GetMeasurementMatrix", "SetStateTransition",
            "SetInputControl", "SetMeasurementMatrix", "SetNoise", "Reset", "__str__")
    nldkf = ("Prepare", "Predict", "Update", "EKFEnabled", "EnableEKF", "DisableEKF", "PreparePNT", "SetNoise")
    est = ("IsWithinNσ", "State", "Measurement", "Innovation", "Covariance", "PredCovariance", "__str__")
    for cls in (gk.Vanilla, gk.Information, gk.SquareRoot):
        for name in ldkf:
            assert callable(getattr(cls, name, None)), (cls.__name__, name)
    for cls in (gk.HybridKF, gk.SRIF):
        for name in nldkf:
            assert callable(getattr(cls, name, None)), (cls.__name__, name)
    for name in est:
        assert callable(getattr(gk.Estimate, name, None)), name
