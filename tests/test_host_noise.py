"""noise_test.go on the host-side Noise classes (gokalman_b200/api.py): TestImplementsNoise, TestBlankNoise, TestBatchNoise
and the constructor half of TestAWGN (noise_test.go:9-134).  No GPU: AWGN samples come from the device and are covered by
tests/test_gpu_awgn.py."""
import numpy as np
import pytest

import gokalman_b200 as gk


def test_implements_noise():
    """noise_test.go:9-14: every noise type carries the Noise method set (noise.go:13-20)."""
    for cls in (gk.Noiseless, gk.BatchNoise, gk.AWGN, gk.ReplayNoise):
        for method in ("Process", "Measurement", "ProcessMatrix", "MeasurementMatrix", "Reset", "__str__"):
            assert callable(getattr(cls, method)), (cls.__name__, method)


def test_blank_noise():
    """noise_test.go:16-58"""
    with pytest.raises(ValueError):
        gk.NewNoiseless(None, None)  # noise.go:30-32 panics
    nl = gk.NewNoiseless(np.zeros((2, 2)), np.zeros((3, 3)))
    assert "Noiseless" in str(nl)
    nl.Reset()
    assert nl.Process(1).shape == (2,) and nl.Measurement(1).shape == (3,)
    assert not nl.Process(1).any() and not nl.Measurement(1).any()
    Q, R = nl.ProcessMatrix(), nl.MeasurementMatrix()
    assert Q.shape == (2, 2) and not Q.any()
    assert R.shape == (3, 3) and not R.any()


def test_batch_noise():
    """noise_test.go:60-111: stored vectors by step, Q = R = zero matrices of the vectors' sizes, panic past the end."""
    process = [np.array([i + 1.0, i + 2.0, i + 3.0]) for i in range(4)]
    measurements = [np.array([2.0 * i + 1.0, 2.0 * i + 2.0]) for i in range(4)]
    batch = gk.BatchNoise(process, measurements)
    batch.Reset()
    assert str(batch) == "BatchNoise"
    for k in range(4):
        assert np.array_equal(batch.Process(k), process[k])
        assert np.array_equal(batch.Process(k), process[k])  # indexed by k: the same vector on a second call (noise.go:73-86)
        assert np.array_equal(batch.Measurement(k), measurements[k])
    Q, R = batch.ProcessMatrix(), batch.MeasurementMatrix()
    assert Q.shape == (3, 3) and not Q.any()
    assert R.shape == (2, 2) and not R.any()
    with pytest.raises(IndexError):
        batch.Process(4)
    with pytest.raises(IndexError):
        batch.Measurement(4)


def test_awgn_constructor():
    """noise_test.go:113-134: a covariance that is not positive definite panics; the matrices are carried unchanged."""
    bad_q = np.array([[1.0, 1.0], [1.0, 1.0]])
    bad_r = np.array([[2.0, 3, 1], [3, 4, 6], [1, 6, 7]])
    with pytest.raises(ValueError):
        gk.NewAWGN(bad_q, bad_r)
    with pytest.raises(ValueError):
        gk.NewAWGN(np.eye(2), bad_r)
    Q, R = np.eye(2), np.array([[20.0, 0.05], [0.05, 20.0]])
    n = gk.NewAWGN(Q, R, seed=7)
    assert np.array_equal(n.ProcessMatrix(), Q) and np.array_equal(n.MeasurementMatrix(), R)
    assert "AWGN" in str(n)
    seed = n.seed
    n.Reset()
    assert n.seed == seed  # an explicit seed is kept across Reset(); seed=None re-seeds like the reference (noise.go:145-159)
    a, b = gk.NewAWGN(Q, R), gk.NewAWGN(Q, R)
    assert a.seed != b.seed
