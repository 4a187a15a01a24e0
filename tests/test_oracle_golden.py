"""Pins the CPU oracle to the reference's own golden vectors and known-answer tests."""
import numpy as np

import fixtures as fx


def _make_oracle_filters(gko):
    def make(f):
        v = gko.NewVanilla(f["x0"], f["P0"], f["F"], f["G"], f["H2"], f["Q"], f["Ra"])
        i = gko.NewInformation(np.zeros(4), np.zeros((4, 4)), f["F"], f["G"], f["H2"], f["Q"], f["Ra"])
        s = gko.NewSquareRoot(f["x0"], f["P0"], f["F"], f["G"], f["H2"], f["Q"], f["Ra"])
        return [("vanilla", v, v.InitialEstimate()), ("information", i, i.InitialEstimate()),
                ("sqrt", s, s.InitialEstimate())]
    return make


def test_jerkcar_golden_csvs(oracle):
    """examples/jerkcar/{vanilla,information,sqrt}.csv: 2001 rows x 12 columns, printed with %f.
    The restatement must reproduce every cell to print precision (5e-7, plus 1e-9 slack)."""
    out = fx.run_jerkcar(_make_oracle_filters(oracle))
    g = fx.load_jerkcar_golden()
    for name in ("vanilla", "information", "sqrt"):
        assert out[name].shape == g[name].shape == (2001, 12)
        d = np.abs(out[name] - g[name])
        assert d.max() <= 5.0e-7 + 1e-9, (name, d.max())


def test_information_rows_zero_until_observable(oracle):
    """information.csv rows 1..19 are all-zero: the information matrix is singular / cond > 1e16
    until the second position fix (information.go:284-288 returns a zero covariance)."""
    out = fx.run_jerkcar(_make_oracle_filters(oracle), steps=25)
    g = fx.load_jerkcar_golden()
    assert np.all(g["information"][1:20] == 0.0)
    assert np.all(out["information"][1:20] == 0.0)
    assert np.any(out["information"][20] != 0.0)


def test_householder_kat(oracle):
    """helper_test.go:108-117, tolerance 1e-15"""
    A = np.array([[1, -2, -1], [2, -1, 1], [1, 1, 2.0]])
    exp = np.array([[-2.449489742783178, 1.224744871391589, -1.2247448713915892],
                    [0, -2.121320343559643, -2.121320343559643], [0, 0, 0]])
    got = oracle.householder_transf(A, 2, 1)
    assert np.max(np.abs(got - exp)) <= 1e-15


def test_srif_update_kat(oracle):
    """srif_test.go:31-56, tolerance 1e-4"""
    R = np.array([[0.1, 0], [0, 0.1]])
    H = np.array([[1, -2], [2, -1], [1, 1.0]])
    b = np.array([0.2, 0.2])
    y = np.array([-1.1, 1.2, 1.8])
    Rk, bk, ek = oracle.measurement_srif_update(R, H, b, y)
    assert np.max(np.abs(ek - np.array([-0.1319, 0.0871, -0.2810]))) <= 1e-4
    assert np.max(np.abs(bk - np.array([-1.2727, -2.0607]))) <= 1e-4
    assert np.max(np.abs(Rk - np.array([[-2.4515, 1.2237], [0, -2.1243]]))) <= 1e-4


def test_srif_r0_kat(oracle):
    """srif_test.go:15-29: est0.Covariance() == P0 within 1e-12"""
    x0 = np.array([0, 0.35, 0])
    P0 = 10.0 * np.eye(3)
    R = np.diag([(5e-3) ** 2, (5e-6) ** 2])
    kf = oracle.NewSRIF(x0, P0, 2, True, R)
    est0 = kf.InitialEstimate()
    assert np.max(np.abs(est0.Covariance() - P0)) <= 1e-12
    assert np.max(np.abs(est0.State() - x0)) <= 1e-12


def test_philox_kat(oracle):
    """Random123 v1.09 kat_vectors, philox4x32 10 rounds."""
    kats = [
        ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
        ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
        ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
         [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
    ]
    for ctr, key, exp in kats:
        got = oracle.philox4x32_10(ctr, key)
        assert [int(v) for v in got] == exp


def test_icdf_normal_transform(oracle):
    """The engine's Gaussian transform (include/gokalman_b200_icdf.inc): piecewise-quintic inverse normal CDF of
    u = (k + 0.5) 2^-32.  Against scipy's ndtri on random and extreme words (<= 2e-12 absolute), antisymmetric,
    monotone, and standard normal on a uniform sample (moments, Kolmogorov-Smirnov)."""
    from scipy import stats
    from scipy.special import ndtri
    rng = np.random.default_rng(11)
    ks = np.concatenate([rng.integers(0, 2 ** 32, 100000, dtype=np.uint64),
                         np.array([0, 1, 2, 3, 5, 2 ** 31 - 1, 2 ** 31, 2 ** 31 + 1, 2 ** 32 - 2, 2 ** 32 - 1, 2 ** 27, 2 ** 27 - 1],
                                  dtype=np.uint64)])
    z = np.array([oracle.icdf_normal(int(k)) for k in ks])
    ref = ndtri((ks.astype(np.float64) + 0.5) * 2.0 ** -32)
    assert np.max(np.abs(z - ref)) <= 2e-12
    assert abs(oracle.icdf_normal(0) + 6.3379577545) <= 1e-9 and oracle.icdf_normal(2 ** 32 - 1) == -oracle.icdf_normal(0)
    for k in (0, 17, 2 ** 20 + 3, 2 ** 31 - 1):
        assert oracle.icdf_normal(k) == -oracle.icdf_normal(2 ** 32 - 1 - k)
    grid = np.sort(rng.integers(0, 2 ** 32, 5000, dtype=np.uint64))
    zg = np.array([oracle.icdf_normal(int(k)) for k in grid])
    assert np.all(np.diff(zg) >= 0)
    sample = z[:100000]
    assert abs(sample.mean()) <= 0.02 and abs(sample.var() - 1.0) <= 0.02
    assert stats.kstest(sample, "norm").pvalue > 1e-3
    # the committed table is what tools/gen_icdf_table.py produces
    z4 = oracle.philox_normals(0x5EED, 12345, 7, 4)
    words = oracle.philox4x32_10([12345, 0, 7, 0], [0x5EED, 0])
    assert np.array_equal(z4, [oracle.icdf_normal(int(w)) for w in words])
