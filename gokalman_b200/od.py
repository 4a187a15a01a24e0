"""Host side of the device OD-input synthesis (`gkb_od_synthesize`, `gkb_od_run`): the statOD scenario tables.

The reference's orbit-determination tests (hybrid_test.go:71-82,159-294; srif_test.go:70-80) build, with the external
`smd` propagator, a LEO truth orbit, three DSN-like ground stations and range / range-rate measurements, and feed
`Prepare(Phi, Htilde)` / `Update(real, computed)` one epoch at a time.  Here the per-FILTER work (reference orbit,
STM, partials, computed observations) runs on the GPU; what remains on the host is small and shared by the whole
batch: the truth orbit, which station tracks at each epoch, where that station is (ECI) and what it measures on the
truth.  numpy only -- nothing here touches the CPU oracle.
"""
import ctypes as C

import numpy as np

from . import _lib

MU_EARTH = 398600.4415          # km^3/s^2
J2_EARTH = 1.0826269e-3
R_EARTH = 6378.1363             # km
OMEGA_EARTH = 7.292115900231276e-5  # rad/s
# (latitude, longitude) in degrees: hybrid_test.go:71-75 / SURVEY App. D
STATIONS_DEG = ((-35.398333, 148.981944), (40.427222, 355.749444), (35.247164, 243.205))


def coe_to_rv(a, e, inc, raan, argp, nu, mu=MU_EARTH):
    """Classical orbital elements (km, degrees) -> ECI position / velocity."""
    inc, raan, argp, nu = np.radians([inc, raan, argp, nu])
    p = a * (1.0 - e * e)
    r = p / (1.0 + e * np.cos(nu))
    rp = np.array([r * np.cos(nu), r * np.sin(nu), 0.0])
    vp = np.sqrt(mu / p) * np.array([-np.sin(nu), e + np.cos(nu), 0.0])

    def rot3(t):
        return np.array([[np.cos(t), -np.sin(t), 0], [np.sin(t), np.cos(t), 0], [0, 0, 1.0]])

    def rot1(t):
        return np.array([[1.0, 0, 0], [0, np.cos(t), -np.sin(t)], [0, np.sin(t), np.cos(t)]])
    Q = rot3(raan) @ rot1(inc) @ rot3(argp)
    return np.concatenate([Q @ rp, Q @ vp])


def accel(r, mu=MU_EARTH, j2=J2_EARTH, re=R_EARTH):
    x, y, z = r
    rn = np.sqrt(x * x + y * y + z * z)
    k = 1.5 * j2 * mu * re * re
    f = 1.0 / rn ** 5 - 5.0 * z * z / rn ** 7
    g = 3.0 / rn ** 5 - 5.0 * z * z / rn ** 7
    return -mu * r / rn ** 3 - k * np.array([x * f, y * f, z * g])


def rk4_orbit(X, h, mu=MU_EARTH, j2=J2_EARTH, re=R_EARTH):
    """One classical RK4 step of the two-body + J2 dynamics (the device kernel's propagation)."""
    def f(X):
        return np.concatenate([X[3:], accel(X[:3], mu, j2, re)])
    k1 = f(X)
    k2 = f(X + 0.5 * h * k1)
    k3 = f(X + 0.5 * h * k2)
    k4 = f(X + h * k3)
    return X + h / 6.0 * (k1 + 2 * k2 + 2 * k3 + k4)


def station_eci(lat_deg, lon_deg, t, theta0=0.0, re=R_EARTH, omega=OMEGA_EARTH):
    """ECI position / velocity of a ground station on the spherical Earth at time t."""
    lat, lon = np.radians(lat_deg), np.radians(lon_deg)
    th = theta0 + omega * t + lon
    rs = re * np.array([np.cos(lat) * np.cos(th), np.cos(lat) * np.sin(th), np.sin(lat)])
    vs = omega * np.array([-rs[1], rs[0], 0.0])
    return rs, vs


def range_rate(X, rs, vs):
    d, dv = X[:3] - rs, X[3:] - vs
    rho = np.sqrt(d @ d)
    return rho, (d @ dv) / rho


class Scenario:
    """Per-epoch tables of an OD run, shared by every filter of the batch."""

    def __init__(self, steps, dt, truth0, mask_deg=10.0, always_track=False, ekf_after=15, stations=STATIONS_DEG,
                 mu=MU_EARTH, j2=J2_EARTH, re=R_EARTH, theta0=0.0):
        """truth0: the truth's ECI state at t = 0.  Epoch k ends at t = (k + 1) dt, where the measurement is taken.
        always_track=False: an epoch has a measurement only when a station sees the truth above `mask_deg` of
        elevation (hybrid_test.go:75: 10 degrees), otherwise it is a Predict() epoch; True: the station with the
        highest elevation tracks at every epoch (the throughput configuration: every epoch is an Update).
        ekf_after: CKF for that many measurement epochs, EKF afterwards (hybrid_test.go:65,270-273).
        theta0: Greenwich sidereal angle at t = 0 [rad] (moves the passes within the run)."""
        self.steps, self.dt, self.mu, self.j2, self.re = int(steps), float(dt), mu, j2, re
        self.station = np.zeros((steps, 6))
        self.truth_obs = np.zeros((steps, 2))
        self.truth = np.zeros((steps, 6))
        self.flags = np.zeros(steps, dtype=np.uint8)
        self.station_index = np.full(steps, -1, dtype=np.int32)
        X = np.asarray(truth0, dtype=np.float64).copy()
        seen = 0
        for k in range(steps):
            X = rk4_orbit(X, dt, mu, j2, re)
            self.truth[k] = X
            t = (k + 1) * dt
            best, best_el = -1, -np.inf
            for s, (lat, lon) in enumerate(stations):
                rs, vs = station_eci(lat, lon, t, theta0=theta0, re=re)
                d = X[:3] - rs
                el = np.degrees(np.arcsin((d @ rs) / (np.linalg.norm(d) * np.linalg.norm(rs))))
                if el > best_el:
                    best, best_el = s, el
            rs, vs = station_eci(*stations[best], t, theta0=theta0, re=re)
            self.station[k, :3], self.station[k, 3:] = rs, vs
            self.truth_obs[k] = range_rate(X, rs, vs)
            if always_track or best_el >= mask_deg:
                self.station_index[k] = best
                self.flags[k] = _lib.F_MEAS | (_lib.F_EKF if seen >= ekf_after else 0)
                seen += 1
            else:
                self.flags[k] = _lib.F_EKF if seen >= ekf_after else 0

    def config(self, orbit0, sigma_range, sigma_rate, seed, filter_offset=0, orbit_mem=_lib.HOST):
        """gkb_od_config over this scenario's tables; orbit0: [6, n_filters] array (or a device pointer int)."""
        cfg = _lib.OdConfig()
        cfg.mu, cfg.j2, cfg.re, cfg.dt = self.mu, self.j2, self.re, self.dt
        if orbit0 is None:
            cfg.orbit0 = None
        elif isinstance(orbit0, int):
            cfg.orbit0 = orbit0
        else:
            self._orbit0 = np.ascontiguousarray(np.asarray(orbit0, dtype=np.float64))
            cfg.orbit0 = self._orbit0.ctypes.data
        cfg.orbit_mem = orbit_mem
        cfg.station, cfg.truth_obs = self.station.ctypes.data, self.truth_obs.ctypes.data
        cfg.sigma_range, cfg.sigma_rate = float(sigma_range), float(sigma_rate)
        cfg.seed, cfg.filter_offset = int(seed), int(filter_offset)
        return cfg


def leo_truth0():
    """SURVEY App. D / hybrid_test.go:159-160: a = 7000 km, e = 0.001, i = 30, RAAN = 80, argp = 40, nu = 0 (degrees)."""
    return coe_to_rv(7000.0, 0.001, 30.0, 80.0, 40.0, 0.0)


def perturbed_orbits(truth0, n_filters, sigma_r=1.0, sigma_v=1e-3, seed=0):
    """[6, n_filters] initial reference orbits: the truth plus N(0, sigma_r) km / N(0, sigma_v) km/s per component
    (1 km, 1 m/s), seeded."""
    rng = np.random.default_rng(seed)
    X = np.repeat(np.asarray(truth0, dtype=np.float64)[:, None], n_filters, axis=1)
    X[:3] += sigma_r * rng.standard_normal((3, n_filters))
    X[3:] += sigma_v * rng.standard_normal((3, n_filters))
    return np.ascontiguousarray(X)


def synthesize(scn, orbit0, sigma_range, sigma_rate, seed, device=0, filter_offset=0):
    """Host-buffer call of gkb_od_synthesize: returns Phi [steps, 36, N], Htilde [steps, 12, N], real / computed
    [steps, 2, N] and the final reference orbits [6, N]."""
    orbit0 = np.ascontiguousarray(np.asarray(orbit0, dtype=np.float64))
    nf, steps = orbit0.shape[1], scn.steps
    Phi, Ht = np.zeros((steps, 36, nf)), np.zeros((steps, 12, nf))
    real, comp, orb = np.zeros((steps, 2, nf)), np.zeros((steps, 2, nf)), np.zeros((6, nf))
    cfg = scn.config(orbit0, sigma_range, sigma_rate, seed, filter_offset)
    _lib.check(_lib.load().gkb_od_synthesize(C.byref(cfg), steps, nf, device, Phi.ctypes.data, Ht.ctypes.data,
                                             real.ctypes.data, comp.ctypes.data, _lib.HOST, orb.ctypes.data))
    return Phi, Ht, real, comp, orb
