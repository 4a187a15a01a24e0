"""CPU-side checks of bench.py's bookkeeping (no GPU, no engine call): both arms print the same `config` object for the
hybrid workloads, the workload tables agree, and the reference-formulas FMA-spread probe of the cpu_baseline leg runs."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_both_arms_share_the_config_object():
    import bench
    import bench_hybrid
    for wl in ("hybrid6", "hybrid6_strict", "hybrid6_fma", "srif6"):
        assert wl in bench_hybrid.WORKLOADS and wl in bench.FILTER_WORKLOADS
        a = bench_hybrid.nl_config(wl, 100000, 1000)
        b = bench_hybrid.nl_config(wl, 100000, 1000)
        assert a == b and a["filters_per_gpu"] == 100000 and a["epochs"] == 1000 and a["n"] == 6 and a["m"] == 2
    # the headline names the arithmetic it runs in; the fast mode says what it is
    assert "reference-order" in bench_hybrid.WORKLOADS["hybrid6"]
    assert bench_hybrid.WORKLOADS["hybrid6"] == bench_hybrid.WORKLOADS["hybrid6_strict"]
    assert "production" in bench_hybrid.WORKLOADS["hybrid6_fma"]


def test_bench_cli_parses_every_workload():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0
    for wl in ("hybrid6", "hybrid6_fma", "srif6", "mc_jerk3", "mc_robot_info", "mc_robot_sqrt", "vanilla32", "vanilla64"):
        assert wl in out.stdout


def test_reference_formulas_fma_spread_probe(oracle):
    """cpu_baseline.fma_spread: the oracle built with -ffp-contract=fast against the same oracle unfused on the bench's
    orbit scenario.  Small here (16 filters x 60 epochs); the point is that the two builds DO differ far above 1e-10 on
    this scenario -- the reason the bench headline runs in reference-order arithmetic."""
    import bench
    r = bench.oracle_fma_spread(nf=16, steps=60)
    assert r["filters"] == 16 and r["epochs"] == 60
    for key in ("state", "covariance"):
        assert np.isfinite(r[key]["median"]) and np.isfinite(r[key]["max"]) and r[key]["max"] >= r[key]["median"] >= 0.0
    assert r["covariance"]["median"] > 1e-8 and r["cond_final_covariance_median"] > 1e9
