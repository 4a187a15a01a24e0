#!/usr/bin/env python
"""Times the batched LDKF.Update kernels (kernels_lti.cu) over the compiled shapes: vanilla / information / square-root at
(n, m) in SHAPES on a random stable LTI model, NF filters x STEPS steps, per-filter measurement stream, final-estimate
outputs, device-resident.  One JSON line per (kind, n, m).  GKB_BENCH_LIB=<path> times another build of the library
(A/B of a kernel change: same process, same inputs).  Not part of bench.py's contract."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import gokalman_b200 as gk  # noqa: E402
from gokalman_b200 import _lib as L  # noqa: E402

if os.environ.get("GKB_BENCH_LIB"):
    L.LIB_PATH = os.path.abspath(os.environ["GKB_BENCH_LIB"])
lib = gk.load()
dev = torch.device("cuda", 0)
nf, steps = int(os.environ.get("NF", 400000)), int(os.environ.get("STEPS", 100))
SHAPES = [(3, 1), (4, 1), (4, 2), (6, 2), (6, 3), (7, 2), (8, 1), (8, 2), (8, 3)]


def model(n, m, rng):
    A = rng.uniform(-1, 1, (n, n)) / n
    F = np.eye(n) + 0.05 * A
    G = np.zeros((n, 1))
    H = np.linalg.qr(rng.standard_normal((n, n)))[0][:m]
    Lq = 0.03 * rng.standard_normal((n, n))
    return F, G, H, Lq @ Lq.T + 1e-4 * np.eye(n), np.diag(rng.uniform(0.1, 1.0, m))


for n, m in SHAPES:
    rng = np.random.default_rng(100 * n + m)
    F, G, H, Q, R = model(n, m, rng)
    y = torch.randn(steps, m, nf, dtype=torch.float64, device=dev)
    u = torch.zeros(steps, 1, dtype=torch.float64, device=dev)
    for kind, ctor in (("vanilla", gk.NewVanilla), ("information", gk.NewInformationFromState), ("sqrt", gk.NewSquareRoot)):
        kf, _ = ctor(np.zeros(n), np.eye(n), F, G, H, gk.NewNoiseless(Q, R), n_filters=nf)
        xs = torch.zeros(n, nf, dtype=torch.float64, device=dev)
        st = torch.zeros(nf, dtype=torch.int32, device=dev)
        out = L.Outputs()
        out.mem, out.every_step, out.state, out.status = L.DEVICE, 0, xs.data_ptr(), st.data_ptr()
        ms = []
        for it in range(5):
            L.check(lib.gkb_reset(kf._h))
            L.check(lib.gkb_update(kf._h, steps, y.data_ptr(), 0, u.data_ptr(), L.DEVICE, C.byref(out)))
            torch.cuda.synchronize()
            if it >= 2:
                ms.append(lib.gkb_last_kernel_ms())
        t = sum(ms) / len(ms)
        print(json.dumps({"kind": kind, "n": n, "m": m, "filters": nf, "steps": steps, "kernel_ms": round(t, 4),
                          "updates_per_s": nf * steps / (t * 1e-3), "failed": int((st != 0).sum().item()),
                          "checksum": float(xs.double().abs().sum().item())}), flush=True)
        del kf
