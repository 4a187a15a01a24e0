#!/usr/bin/env python
"""Times the NLDKF kernels at the shapes beyond the TMA production path (n = 7, 8): HybridKF in production and in
reference-order (strict) arithmetic, and SRIF, on random well-conditioned per-filter streams, device-resident.
One JSON line per (kind, n, m).  Not part of bench.py's contract: a coverage measurement for DESIGN.md section 3."""
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

lib = gk.load()
dev = torch.device("cuda", 0)
nf, steps = int(os.environ.get("NF", 100000)), int(os.environ.get("STEPS", 100))


def run(kind, n, m, strict):
    g = torch.Generator(device=dev).manual_seed(7)
    Phi = (torch.eye(n, dtype=torch.float64, device=dev).reshape(1, n * n, 1)
           + 0.02 * torch.randn(steps, n * n, nf, dtype=torch.float64, device=dev, generator=g)).contiguous()
    Ht = torch.randn(steps, m * n, nf, dtype=torch.float64, device=dev, generator=g)
    real = torch.randn(steps, m, nf, dtype=torch.float64, device=dev, generator=g)
    comp = real + 0.05 * torch.randn(steps, m, nf, dtype=torch.float64, device=dev, generator=g)
    P0, R, Q = np.eye(n) * 10.0, np.eye(m) * 1e-2, np.eye(3) * 1e-12
    if kind == "srif":
        kf = gk.NewSRIF(np.zeros(n), P0, m, False, gk.NewNoiseless(Q, R), n_filters=nf)[0]
    else:
        kf = gk.NewHybridKF(np.zeros(n), P0, gk.NewNoiseless(Q, R), m, n_filters=nf)[0]
    if kind != "srif":
        kf.SetStrict(strict)
    flags = torch.from_numpy(np.array([L.F_MEAS | (L.F_EKF if (k >= 15 and kind != "srif") else 0) for k in range(steps)],
                                      dtype=np.uint8)).to(dev)
    out = L.Outputs()
    o_state = torch.zeros(n, nf, dtype=torch.float64, device=dev)
    o_cov = torch.zeros(n * n, nf, dtype=torch.float64, device=dev)
    status = torch.zeros(nf, dtype=torch.int32, device=dev)
    out.mem, out.every_step = L.DEVICE, 0
    out.state, out.covar, out.status = o_state.data_ptr(), o_cov.data_ptr(), status.data_ptr()
    ms = []
    for it in range(5):
        L.check(lib.gkb_reset(kf._h))
        L.check(lib.gkb_nl_run(kf._h, steps, flags.data_ptr(), Phi.data_ptr(), 0, Ht.data_ptr(), 0, real.data_ptr(),
                               comp.data_ptr(), None, L.DEVICE, C.byref(out)))
        torch.cuda.synchronize()
        if it >= 2:
            ms.append(lib.gkb_last_main_kernel_ms())
    t = sum(ms) / len(ms)
    ups = nf * steps / (t * 1e-3)
    bytes_unit = 8.0 * (n * n + m * n + 2 * m)
    print(json.dumps({"kind": kind, "strict": bool(strict), "n": n, "m": m, "filters": nf, "epochs": steps, "kernel_ms": t,
                      "updates_per_s": ups, "hbm_gbs": ups * bytes_unit / 1e9, "failed": int((status != 0).sum().item())}),
          flush=True)


for (n, m) in ((6, 2), (7, 2), (8, 2), (8, 3)):
    run("hybrid", n, m, False)
    run("hybrid", n, m, True)
    run("srif", n, m, False)
