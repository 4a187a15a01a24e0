#!/usr/bin/env python
"""Times gkb_batch_solve (batch.go:34-79 on the GPU: `steps` SetNextMeasurement accumulations + Solve() per filter) on the
hybrid6 bench streams: 10^5 batch filters x 200 measurement epochs, device-resident.  Prints one JSON line (not part of
bench.py's contract: BatchKF is a section-8(f) row).  Algorithmic HBM bytes: (m n + 2 m) x 8 = 128 B per measurement."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import gokalman_b200 as gk  # noqa: E402
from gokalman_b200 import _lib as L  # noqa: E402
from bench_hybrid import make_streams  # noqa: E402

if os.environ.get("GKB_BENCH_LIB"):
    L.LIB_PATH = os.path.abspath(os.environ["GKB_BENCH_LIB"])
lib = gk.load()
nf, steps = 100000, 200
n, m = int(os.environ.get("N", 6)), int(os.environ.get("M", 2))
dev = torch.device("cuda", 0)
if (n, m) == (6, 2):
    Phi, Ht, real, comp = make_streams(torch, nf, steps, 1234, dev)
    del Phi
else:  # other compiled shapes (N=, M= in the environment): random partials
    g = torch.Generator(device=dev).manual_seed(7)
    Ht = torch.randn(steps, m * n, nf, dtype=torch.float64, device=dev, generator=g)
    real = torch.randn(steps, m, nf, dtype=torch.float64, device=dev, generator=g)
    comp = real + 0.05 * torch.randn(steps, m, nf, dtype=torch.float64, device=dev, generator=g)
R = np.ascontiguousarray(np.diag([1e-6] * m))
xhat = torch.zeros(n, nf, dtype=torch.float64, device=dev)
P0 = torch.zeros(n * n, nf, dtype=torch.float64, device=dev)
status = torch.zeros(nf, dtype=torch.int32, device=dev)
ms = []
for it in range(6):
    L.check(lib.gkb_batch_solve(n, m, steps, nf, 0, R.ctypes.data_as(C.c_void_p), Ht.data_ptr(), 0, real.data_ptr(),
                                comp.data_ptr(), L.DEVICE, xhat.data_ptr(), P0.data_ptr(), status.data_ptr()))
    torch.cuda.synchronize()
    if it >= 2:
        ms.append(lib.gkb_last_kernel_ms())
t = sum(ms) / len(ms)
ups = nf * steps / (t * 1e-3)
print(json.dumps({"kernel": "batch_solve_kernel<%d,%d>" % (n, m), "kernel_ms": t, "measurements_per_s": ups,
                  "hbm_gbs": ups * 8 * (m * n + 2 * m) / 1e9, "bad": int((status != 0).sum().item())}))
