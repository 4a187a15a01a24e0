#!/usr/bin/env python
"""Times gkb_smooth_all (hybrid.go:209-238 on the GPU) on the hybrid6 bench streams: 10^5 filters x 200 epochs,
device-resident.  Prints one JSON line (not part of bench.py's contract: SmoothAll is a section-8(f) row)."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import gokalman_b200 as gk  # noqa: E402
from gokalman_b200 import _lib as L  # noqa: E402
from bench_hybrid import make_streams  # noqa: E402

lib = gk.load()
nf, steps, n = 100000, 200, 6
dev = torch.device("cuda", 0)
Phi, Ht, real, comp = make_streams(torch, nf, steps, 1234, dev)
del Ht, real, comp
xs = torch.randn(steps, n, nf, dtype=torch.float64, device=dev)
A = torch.randn(steps, n, n, nf, dtype=torch.float64, device=dev)
Ps = (A.transpose(1, 2) * 0 + torch.einsum("sijf,skjf->sikf", A, A)).reshape(steps, n * n, nf).contiguous()
del A
status = torch.zeros(nf, dtype=torch.int32, device=dev)
ms = []
for it in range(6):
    L.check(lib.gkb_smooth_all(n, steps, nf, 0, Phi.data_ptr(), 0, xs.data_ptr(), Ps.data_ptr(), L.DEVICE, status.data_ptr()))
    torch.cuda.synchronize()
    if it >= 2:
        ms.append(lib.gkb_last_kernel_ms())
t = sum(ms) / len(ms)
ups = nf * (steps - 1) / (t * 1e-3)
print(json.dumps({"kernel": "smooth_all_kernel<6>", "kernel_ms": t, "filter_steps_per_s": ups,
                  "hbm_gbs": ups * (288 + 336) / 1e9, "bad": int((status != 0).sum().item())}))
