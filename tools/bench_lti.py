#!/usr/bin/env python
"""Times the batched LDKF.Update kernels (kernels_lti.cu) on the examples/jerkcar 4-state model: 10^6 filters x 200 steps,
per-filter measurement stream, final-estimate outputs, device-resident.  One JSON line per filter kind."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import fixtures as fx  # noqa: E402
import gokalman_b200 as gk  # noqa: E402
from gokalman_b200 import _lib as L  # noqa: E402

lib = gk.load()
nf, steps = 1000000, 200
dev = torch.device("cuda", 0)
f = fx.jerkcar4()
y = torch.randn(steps, 1, nf, dtype=torch.float64, device=dev)
u = torch.zeros(steps, 1, dtype=torch.float64, device=dev)
FLOPS = {"vanilla": 765.0, "information": 1033.0, "sqrt": 624.0}
for kind, ctor in (("vanilla", gk.NewVanilla), ("information", gk.NewInformationFromState), ("sqrt", gk.NewSquareRoot)):
    kf, _ = ctor(f["x0"], f["P0"], f["F"], f["G"], f["H2"], gk.NewNoiseless(f["Q"], f["Ra"]), n_filters=nf)
    xs = torch.zeros(4, nf, dtype=torch.float64, device=dev)
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
    ups = nf * steps / (t * 1e-3)
    print(json.dumps({"kind": kind, "kernel_ms": t, "updates_per_s": ups, "alg_tflops": ups * FLOPS[kind] / 1e12,
                      "failed": int((st != 0).sum().item())}))
