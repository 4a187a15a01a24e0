"""Sweep of the chunk count of the persistent NLDKF scheduler (GKB_NL_CHUNKS) on the bench's hybrid6 / srif6
configuration (10^5 filters x 1000 epochs of device-synthesised statOD streams): kernel ms per chunk count."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import gokalman_b200 as gk
    from gokalman_b200 import _lib as L
    from bench_hybrid import make_streams_od, SIGMA
    lib = gk.load()
    nf, steps = 100000, int(os.environ.get("SWEEP_EPOCHS", "1000"))
    Phi, Ht, real, comp, scn, orbit0 = make_streams_od(torch, L, lib, nf, steps, 1234, 0)
    dev = torch.device("cuda", 0)
    for kind in sys.argv[1:] or ["hybrid"]:
        R, Q = np.diag([SIGMA ** 2] * 2), np.diag([1e-12] * 3)
        if kind == "srif":
            kf = gk.NewSRIF(np.zeros(6), np.diag([50, 50, 50, 1, 1, 1.0]), 2, False, gk.NewNoiseless(Q, R), n_filters=nf)[0]
            flags_np = np.full(steps, L.F_MEAS, dtype=np.uint8)
        else:
            kf = gk.NewHybridKF(np.zeros(6), np.diag([10, 10, 10, 1, 1, 1.0]), gk.NewNoiseless(Q, R), 2, n_filters=nf)[0]
            flags_np = np.ascontiguousarray(scn.flags)
        flags = torch.from_numpy(flags_np).to(dev)
        xs = torch.zeros(6, nf, dtype=torch.float64, device=dev)
        Ps = torch.zeros(36, nf, dtype=torch.float64, device=dev)
        out = L.Outputs()
        out.mem, out.every_step = L.DEVICE, 0
        out.state, out.covar = xs.data_ptr(), Ps.data_ptr()
        order = os.environ.get("SWEEP_CHUNKS")  # e.g. "default,6,default,6,3,5,7": an interleaved A/B (the GPU is power-capped
        # when this kernel runs back to back, so the position in the sequence matters as much as the chunk count)
        for chunks in (order.split(",") if order else ["default"] + [str(c) for c in (2, 3, 4, 6, 8, 12, 16, 24, 32, 48, 64)]):
            if chunks == "default":
                os.environ.pop("GKB_NL_CHUNKS", None)
            else:
                os.environ["GKB_NL_CHUNKS"] = chunks
            ms = []
            for it in range(6):
                L.check(lib.gkb_reset(kf._h))
                L.check(lib.gkb_nl_run(kf._h, steps, flags.data_ptr(), Phi.data_ptr(), 0, Ht.data_ptr(), 0, real.data_ptr(),
                                       comp.data_ptr(), None, L.DEVICE, C.byref(out)))
                torch.cuda.synchronize()
                if it >= 2:
                    ms.append(lib.gkb_last_kernel_ms())
            best = min(ms)
            print("%s chunks=%s: %.3f ms (min of 4; mean %.3f) -> %.3e updates/s, %.1f%% of 6455.6 GB/s" %
                  (kind, chunks, best, sum(ms) / len(ms), nf * steps / (best * 1e-3), 100 * nf * steps * 416 / (best * 1e-3) / 6455.6e9), flush=True)


if __name__ == "__main__":
    main()
