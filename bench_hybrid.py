"""bench.py workload `hybrid6` (BASELINE configs[3]): 10^5 six-state hybrid CKF->EKF filters with
per-filter, per-epoch Phi (6x6), Htilde (2x6), real and computed range / range-rate observations
streamed from HBM in SoA [epoch][component][filter] (416 B per filter-update in, state out only at
the end).  Synthetic orbit-determination-like inputs are generated ON THE DEVICE with torch (input
synthesis, not the measured path): Phi = I + dt*[[0, I],[G_k, 0]] with a random symmetric
gravity-gradient block, Htilde from random line-of-sight unit vectors."""
import ctypes as C
import os
import statistics
import time

import numpy as np

FLOPS_CKF, FLOPS_EKF, BYTES_IN = 2526.0, 2422.0, 416.0


def make_streams(torch, nf, steps, seed, device):
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    n, m, dt = 6, 2, 10.0
    Phi = torch.zeros(steps, n, n, nf, dtype=torch.float64, device=device)
    eye = torch.eye(n, dtype=torch.float64, device=device)
    Phi += eye[None, :, :, None]
    for i in range(3):
        Phi[:, i, 3 + i, :] += dt
    A = 1e-6 * torch.randn(steps, 3, 3, nf, dtype=torch.float64, device=device, generator=g)
    Gm = A + A.transpose(1, 2)
    Phi[:, 3:, :3, :] += dt * Gm
    Phi[:, :3, :3, :] += 0.5 * dt * dt * Gm
    del A, Gm
    los = torch.randn(steps, 3, nf, dtype=torch.float64, device=device, generator=g)
    los /= los.norm(dim=1, keepdim=True)
    Ht = torch.zeros(steps, m, n, nf, dtype=torch.float64, device=device)
    Ht[:, 0, :3, :] = los
    Ht[:, 1, 3:, :] = los
    Ht[:, 1, :3, :] = 1e-3 * torch.randn(steps, 3, nf, dtype=torch.float64, device=device, generator=g)
    real = torch.randn(steps, m, nf, dtype=torch.float64, device=device, generator=g)
    comp = real + 1e-3 * torch.randn(steps, m, nf, dtype=torch.float64, device=device, generator=g)
    return (Phi.reshape(steps, n * n, nf).contiguous(), Ht.reshape(steps, m * n, nf).contiguous(), real.contiguous(),
            comp.contiguous())


def run_ours_hybrid(args, rank, world, local):
    import torch
    import torch.distributed as dist
    import gokalman_b200 as gk
    from gokalman_b200 import _lib as L
    from bench import ClockSampler, fp64_peak, hbm_peak, measured_traffic

    lib = gk.load()
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nf = args.trials if args.trials != 1000000 else 100000
    steps = args.filter_steps if args.filter_steps != 1000 else 200
    n, m, q = 6, 2, 3
    dev = torch.device("cuda", local)
    Phi, Ht, real, comp = make_streams(torch, nf, steps, 1234 + rank, dev)
    flags_np = np.array([L.F_MEAS | (L.F_EKF if k >= 15 else 0) for k in range(steps)], dtype=np.uint8)  # hybrid_test.go:65
    flags = torch.from_numpy(flags_np).to(dev)
    P0 = np.diag([10, 10, 10, 1, 1, 1.0])
    R = np.diag([1e-6, 1e-6])
    Q = np.diag([1e-12] * 3)
    srif = args.workload == "srif6"
    if srif:  # srif_test.go:70-80: P0 = diag(50,50,50,1,1,1)
        P0 = np.diag([50, 50, 50, 1, 1, 1.0])
        flags_np = np.full(steps, L.F_MEAS, dtype=np.uint8)
        flags = torch.from_numpy(flags_np).to(dev)
        make = lambda: gk.NewSRIF(np.zeros(n), P0, m, False, gk.NewNoiseless(Q, R), n_filters=nf, device=local)[0]
    else:
        make = lambda: gk.NewHybridKF(np.zeros(n), P0, gk.NewNoiseless(Q, R), m, n_filters=nf, device=local)[0]
    kf = make()
    out_state = torch.zeros(n, nf, dtype=torch.float64, device=dev)
    out_cov = torch.zeros(n * n, nf, dtype=torch.float64, device=dev)
    status = torch.zeros(nf, dtype=torch.int32, device=dev)
    every = bool(int(os.environ.get("GKB_BENCH_EVERY_STEP", "0"))) and not srif
    if every:  # Estimate of every epoch streamed out (+336 B per update): what a smoothing pass consumes
        out_state = torch.zeros(steps, n, nf, dtype=torch.float64, device=dev)
        out_cov = torch.zeros(steps, n * n, nf, dtype=torch.float64, device=dev)
    out = L.Outputs()
    out.mem, out.every_step = L.DEVICE, int(every)
    out.state, out.covar, out.status = out_state.data_ptr(), out_cov.data_ptr(), status.data_ptr()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def step_device():
        L.check(lib.gkb_reset(kf._h))
        L.check(lib.gkb_nl_run(kf._h, steps, flags.data_ptr(), Phi.data_ptr(), 0, Ht.data_ptr(), 0, real.data_ptr(),
                               comp.data_ptr(), None, L.DEVICE, C.byref(out)))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    peak_tf, peak_src = fp64_peak() if rank == 0 else (None, None)
    hbm, hbm_src = hbm_peak()
    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kern_ms = []
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        flush.zero_()  # inputs (>= 8 GB) are far larger than L2 anyway
        ev[i][0].record()
        step_device()
        ev[i][1].record()
        kern_ms.append(lib.gkb_last_main_kernel_ms())
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    total_ms = torch.tensor([sum(a.elapsed_time(b) for a, b in ev)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    units = float(nf) * steps * world * args.steps
    value = units / (total_ms * 1e-3)
    bad = int((status != 0).sum().item())

    # ---- e2e: the public host-buffer API (pinned host arrays in, final estimate out), reduced epochs
    e_steps = min(steps, 50)
    hPhi = Phi[:e_steps].cpu().pin_memory().numpy()
    hHt = Ht[:e_steps].cpu().pin_memory().numpy()
    hreal = real[:e_steps].cpu().pin_memory().numpy()
    hcomp = comp[:e_steps].cpu().pin_memory().numpy()
    kf2 = make()
    kf2.RunBatch(flags_np[:e_steps], hPhi, hHt, hreal, hcomp, None, every_step=False, want=("state", "covar"))
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n_e2e = 2
    for _ in range(n_e2e):
        L.check(lib.gkb_reset(kf2._h))
        est = kf2.RunBatch(flags_np[:e_steps], hPhi, hHt, hreal, hcomp, None, every_step=False, want=("state", "covar"))
    e1.record()
    barrier()
    e2e_ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = float(nf) * e_steps * world * n_e2e / (float(e2e_ms.item()) * 1e-3)
    h2d = 8 * e_steps * nf * (36 + 12 + 4) + e_steps
    d2h = 8 * nf * (6 + 36) + 4 * nf
    if rank != 0:
        return None
    main_ms = statistics.mean(kern_ms)
    ups = float(nf) * steps / (main_ms * 1e-3)
    bytes_unit = BYTES_IN + (336.0 if every else 0.0)
    gbs = ups * bytes_unit / 1e9
    flops = 2168.0 if srif else FLOPS_EKF
    tf = ups * flops / 1e12
    # The input streams are read exactly once (ncu: DRAM traffic = algorithmic bytes), so the HBM roof is the one the
    # headline fraction is quoted against.  The FP64 figure below counts the DENSE flops the reference executes (SURVEY
    # App. B); the kernels execute fewer (symmetric / triangular shortcuts), so it is context, not a pipe utilisation.
    bound_hbm = True
    line = {
        "metric": "filter-updates/sec (batch x steps, FP64)", "value": value, "unit": "filter-updates/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": ("srif6: 6-state SRIF, range + range-rate, per-filter Phi/Htilde streams (BASELINE configs[3])" if srif else
                                "hybrid6: 6-state hybrid CKF->EKF, range + range-rate, per-filter Phi/Htilde streams "
                                "(BASELINE configs[3])"), "filters_per_gpu": nf, "epochs": steps, "n": 6, "m": 2,
                   "ekf_after": 15, "outputs": "state + covariance of every epoch" if every else "final state + covariance only", "failed_filters": bad,
                   "l2": "inputs (%.1f GB) exceed L2; flushed anyway" % (nf * steps * BYTES_IN / 1e9)},
        "roofline": {"bound": "hbm" if bound_hbm else "fp64", "achieved": gbs if bound_hbm else tf,
                     "peak": hbm if bound_hbm else peak_tf, "unit": "GB/s" if bound_hbm else "TFLOP/s",
                     "frac": (gbs / hbm) if bound_hbm else (tf / peak_tf),
                     "traffic": measured_traffic(args.workload, nf == 100000 and steps == 200),
                     "algorithmic_bytes": float(nf) * steps * BYTES_IN,
                     "kernel": "nl_run_wtma_sched_kernel<6,2,SRIF>" if srif else "nl_run_wtma_sched_kernel<6,2> (warp-private TMA tensor-map pipelines, persistent chunk scheduler)", "kernel_ms": main_ms,
                     "hbm": {"achieved_gbs": gbs, "peak_gbs": hbm, "frac": gbs / hbm, "bytes_per_unit": bytes_unit, "source": hbm_src},
                     "fp64": {"achieved_tflops": tf, "peak_tflops": peak_tf, "frac": tf / peak_tf, "flops_per_unit": flops,
                              "source": peak_src}},
        "e2e": {"value": e2e_value, "unit": "filter-updates/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "HybridKF.RunBatch (pinned host buffers), %d epochs" % e_steps},
        "gpu_launches": args.steps, "clocks": clocks, "wall_s": wall,
    }
    return line
