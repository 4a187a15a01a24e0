#!/usr/bin/env python
"""bench.py -- filter-updates/sec of the batched Kalman hot path on N B200s (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload W] [--impl ours|reference] [--no-sub]

One "step" = one pass of the hot path over one batch of synthetic input.  Workloads:
  hybrid6 (DEFAULT: BASELINE configs[3], the north star's Target config): 10^5 six-state hybrid CKF->EKF filters per
      GPU x 1000 epochs, per-filter per-epoch Phi / Htilde / observations (416 B per filter-update) resident in HBM:
      41.6 GB per GPU, synthesised on the device from a LEO statOD scenario (gkb_od_synthesize: two-body + J2 STM,
      range / range-rate partials, App. D constants).  `value` = the REFERENCE-ORDER kernel (gkb_set_strict: dense
      products in the written order, no FMA contraction, dense Joseph form): bit-identical to the CPU oracle on these
      streams, which drive the covariance to cond ~ 1e13 -- nothing but the reference's own rounding sequence holds the
      north star's 1e-10 there.  `e2e` = the fused OD run through the public API in the same arithmetic (host: initial
      orbits in, final estimates out); the same line carries the host-fed-streams figure against the measured PCIe
      bandwidth.  Multi-GPU: disjoint filter ranges, no collective.
  hybrid6_fma: the same run through the production (FMA, TMA-pipelined) kernel -- 2.5x faster, HBM-bound, equal to the
      reference up to the rounding sensitivity of the run (`production_vs_strict` in its record).
  srif6 (configs[3], SRIF arm), mc_jerk3 (configs[1]: 10^6 Monte Carlo trials x 1000 steps + chi-square, one NCCL
      all-reduce of the NEES / NIS sums), mc_robot_info / mc_robot_sqrt (configs[2]), vanilla32 / vanilla64 (configs[4]).
Unless --no-sub is given the default run appends `"sub"`: records of hybrid6_fma, srif6, mc_jerk3 (weak and, at N > 1, strong scaling) and vanilla32, each with its own clocks.
Prints ONE JSON line (rank 0).  --impl reference times the CPU oracle on the same workload (the reference itself is
Go + un-vendored gonum and cannot be built in this image).
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

import fixtures as fx  # noqa: E402

FLOPS_PER_UNIT = {"mc_jerk3": 537.0, "hybrid6": 2526.0}  # BASELINE.md section 3 (algorithmic, dense)
BYTES_PER_UNIT = {"mc_jerk3": 0.0, "hybrid6": 416.0}
SEED = 0x5EED

# Monte Carlo + chi-square workloads: (fixture, tested kind, NEES, NIS, controls, algorithmic flop per (trial, step))
#   mc_jerk3       BASELINE configs[1]: truth 51 + vanilla 374 + NEES 81 + NIS 31 = 537
#   mc_robot_info  BASELINE configs[2]: truth 29 + information 173 + NEES 30 = 232 (NIS of an information filter needs
#                  n == m in the reference, information.go:272-274 -- the robot has n = 2, m = 1, so NEES only)
#   mc_robot_sqrt  BASELINE configs[2]: truth 29 + square-root 123 + P = S S^T 16 + NEES 30 + NIS 19 = 217
MC_WORKLOADS = {
    "mc_jerk3": dict(fixture="jerk3", kind="VANILLA", nees=1, nis=1, controls="zero", flops=537.0,
                     label="mc_jerk3: jerkcar 3-state vanilla KF Monte Carlo + chi-square (BASELINE configs[1])",
                     kernel="mc_chisquare_kernel<3,1,VanillaTested>"),
    "mc_robot_info": dict(fixture="robot_1d", kind="INFORMATION", nees=1, nis=0, controls="robot", flops=232.0,
                          label="mc_robot_info: examples/robot 2-state information filter Monte Carlo + NEES (BASELINE configs[2])",
                          kernel="mc_chisquare_kernel<2,1,InfoTested>"),
    "mc_robot_sqrt": dict(fixture="robot_1d", kind="SQRT", nees=1, nis=1, controls="robot", flops=217.0,
                          label="mc_robot_sqrt: examples/robot 2-state square-root filter Monte Carlo + chi-square (BASELINE configs[2])",
                          kernel="mc_chisquare_kernel<2,1,SqrtTested>"),
}


# ------------------------------------------------------------------------------------------------
# helpers
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock / power / clock-event reasons sampled DURING the timed region.  In-process NVML at a 5 ms period
    (the timed regions here are tenths of a second: `nvidia-smi -lms 100` would see one or two samples), with the
    nvidia-smi loop of B200_PROFILING.md as the fallback when NVML is unavailable."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, gpu_index, period_s=0.005):
        self.idx, self.period, self.proc, self.lines = gpu_index, period_s, None, []
        self.samples, self._stop, self.t, self.h, self.src = [], threading.Event(), None, None, None

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        try:  # CUDA_VISIBLE_DEVICES may renumber the devices: go through the UUID
            import torch
            uuid = str(torch.cuda.get_device_properties(self.idx).uuid)
            uuid = uuid if uuid.startswith("GPU-") else "GPU-" + uuid
            return pynvml, pynvml.nvmlDeviceGetHandleByUUID(uuid.encode() if hasattr(uuid, "encode") else uuid)
        except Exception:
            return pynvml, pynvml.nvmlDeviceGetHandleByIndex(self.idx)

    def start(self):
        try:
            self.nv, self.h = self._nvml_handle()
            self.max_sm = float(self.nv.nvmlDeviceGetMaxClockInfo(self.h, self.nv.NVML_CLOCK_SM))
            self.src = "nvml, %g ms period" % (1e3 * self.period)
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.h = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.src = "nvidia-smi -lms 20"
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv, h = self.nv, self.h
        reasons_fn = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self._stop.is_set():
            try:
                self.samples.append((float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)),
                                     nv.nvmlDeviceGetPowerUsage(h) / 1000.0, int(reasons_fn(h))))
            except Exception:
                pass
            time.sleep(self.period)

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.h is not None:
            self._stop.set()
            self.t.join(timeout=1.0)
            if not self.samples:
                return {"sm_mhz": None, "sm_max_mhz": self.max_sm, "samples": 0, "reasons": ["no samples"], "source": self.src}
            sm = [x[0] for x in self.samples]
            reasons = sorted({name for _, _, r in self.samples for bit, name in self.REASONS if r & bit})
            load = [x for x in sm if x >= 0.5 * self.max_sm] or sm
            return {"sm_mhz": statistics.median(load), "sm_max_mhz": self.max_sm, "power_w_max": max(x[1] for x in self.samples),
                    "samples": len(sm), "reasons": reasons, "source": self.src}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["nvml and nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 8:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2])); pw.append(float(p[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["no samples"], "source": self.src}
        load = [s for s in sm if s >= 0.5 * max(mx)] or sm
        return {"sm_mhz": statistics.median(load), "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons), "source": self.src}


# The live FP64 peak measurement (tools/peak_fp64) is seconds of back-to-back DFMA / DMMA launches: it must never run
# before a timed region (it would pre-heat the GPU into its power cap).  Records are therefore built with their
# FP64-peak-dependent fields left to a fix-up that runs once every timed region of the process is over.
_PEAK_FIXUPS = []


def when_fp64_peak_known(fn):
    """fn(dfma_tflops, dmma_tflops, source) fills the peak-dependent fields of a record; called at the end of main()."""
    _PEAK_FIXUPS.append(fn)


def apply_fp64_peak_fixups():
    # bench.py runs as __main__ while bench_hybrid / bench_tile import it as `bench`: two module objects, two lists
    lists = [m._PEAK_FIXUPS for m in {sys.modules.get("__main__"), sys.modules.get("bench")} if m is not None and hasattr(m, "_PEAK_FIXUPS")]
    if not any(lists):
        return
    dfma, dmma, src = fp64_peaks_all()
    for lst in lists:
        for fn in lst:
            fn(dfma, dmma, src)
        lst.clear()


def settle_clocks(step_fn, sync_fn, seconds=0.8, max_steps=400):
    """Extra untimed warm-up: the workload's own step repeated for ~`seconds` so that the SM / memory clocks have left
    their idle state and the GPU is in the sustained regime before the timed region starts (a 40 ms warm-up of an
    HBM-bound kernel is over before the clocks have ramped: the NVML samples then show 1.6 GHz with no reason)."""
    t0 = time.perf_counter()
    n = 0
    while time.perf_counter() - t0 < seconds and n < max_steps:
        step_fn()
        if n % 8 == 7:
            sync_fn()
        n += 1
    sync_fn()
    return n


def fp64_peak():
    """DFMA TFLOP/s: measured live with tools/peak_fp64 when built, else the committed measurement."""
    dfma, _, src = fp64_peaks_all()
    return dfma, src + ", DFMA sustained"


_PEAK_CACHE = {}


def fp64_peaks_all():
    """(dfma_sustained, dmma_burst, source) in TFLOP/s: tools/peak_fp64 run once per process, else the committed
    measurement."""
    if "v" not in _PEAK_CACHE:
        d, src = None, None
        exe = os.path.join(ROOT, "tools", "peak_fp64")
        if os.path.exists(exe):
            try:
                r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
                d = json.loads(r.stdout.strip().splitlines()[-1])
                src = "measured live: tools/peak_fp64 (MEASURED_PEAKS.json has no FP64 entry)"
            except Exception:
                d = None
        if d is None:
            try:
                d = json.load(open(os.path.join(ROOT, "profiles", "r01_peak_fp64.json")))
                src = "profiles/r01_peak_fp64.json (measured on this pool's B200)"
            except Exception:
                d, src = {"dfma_tflops_sustained": 34.2, "dmma_m8n8k4_tflops_burst": 37.1}, "fallback constants (profiles/r01_peak_fp64.json)"
        _PEAK_CACHE["v"] = (d["dfma_tflops_sustained"], d.get("dmma_m8n8k4_tflops_burst", 37.1), src)
    return _PEAK_CACHE["v"]


def measured_traffic(workload, is_default_config):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel, from the committed ncu
    capture of this same bench configuration (tools/measure_traffic.sh -> profiles/r01_traffic.json); None when
    the run uses another size."""
    if not is_default_config:
        return None
    try:
        # the hybrid / SRIF bench configuration changed in round 2 (1000 epochs): only the round-2 capture applies
        names = ("r02_traffic.json",) if workload.startswith(("hybrid", "srif")) else ("r02_traffic.json", "r01_traffic.json")
        for name in names:
            d = json.load(open(os.path.join(ROOT, "profiles", name)))
            if workload in d and d[workload].get("traffic") is not None:
                return d[workload]["traffic"]
        return None
    except Exception:
        return None


def hbm_peak():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "MEASURED_PEAKS.json (of measured)"
    except Exception:
        return 6650.0, "B200_PROFILING.md fallback (of fallback)"


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ------------------------------------------------------------------------------------------------
# workloads
# ------------------------------------------------------------------------------------------------
def mc_model(workload="mc_jerk3"):
    # jerk3: helper_test.go:17-22 + montecarlo_test.go:12-19; robot_1d: examples/robot/main.go:16-26
    return getattr(fx, MC_WORKLOADS[workload]["fixture"])()


def mc_controls(workload, steps):
    if MC_WORKLOADS[workload]["controls"] == "robot":
        return fx.robot_controls(steps)  # examples/robot/main.go:36-39
    return np.zeros((steps, 1))          # montecarlo.go:98-104: a single control vector means zero controls


def mc_config_struct(L, f, trials, trial_offset, steps, device, workload="mc_jerk3"):
    """gkb_mc_config for the device-resident leg (PHILOX noise)."""
    spec = MC_WORKLOADS[workload]
    cfg = L.McConfig()
    keep = {k: np.ascontiguousarray(np.asarray(f[k], dtype=np.float64)) for k in ("F", "G", "H", "Q", "R", "x0", "P0")}
    keep["u"] = np.ascontiguousarray(mc_controls(workload, steps))
    cfg.kind, cfg.n, cfg.m, cfg.c = getattr(L, spec["kind"]), keep["F"].shape[0], keep["H"].shape[0], 1
    cfg.F, cfg.G, cfg.H, cfg.Q, cfg.R = (keep[k].ctypes.data for k in ("F", "G", "H", "Q", "R"))
    cfg.x0_truth = cfg.x0_filter = keep["x0"].ctypes.data
    cfg.P0 = keep["P0"].ctypes.data
    cfg.trials, cfg.trial_offset, cfg.steps = trials, trial_offset, steps
    cfg.controls = keep["u"].ctypes.data
    cfg.noise_mode, cfg.seed = L.NOISE_PHILOX, SEED
    cfg.with_nees, cfg.with_nis = spec["nees"], spec["nis"]
    cfg.device = device
    return cfg, keep


def run_ours_mc(args, rank, world, local, workload=None, sub=False, strong=False):
    """strong=False: `--trials` Monte Carlo trials PER GPU (weak scaling); strong=True: `--trials` trials in TOTAL,
    sharded over the ranks by contiguous global trial ranges (Philox is keyed by the global trial index)."""
    import torch
    import torch.distributed as dist
    import gokalman_b200 as gk
    from gokalman_b200 import _lib as L
    from gokalman_b200.sharding import shard_range

    lib = gk.load()
    torch.cuda.set_device(local)
    total_trials, steps = args.trials, args.filter_steps
    wl = workload or args.workload
    spec = MC_WORKLOADS[wl]
    f = mc_model(wl)
    if strong:
        lo, hi = shard_range(total_trials, rank, world)
        trials, offset, all_trials = hi - lo, lo, total_trials
    else:
        trials, offset, all_trials = total_trials, rank * total_trials, total_trials * world
    n_steps = args.steps if not sub else max(3, min(args.steps, 10))
    n_warm = args.warmup if not sub else 3
    cfg, keep = mc_config_struct(L, f, trials, offset, steps, local, wl)
    sums = torch.zeros(2, steps, dtype=torch.float64, device="cuda")
    out = L.McOutputs()
    out.mem, out.sums_only = L.DEVICE, 1
    out.nis, out.nees = sums[0].data_ptr(), sums[1].data_ptr()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def step_device():
        L.check(lib.gkb_mc_chisquare(C.byref(cfg), C.byref(out)))
        if world > 1:
            dist.all_reduce(sums)  # the one collective of the path: 2 x steps doubles (NCCL)
        return sums / float(all_trials)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(n_warm):
        flush.zero_()
        means = step_device()
    barrier()
    settle_clocks(step_device, torch.cuda.synchronize)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_steps)]
    kern_ms = []
    barrier()
    t_wall0 = time.perf_counter()
    for i in range(n_steps):
        flush.zero_()  # L2 flush between timed iterations (outside the per-step event bracket)
        ev[i][0].record()
        means = step_device()
        ev[i][1].record()
        kern_ms.append(lib.gkb_last_main_kernel_ms())  # CUDA events on the launch stream, inside the lib
    barrier()
    wall = time.perf_counter() - t_wall0
    clocks = sampler.stop() if rank == 0 else None
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)  # max over ranks
    total_ms = float(total_ms.item())
    units = float(all_trials) * steps * n_steps
    value = units / (total_ms * 1e-3)
    nis_mean, nees_mean = float(means[0].mean().item()), float(means[1].mean().item())
    if sub:  # a sub-record of the default run: device-resident figure only
        if rank != 0:
            return None
        main_ms = statistics.mean(kern_ms)
        achieved_tf = spec["flops"] * float(trials) * steps / (main_ms * 1e-3) / 1e12
        peak_tf = peak_src = None
        rec = {"value": value, "unit": "filter-updates/s", "n_gpus": world, "steps": n_steps, "warmup": n_warm,
                "ms_per_step": total_ms / n_steps, "scaling": "strong" if strong else "weak",
                "config": {"workload": spec["label"], "trials_total": all_trials, "trials_this_gpu": trials, "filter_steps": steps,
                           "collective": "one NCCL all-reduce of 2 x %d doubles per step" % steps if world > 1 else "none (1 GPU)",
                           "nis_mean": nis_mean, "nees_mean": nees_mean},
                "roofline": {"bound": "fp64", "achieved": achieved_tf, "peak": None, "unit": "TFLOP/s", "frac": None,
                             "kernel": spec["kernel"], "kernel_ms": main_ms, "flops_per_unit": spec["flops"], "peak_source": None},
                "allreduce_and_glue_ms": total_ms / n_steps - main_ms,
                "gpu_launches": int(lib.gkb_last_kernel_launches()) * n_steps, "clocks": clocks}
        rec["roofline"]["frac"] = None

        def fix(dfma, dmma, src, r=rec["roofline"]):
            r["peak"], r["frac"], r["peak_source"] = dfma, r["achieved"] / dfma, src + ", DFMA sustained"
        when_fp64_peak_known(fix)
        return rec

    # ---- e2e: the public API with HOST buffers (model + controls in, NIS/NEES means out), every step
    controls = [np.zeros(1)] if spec["controls"] == "zero" else list(mc_controls(wl, steps))
    def step_e2e():
        runs = gk.NewMonteCarloRuns(trials, steps, 1, controls, mckf, trial_offset=offset)
        nis, nees = gk.NewChiSquare(chikf, runs, controls, bool(spec["nees"]), bool(spec["nis"]))
        if world > 1:
            t = torch.from_numpy(np.stack([nis, nees]) * trials).cuda()
            dist.all_reduce(t)
            nis, nees = (t / float(all_trials)).cpu().numpy()
        return nis, nees
    mckf, _ = gk.NewPurePredictorVanilla(f["x0"], f["P0"], f["F"], f["G"], f["H"], gk.NewAWGN(f["Q"], f["R"], seed=SEED),
                                         device=local)
    make_tested = {"VANILLA": gk.NewVanilla, "INFORMATION": gk.NewInformationFromState, "SQRT": gk.NewSquareRoot}[spec["kind"]]
    chikf, _ = make_tested(f["x0"], f["P0"], f["F"], f["G"], f["H"], gk.NewNoiseless(f["Q"], f["R"]), device=local)
    step_e2e()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n_e2e = max(1, min(args.steps, 5))
    for _ in range(n_e2e):
        nis_h, nees_h = step_e2e()
    e1.record()
    barrier()
    e2e_ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = float(all_trials) * steps * n_e2e / (float(e2e_ms.item()) * 1e-3)
    nn, mm = keep["F"].shape[0], keep["H"].shape[0]
    h2d = 8 * (nn * nn + nn + mm * nn + nn * nn + mm * mm + nn + nn + nn * nn + steps * 1)  # F,G,H,Q,R,x0,x0,P0 + controls
    d2h = 8 * 2 * steps + 4                                   # NIS, NEES means + the error word
    assert np.allclose(nees_h.mean(), nees_mean, rtol=1e-9), (nees_h.mean(), nees_mean)

    if rank != 0:
        return None
    main_ms = statistics.mean(kern_ms)
    achieved_tf = spec["flops"] * float(trials) * steps / (main_ms * 1e-3) / 1e12
    line = {
        "metric": "filter-updates/sec (batch x steps, FP64)", "value": value, "unit": "filter-updates/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
        "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": spec["label"],
                   "trials_per_gpu": trials, "trials_total": all_trials, "filter_steps": steps, "n": nn, "m": mm, "c": 1, "noise": "philox4x32-10 + table inverse normal CDF, in-kernel",
                   "sharding": "trials split by rank, one NCCL all-reduce of 2 x %d doubles per step" % steps,
                   "l2": "flushed between timed iterations (256 MiB memset)", "nis_mean": nis_mean, "nees_mean": nees_mean},
        "roofline": {"bound": "fp64", "achieved": achieved_tf, "peak": None, "unit": "TFLOP/s",
                     "frac": None,
                     "traffic": measured_traffic(wl, trials == 1000000 and steps == 1000),
                     "kernel": spec["kernel"], "kernel_ms": main_ms,
                     "flops_per_unit": spec["flops"], "peak_source": None,
                     "note": "%g algorithmic flop per (trial, step) per SURVEY App. B; RNG / inverse-CDF work not counted" % spec["flops"]},
        "e2e": {"value": e2e_value, "unit": "filter-updates/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "NewMonteCarloRuns + NewChiSquare (host buffers)"},
        # fused MC kernel + finish kernel per step, as counted by the library (the model-setup kernel runs once, before
        # the timed region: its result is cached); torch's memset / division and NCCL are not ours and not counted
        "gpu_launches": int(lib.gkb_last_kernel_launches()) * args.steps,
        "clocks": clocks,
        "wall_s": wall,
        "step_ms": [round(x, 3) for x in step_ms],
    }

    def fix(dfma, dmma, src, r=line["roofline"]):
        r["peak"], r["frac"], r["peak_source"] = dfma, r["achieved"] / dfma, src + ", DFMA sustained"
    when_fp64_peak_known(fix)
    return line


# ------------------------------------------------------------------------------------------------
# CPU baseline / reference arm
# ------------------------------------------------------------------------------------------------
def oracle_mc_rate(trials, steps, threads, workload="mc_jerk3"):
    from oracle import gko
    gko.build()
    spec = MC_WORKLOADS[workload]
    f = mc_model(workload)
    ctrl = None if spec["controls"] == "zero" else mc_controls(workload, steps)
    t0 = time.perf_counter()
    r = gko.mc_chisquare(getattr(gko, spec["kind"]), f["F"], f["G"], f["H"], f["Q"], f["R"], f["x0"], f["x0"], f["P0"],
                         trials, steps, controls=ctrl, seed=SEED, threads=threads, with_nees=bool(spec["nees"]),
                         with_nis=bool(spec["nis"]))
    dt = time.perf_counter() - t0
    return trials * steps / dt, dt, float(r["NEES"].mean())


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def cpu_baseline(target_s=12.0, workload="mc_jerk3"):
    cores = os.cpu_count() or 1
    steps = 1000
    rate, _, _ = oracle_mc_rate(cores * 4, steps, cores, workload)  # calibration
    trials = max(cores, int(rate * target_s / steps / cores) * cores)
    single, _, _ = oracle_mc_rate(max(8, trials // (cores * 8)), steps, 1, workload)  # ~1.5 s on one thread
    rate, dt, _ = oracle_mc_rate(trials, steps, cores, workload)
    return {"value": rate, "unit": "filter-updates/s", "cores": cores, "kind": "port", "single_thread_value": single,
            "cpu_model": cpu_model(), "omp_num_threads": cores,
            "sample": "%d trials x %d steps of %s (%.1f s), C oracle restatement with OpenMP over trials; "
                      "the Go/gonum reference cannot be built here (no Go toolchain)" % (trials, steps, workload, dt)}


def np_od_streams(nf, steps, seed):
    """numpy twin of bench_hybrid.make_streams (same construction, host side) for the CPU arm."""
    rng = np.random.default_rng(seed)
    n, m, dt = 6, 2, 10.0
    Phi = np.zeros((steps, n, n, nf))
    Phi += np.eye(n)[None, :, :, None]
    for i in range(3):
        Phi[:, i, 3 + i, :] += dt
    A = 1e-6 * rng.standard_normal((steps, 3, 3, nf))
    Gm = A + A.transpose(0, 2, 1, 3)
    Phi[:, 3:, :3, :] += dt * Gm
    Phi[:, :3, :3, :] += 0.5 * dt * dt * Gm
    los = rng.standard_normal((steps, 3, nf))
    los /= np.linalg.norm(los, axis=1, keepdims=True)
    Ht = np.zeros((steps, m, n, nf))
    Ht[:, 0, :3, :] = los
    Ht[:, 1, 3:, :] = los
    Ht[:, 1, :3, :] = 1e-3 * rng.standard_normal((steps, 3, nf))
    real = rng.standard_normal((steps, m, nf))
    comp = real + 1e-3 * rng.standard_normal((steps, m, nf))
    return Phi.reshape(steps, n * n, nf), Ht.reshape(steps, m * n, nf), real, comp


def oracle_filter_rate(workload, nf, steps, threads, reps=1):
    """CPU oracle over `nf` independent filters x `steps` (OpenMP over filters): hybrid6 / srif6 / vanilla32.  The
    sample is bounded by host memory (its input streams), so `reps` passes over the same streams make up the CPU time;
    returns (updates/s, seconds of filter work -- input synthesis excluded)."""
    from oracle import gko
    gko.build()
    if workload in ("vanilla32", "vanilla64"):
        f = fx.synth_lti(int(workload[7:]), 8, seed=5)
        y = np.random.default_rng(4321).standard_normal((steps, nf, 8))
        t0 = time.perf_counter()
        for _ in range(reps):
            gko.run_vanilla_batch(f["x0"], f["P0"], f["F"], f["H"], f["Q"], f["R"], y, threads=threads, want_covar=False)
        dt = time.perf_counter() - t0
    else:
        # the bench's statOD scenario (bench_hybrid.od_scenario), streams made by the oracle's own synthesis
        from gokalman_b200 import od  # host-side numpy tables only: no engine call on this path
        scn = od.Scenario(steps, 10.0, od.leo_truth0(), always_track=True, theta0=2.5)
        orbit0 = od.perturbed_orbits(od.leo_truth0(), nf, sigma_r=1.0, sigma_v=1e-3, seed=1234)
        Phi, Ht, real, comp, _ = gko.od_synth(scn.mu, scn.j2, scn.re, scn.dt, orbit0, scn.station, scn.truth_obs, 1e-3, 1e-3, 1234)
        srif = workload == "srif6"
        P0 = np.diag([50, 50, 50, 1, 1, 1.0]) if srif else np.diag([10, 10, 10, 1, 1, 1.0])
        flags = np.full(steps, 1, dtype=np.uint8) if srif else np.ascontiguousarray(scn.flags)
        t0 = time.perf_counter()
        for _ in range(reps):
            gko.run_nl_batch(gko.SRIF if srif else gko.HYBRID, np.zeros(6), P0, np.diag([1e-6, 1e-6]), flags, Phi, Ht, real, comp,
                             threads=threads)
        dt = time.perf_counter() - t0
    return nf * steps * reps / dt, dt


FILTER_WORKLOADS = {
    "hybrid6": "hybrid6: 6-state hybrid CKF->EKF, range + range-rate, per-filter Phi/Htilde streams (BASELINE configs[3])",
    "hybrid6_strict": "hybrid6: 6-state hybrid CKF->EKF, range + range-rate, per-filter Phi/Htilde streams (BASELINE configs[3])",
    "hybrid6_fma": "hybrid6: 6-state hybrid CKF->EKF, range + range-rate, per-filter Phi/Htilde streams (BASELINE configs[3])",
    "srif6": "srif6: 6-state SRIF, range + range-rate, per-filter Phi/Htilde streams (BASELINE configs[3])",
    "vanilla32": "vanilla32: synthetic 32-state vanilla KF, m = 8 (BASELINE configs[4])",
    "vanilla64": "vanilla64: synthetic 64-state vanilla KF, m = 8 (the n = 64 shape of BASELINE configs[4])",
}


def filter_sample_size(workload, cores, target_s):
    steps = 200
    nf0 = cores * 8
    rate, _ = oracle_filter_rate(workload, nf0, steps, cores)  # calibration
    per_filter_bytes = steps * (64 if workload.startswith("vanilla") else 416)
    nf = int(min(rate * target_s / steps, 4e9 / per_filter_bytes))
    nf = max(cores, nf // cores * cores)
    reps = int(min(40, max(1, round(rate * target_s / (nf * steps)))))  # the streams are bounded by memory: repeat them
    return nf, steps, reps


def oracle_fma_spread(nf=256, steps=200):
    """How far the REFERENCE'S OWN formulas move on the bench's statOD streams when a*b+c is merely contracted: the CPU
    oracle built with -ffp-contract=fast against the same oracle unfused, same streams, same metric as the GPU arm's
    production_vs_strict (per filter: max |fma - unfused| over the final array / max |unfused|)."""
    from oracle import gko
    from gokalman_b200 import od  # host-side numpy tables only
    scn = od.Scenario(steps, 10.0, od.leo_truth0(), always_track=True, theta0=2.5)
    orbit0 = od.perturbed_orbits(od.leo_truth0(), nf, sigma_r=1.0, sigma_v=1e-3, seed=1234)
    Phi, Ht, real, comp, _ = gko.od_synth(scn.mu, scn.j2, scn.re, scn.dt, orbit0, scn.station, scn.truth_obs, 1e-3, 1e-3, 1234)
    P0, R, flags = np.diag([10, 10, 10, 1, 1, 1.0]), np.diag([1e-6, 1e-6]), np.ascontiguousarray(scn.flags)
    xr, Pr = gko.run_nl_batch(gko.HYBRID, np.zeros(6), P0, R, flags, Phi, Ht, real, comp, threads=os.cpu_count() or 1)
    xf, Pf = gko.run_nl_batch(gko.HYBRID, np.zeros(6), P0, R, flags, Phi, Ht, real, comp, threads=os.cpu_count() or 1, fma=True)

    def scaled(a, b):
        e = np.abs(a - b).max(axis=0) / np.abs(b).max(axis=0)
        return {"median": float(np.median(e)), "p99": float(np.quantile(e, 0.99)), "max": float(e.max())}
    cond = [float(np.linalg.cond(Pr[:, j].reshape(6, 6))) for j in range(min(nf, 32))]
    return {"filters": nf, "epochs": steps, "state": scaled(xf, xr), "covariance": scaled(Pf, Pr),
            "cond_final_covariance_median": float(np.median(cond)),
            "what": "CPU oracle with -ffp-contract=fast vs the same oracle unfused (the reference's formulas, both)"}


def cpu_baseline_filters(workload, target_s=10.0):
    cores = os.cpu_count() or 1
    nf, steps, reps = filter_sample_size(workload, cores, target_s)
    rate, dt = oracle_filter_rate(workload, nf, steps, cores, reps)
    extra = {}
    if workload.startswith("hybrid6"):
        try:
            extra["fma_spread"] = oracle_fma_spread()
        except Exception as e:
            extra["fma_spread"] = {"error": "%s: %s" % (type(e).__name__, e)}
    return {**extra, "value": rate, "unit": "filter-updates/s", "cores": cores, "kind": "port", "cpu_model": cpu_model(),
            "omp_num_threads": cores,
            "sample": "%d filters x %d epochs of %s, %d passes (%.1f s of filter work), C oracle restatement with OpenMP over "
                      "filters; the Go/gonum reference cannot be built here (no Go toolchain)" % (nf, steps, workload, reps, dt)}


def run_reference_filters(args, rank, world):
    if rank != 0:
        return None
    cores = os.cpu_count() or 1
    wl = args.workload
    nf, steps, reps = filter_sample_size(wl, cores, 4.0)
    times = []
    for i in range(args.warmup + args.steps):
        _, dt = oracle_filter_rate(wl, nf, steps, cores, reps)
        if i >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = nf * steps * reps * args.steps / total
    sample = ("%d filters x %d epochs x %d passes per step (bounded sample of the 10^5-filter workload), OpenMP x %d"
              % (nf, steps, reps, cores))
    if wl in ("hybrid6", "hybrid6_strict", "hybrid6_fma", "srif6"):  # the same config object as our arm (the sample is in cpu_baseline)
        from bench_hybrid import nl_config
        config = nl_config(wl, 100000 if args.trials == 1000000 else args.trials, args.filter_steps)
    else:
        n = int(wl[7:])
        config = {"workload": FILTER_WORKLOADS[wl], "filters_per_gpu": 26640 if wl == "vanilla64" else 100000,
                  "epochs": 100 if wl == "vanilla64" else 200, "n": n, "m": 8}
    return {
        "impl": "reference", "metric": "filter-updates/sec (batch x steps, FP64)", "value": value, "unit": "filter-updates/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config,
        "cpu_baseline": {"value": value, "unit": "filter-updates/s", "cores": cores, "kind": "port", "sample": sample,
                         "sample_filters": nf, "sample_epochs": steps, "sample_passes": reps},
        "e2e": {"value": value, "unit": "filter-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "CPU oracle port (C, OpenMP); the reference is Go + un-vendored gonum and cannot be built in this image",
    }


def run_reference(args, rank, world):
    if rank != 0:
        return None
    if args.workload in FILTER_WORKLOADS:
        return run_reference_filters(args, rank, world)
    cores = os.cpu_count() or 1
    steps = args.filter_steps
    wl = args.workload if args.workload in MC_WORKLOADS else "mc_jerk3"
    f = mc_model(wl)
    rate, _, _ = oracle_mc_rate(cores * 4, steps, cores, wl)
    trials = max(cores, int(rate * 6.0 / steps / cores) * cores)  # ~6 s of CPU per step
    times = []
    for i in range(args.warmup + args.steps):
        r, dt, _ = oracle_mc_rate(trials, steps, cores, wl)
        if i >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = trials * steps * args.steps / total
    sample = "%d trials x %d steps per step (bounded sample of the 10^6-trial workload), OpenMP x %d" % (trials, steps, cores)
    return {
        "impl": "reference", "metric": "filter-updates/sec (batch x steps, FP64)", "value": value, "unit": "filter-updates/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": MC_WORKLOADS[wl]["label"], "filter_steps": steps, "n": int(np.asarray(f["F"]).shape[0]),
                   "m": int(np.asarray(f["H"]).shape[0]), "c": 1},
        "cpu_baseline": {"value": value, "unit": "filter-updates/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "filter-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "CPU oracle port (C, OpenMP); the reference is Go + un-vendored gonum and cannot be built in this image",
    }


def run_subrecords(args, rank, world, local, shared):
    """Short records of the other BASELINE configs, appended to the default (hybrid6) line: same process, same
    clocks discipline, device-resident figures only."""
    from bench_hybrid import run_ours_hybrid
    from bench_tile import run_ours_tile
    sub = {}

    def add(name, fn):
        try:
            sub[name] = fn()
        except Exception as e:  # a sub-record must never take the headline down with it
            sub[name] = {"error": "%s: %s" % (type(e).__name__, e)}
    add("hybrid6_fma", lambda: run_ours_hybrid(args, rank, world, local, workload="hybrid6_fma", sub=True, shared=shared))
    add("srif6", lambda: run_ours_hybrid(args, rank, world, local, workload="srif6", sub=True, shared=shared))
    shared.clear()  # frees the 41.6 GB of streams
    import torch
    torch.cuda.empty_cache()
    mc_args = argparse.Namespace(**vars(args))
    mc_args.trials, mc_args.filter_steps = 1000000, 1000
    add("mc_jerk3", lambda: run_ours_mc(mc_args, rank, world, local, workload="mc_jerk3", sub=True))
    if world > 1:
        add("mc_jerk3_strong", lambda: run_ours_mc(mc_args, rank, world, local, workload="mc_jerk3", sub=True, strong=True))
    tile_args = argparse.Namespace(**vars(args))
    tile_args.workload, tile_args.trials, tile_args.filter_steps = "vanilla32", 1000000, 1000  # = that workload's defaults
    add("vanilla32", lambda: run_ours_tile(tile_args, rank, world, local, sub=True))
    return sub


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="hybrid6", choices=["hybrid6", "hybrid6_strict", "hybrid6_fma", "srif6", "mc_jerk3", "mc_robot_info",
                                                              "mc_robot_sqrt", "vanilla32", "vanilla64"])
    ap.add_argument("--trials", type=int, default=1000000, help="Monte Carlo trials per GPU (MC workloads) / filters per GPU "
                    "(filter workloads: the default means 10^5, or 26 640 for vanilla64)")
    ap.add_argument("--filter-steps", type=int, default=1000, help="filter steps / epochs per trial (vanilla32 / vanilla64: the "
                    "default means 200 / 100)")
    ap.add_argument("--strong", action="store_true", help="MC workloads: --trials is the TOTAL over all GPUs (strong scaling)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sub", action="store_true", help="default workload only: skip the sub-records of the other configs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank, world, local = dist_env()
    if args.impl == "reference":
        line = run_reference(args, rank, world)
        if line is not None:
            print(json.dumps(line), flush=True)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    shared = {}
    if args.workload in ("hybrid6", "hybrid6_strict", "hybrid6_fma", "srif6"):
        from bench_hybrid import run_ours_hybrid
        line = run_ours_hybrid(args, rank, world, local, shared=shared)
    elif args.workload in ("vanilla32", "vanilla64"):
        from bench_tile import run_ours_tile
        line = run_ours_tile(args, rank, world, local)
    else:
        line = run_ours_mc(args, rank, world, local, strong=args.strong)
    sub = None
    if args.workload == "hybrid6" and not args.no_sub:
        sub = run_subrecords(args, rank, world, local, shared)
    apply_fp64_peak_fixups()  # the FP64 peak micro-benchmark runs here, after every timed region
    if line is not None:
        if sub is not None:
            line["sub"] = sub
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = (cpu_baseline(workload=args.workload) if args.workload in MC_WORKLOADS
                                    else cpu_baseline_filters(args.workload))
        elif "cpu_baseline" not in line:
            line["cpu_baseline"] = None
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
