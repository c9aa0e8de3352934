#!/usr/bin/env python
"""Benchmark of the lnpost hot path (BASELINE.json: "lnpost evals/sec (batched walkers)").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1]): StarModel lnpost batch 1e6 — Sun-like single star (Teff, logg, feh + V, J, H, K
+ parallax) on the full MIST-shaped evolution-track grid (15 x 196 x 1710 nodes) and BC grid (70 x 26 x 18 x 13).
A "step" is one pass of the fused lnpost kernel over one batch of 1e6 parameter vectors; the timed steps rotate over
8 distinct batches (8 x 40 MB = 320 MB of inputs, larger than the 126 MB L2).  With N > 1 (torchrun, one rank per
GPU) every rank owns its own 1e6-row shard per step (weak scaling, no data-path collective); the all-gather the
sampler's acceptance step needs is timed separately and reported under "allgather".

Keys beyond the base contract: "roofline" and "cpu_baseline" (see DESIGN.md §5), "alt" (the same kernel on the
prior-like and grid-scattered batches).  `--impl reference` times the CPU path (the C port of the reference's numba
kernels under oracle/, all host threads) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

BATCH = 1_000_000
N_BATCHES = 8
BANDS = ("V", "J", "H", "K")
PACK_COLUMNS = ("Teff", "logg", "feh", "Mbol", "age", "dt_deep", "nu_max", "delta_nu")
B_ALG = 944.0            # algorithmic bytes per single-star, 4-band lnpost evaluation
METRIC = "lnpost evals/sec (batched walkers)"
UNIT = "evals/s"


def build_workload(ctx=None, small=False):
    """Grids, interpolator, star model and the truth point.  Needs the GPU only through ``ctx`` (None: host objects)."""
    import isochrones_b200 as ib
    from isochrones_b200 import synthetic as syn

    if small:
        trk = syn.make_track_grid(n_feh=6, n_mass=24, n_eep=171, columns=PACK_COLUMNS)
        bc = syn.make_bc_grid(bands=BANDS, n_teff=24, n_logg=10, n_feh=8, n_av=7)
    else:
        trk = syn.make_track_grid(columns=PACK_COLUMNS)
        bc = syn.make_bc_grid(bands=BANDS)
    n_eep = len(trk["axes"][2])
    ic = ib.ichrone_from_arrays("track", trk, bc, ctx=ctx)
    truth = syn.default_truth("track", n_eep=n_eep)
    return trk, bc, ic, truth, n_eep


def truth_mags(trk, bc, truth):
    """Observed magnitudes = model magnitudes at the truth (computed with the oracle so that both arms share them)."""
    from oracle import oracle

    mg = oracle.Grid(trk["grid"], trk["axes"])
    bg = oracle.Grid(bc["grid"], bc["axes"])
    ci = {c: i for i, c in enumerate(trk["columns"])}
    _, _, _, mags = oracle.interp_mags(np.asarray(truth).reshape(5, 1), [2, 0, 1, 3, 4], mg, ci["Teff"], ci["logg"],
                                       ci["feh"], ci["Mbol"], bg, list(range(len(BANDS))))
    return [float(np.round(m, 3)) for m in mags[0]], mg, bg


def make_model(ic, mags):
    import isochrones_b200 as ib

    obs = {b: (m, 0.02) for b, m in zip(BANDS, mags)}
    return ib.BasicStarModel(ic, Teff=(5772.0, 80.0), logg=(4.44, 0.1), feh=(0.0, 0.1), parallax=(10.0, 0.1), **obs)


def scattered_batch(n, seed):
    """Rows spread uniformly over the part of the track grid where every EEP is populated and the BC lookup is in
    range: all rows do the full work and the gathers scatter over the grid (HBM/L2-bound case)."""
    rng = np.random.RandomState(seed)
    p = np.empty((n, 5))
    p[:, 0] = np.exp(rng.uniform(np.log(0.7), np.log(5.9), n))     # mass
    p[:, 1] = rng.uniform(1.0, 1700.0, n)                           # eep
    p[:, 2] = rng.uniform(-2.0, 0.45, n)                            # feh
    p[:, 3] = rng.uniform(20.0, 199.0, n)                           # distance
    p[:, 4] = rng.uniform(0.0, 0.99, n)                             # AV
    return p


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons sampled DURING the timed regions (NVML polled every ~2 ms; nvidia-smi fallback)."""

    def __init__(self, device):
        super().__init__(daemon=True)
        self.device = device
        self.sm, self.power, self.reasons = [], [], set()
        self.sm_max = None
        self.stop_flag = threading.Event()
        self.source = "nvml"

    def _run_nvml(self):
        import pynvml as nv

        nv.nvmlInit()
        # LOCAL_RANK indexes the visible devices; map through CUDA_VISIBLE_DEVICES when it is a plain index list
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = self.device
        if vis:
            try:
                idx = int(vis.split(",")[self.device])
            except (ValueError, IndexError):
                idx = self.device
        h = nv.nvmlDeviceGetHandleByIndex(idx)
        self.sm_max = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        names = {nv.nvmlClocksEventReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksEventReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksEventReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksEventReasonSwPowerCap: "sw_power_cap",
                 nv.nvmlClocksEventReasonHwPowerBrakeSlowdown: "hw_power_brake_slowdown"}
        while not self.stop_flag.is_set():
            self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
            try:
                self.power.append(nv.nvmlDeviceGetPowerUsage(h) / 1000.0)
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self.stop_flag.wait(0.002)

    def _run_smi(self):
        self.source = "nvidia-smi"
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(",")]
                if len(parts) >= 7:
                    self.sm.append(float(parts[0]))
                    self.sm_max = float(parts[1])
                    self.power.append(float(parts[2]))
                    for i, n in enumerate(names):
                        if parts[3 + i].lower().startswith("active"):
                            self.reasons.add(n)
            except Exception:
                pass
            self.stop_flag.wait(0.05)

    def run(self):
        try:
            self._run_nvml()
        except Exception:
            self._run_smi()

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": [], "samples": 0, "source": self.source}
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2], "sm_min_mhz": sm[0], "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons),
                "samples": len(sm), "power_w_max": max(self.power) if self.power else None, "source": self.source,
                "window": "device-timed loop + end-to-end loop"}


def cpu_baseline(trk, bc, ic_host_model, mg, bg, seed, budget_s=12.0):
    """The C port of the reference's CPU path (oracle/) on the host cores: bounded sample of the same workload."""
    from isochrones_b200 import synthetic as syn
    from oracle import oracle

    om = oracle.StarModel(ic_host_model, model_grid=mg, bc_grid=bg)
    threads = oracle.max_threads()
    truth = syn.default_truth("track", n_eep=len(trk["axes"][2]))
    probe = syn.posterior_like_batch("track", 200_000, truth, seed=seed)
    t0 = time.perf_counter()
    om.lnpost_batch(probe, n_threads=threads)
    rate = len(probe) / (time.perf_counter() - t0)
    n = int(min(max(rate * budget_s, 200_000), 40_000_000))
    sample = syn.posterior_like_batch("track", n, truth, seed=seed + 1)
    t0 = time.perf_counter()
    om.lnpost_batch(sample, n_threads=threads)
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "%d posterior-like rows of the bench workload, oracle/iso_oracle.c (C port of the reference's numba "
                      "kernels + priors), OpenMP over %d threads, %.1f s" % (n, threads, dt)}


def run_reference(args, rank, world):
    """`--impl reference`: the reference's CPU implementation of the path.  The reference is pure Python + numba and
    cannot travel to the GPU box (no /root/reference there), so its C port under oracle/ is timed: kind = "port"."""
    if rank != 0:
        return
    from isochrones_b200 import synthetic as syn
    from oracle import oracle

    trk, bc, ic, truth, n_eep = build_workload(ctx=None)
    mags, mg, bg = truth_mags(trk, bc, truth)
    mod = make_model(ic, mags)
    om = oracle.StarModel(mod, model_grid=mg, bc_grid=bg)
    threads = oracle.max_threads()
    probe = syn.posterior_like_batch("track", 200_000, truth, seed=100)
    t0 = time.perf_counter()
    om.lnpost_batch(probe, n_threads=threads)
    rate = len(probe) / (time.perf_counter() - t0)
    total_budget = 60.0
    n = int(min(BATCH, max(50_000, rate * total_budget / max(args.steps + args.warmup, 1))))
    batches = [syn.posterior_like_batch("track", n, truth, seed=2 + s) for s in range(min(N_BATCHES, args.steps + args.warmup))]
    for s in range(args.warmup):
        om.lnpost_batch(batches[s % len(batches)], n_threads=threads)
    t0 = time.perf_counter()
    for s in range(args.steps):
        om.lnpost_batch(batches[s % len(batches)], n_threads=threads)
    dt = time.perf_counter() - t0
    value = n * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "StarModel lnpost batch 1e6 (configs[1]): single Sun-like star, Teff/logg/feh + VJHK + "
                               "parallax, MIST-shaped track grid 15x196x1710, BC 70x26x18x13",
                   "batch_rows": BATCH, "distribution": "posterior-like", "sample_rows_per_step": n},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "%d rows per step (bounded sample of the 1e6-row batch), oracle/iso_oracle.c with "
                                   "OpenMP over %d threads" % (n, threads)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def timed_device_loop(ctx, compiled, d_batches, d_out, steps, warmup, barrier=None):
    for s in range(warmup):
        compiled.lnpost_device(d_batches[s % len(d_batches)], BATCH, d_out)
    ctx.sync()
    if barrier:
        barrier()
    l0 = ctx.launch_count()
    ctx.timer_start()
    for s in range(steps):
        compiled.lnpost_device(d_batches[s % len(d_batches)], BATCH, d_out)
    ms = ctx.timer_stop()
    ctx.sync()
    if barrier:
        barrier()
    return ms, ctx.launch_count() - l0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--small", action="store_true", help="small grids (debugging only; not a valid bench)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from isochrones_b200 import _lib, parallel, synthetic as syn

    ctx = _lib.default_context(local_rank)
    trk, bc, ic, truth, n_eep = build_workload(ctx=ctx, small=args.small)
    mags, mg, bg = truth_mags(trk, bc, truth)
    mod = make_model(ic, mags)
    compiled = mod.compiled
    bounds = [mod.bounds(p) for p in mod.param_names]

    def barrier():
        if dist is not None:
            import torch

            t = torch.zeros(1, device="cuda")
            dist.all_reduce(t)
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch

        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident batches: this rank's shard of every step --------------------------------------------
    def stage(batches):
        ptrs = []
        for b in batches:
            d = ctx.dev_alloc(b.nbytes)
            ctx.h2d(d, b)
            ptrs.append(d)
        return ptrs

    post = [syn.posterior_like_batch("track", BATCH, truth, n_eep=n_eep, seed=2 + 1000 * rank + s) for s in range(N_BATCHES)]
    d_post = stage(post)
    d_out = ctx.dev_alloc(BATCH * 8)

    clocks = ClockSampler(local_rank)
    clocks.start()
    ms, launches = timed_device_loop(ctx, compiled, d_post, d_out, args.steps, args.warmup, barrier)
    ms = max_over_ranks(ms)
    value = world * BATCH * args.steps / (ms * 1e-3)
    kernel_ms = ms / args.steps

    # ---- end to end through the public API: pinned host buffers, H2D + kernel + D2H inside the timed region ----
    h_in = [ctx.pinned_empty((BATCH, 5)) for _ in range(2)]
    h_out = ctx.pinned_empty((BATCH,))
    for i in range(2):
        h_in[i][:] = post[i]
    for s in range(max(args.warmup, 3)):
        mod.lnpost_batch(h_in[s % 2], out=h_out)
    barrier()
    t0 = time.perf_counter()
    for s in range(args.steps):
        mod.lnpost_batch(h_in[s % 2], out=h_out)
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e_value = world * BATCH * args.steps / e2e_s
    clock_summary = clocks.summary()
    finite_frac = float(np.isfinite(h_out).mean())

    # ---- the all-gather the sampler's acceptance step needs (timed apart from the sharded evaluation) ---------
    allgather = None
    if world > 1:
        comm = parallel.NcclGather(ctx, rank, world, exchange=parallel.torch_exchange(dist))
        d_all = ctx.dev_alloc(world * BATCH * 8)
        for s in range(3):
            compiled.lnpost_device(d_post[s % N_BATCHES], BATCH, d_out)
            comm.allgather(d_out, BATCH, d_all)
        ctx.sync()
        barrier()
        ctx.timer_start()
        for s in range(args.steps):
            compiled.lnpost_device(d_post[s % N_BATCHES], BATCH, d_out)
            comm.allgather(d_out, BATCH, d_all)
        ms_g = max_over_ranks(ctx.timer_stop())
        barrier()
        allgather = {"ms_per_step_with_gather": ms_g / args.steps, "value_with_gather": world * BATCH * args.steps / (ms_g * 1e-3),
                     "bytes_per_rank_per_step": BATCH * 8, "collective": "ncclAllGather f64 on the compute stream"}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- rank 0 only: alternative distributions, roofline, CPU baseline ---------------------------------------
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md 6.65 TB/s)"

    alt = {}
    for name, gen in (("prior_like", lambda s: syn.prior_like_batch("track", BATCH, bounds, seed=3 + s)),
                      ("scattered_valid", lambda s: scattered_batch(BATCH, seed=50 + s))):
        batches = [gen(s) for s in range(N_BATCHES)]
        d_b = stage(batches)
        ms_a, _ = timed_device_loop(ctx, compiled, d_b, d_out, args.steps, args.warmup)
        out = np.empty(BATCH)
        ctx.d2h(out, d_out)
        k_ms = ms_a / args.steps
        alt[name] = {"value": BATCH / (k_ms * 1e-3), "unit": UNIT, "ms_per_step": k_ms,
                     "finite_frac": float(np.isfinite(out).mean()),
                     "roofline_frac": B_ALG * BATCH / (k_ms * 1e-3) / 1e9 / peak}
        for d in d_b:
            ctx.dev_free(d)

    achieved = B_ALG * BATCH / (kernel_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            with open(tpath) as f:
                traffic = json.load(f).get("lnpost_posterior_like_dram_bytes_per_launch")
        except Exception:
            traffic = None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": kernel_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "StarModel lnpost batch 1e6 (configs[1]): single Sun-like star, Teff/logg/feh + VJHK + "
                               "parallax, MIST-shaped track grid 15x196x1710 (8-col pack, %d MB in HBM), BC 70x26x18x13"
                               % (trk["grid"].nbytes // 2 ** 20),
                   "batch_rows": BATCH, "rows_per_gpu_per_step": BATCH, "distribution": "posterior-like (sigma: mass .05, "
                   "eep 15, feh .1, d 2 pc, AV .05)", "l2": "inputs larger than L2: steps rotate over %d distinct 1e6-row "
                   "batches (%d MB)" % (N_BATCHES, N_BATCHES * BATCH * 40 // 2 ** 20), "finite_frac": finite_frac,
                   "sharding": "rows sharded across ranks, no data-path collective" if world > 1 else "single GPU"},
        "clocks": clock_summary,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": BATCH * 40, "d2h_bytes_per_step": BATCH * 8,
                "api": "BasicStarModel.lnpost_batch(pinned host array) -> iso_lnpost_batch (C ABI), chunked "
                       "H2D/kernel/D2H pipeline on two streams"},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "kernel": "iso_lnpost_kernel<1,false>", "algorithmic_bytes_per_eval": B_ALG,
                     "peak_source": peak_src, "kernel_ms": kernel_ms},
        "alt": alt,
    }
    if allgather:
        line["allgather"] = allgather
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(trk, bc, mod, mg, bg, seed=900)
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
