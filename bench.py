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


def host_threads():
    """Host cores this process may use (the CPU arm uses all of them through OpenMP's num_threads clause)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


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


def truth_mags(ic, truth):
    """Observed magnitudes of the b200 arm = model magnitudes at the truth, rounded to 1 mmag (the product's own
    ``interp_mag``; the reference arm derives the same numbers from the oracle, see ``oracle_truth_mags``)."""
    _, _, _, mags = ic.interp_mag(list(truth), list(BANDS))
    return [float(np.round(m, 3)) for m in mags]


def oracle_grids(trk, bc):
    """Oracle views of the two grids (CPU arm / cpu_baseline leg only)."""
    from oracle import oracle

    return oracle.Grid(trk["grid"], trk["axes"]), oracle.Grid(bc["grid"], bc["axes"])


def oracle_truth_mags(trk, bc, truth):
    """The same observed magnitudes computed on the CPU (reference arm: no GPU is touched there)."""
    from oracle import oracle

    mg, bg = oracle_grids(trk, bc)
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


_POLLER = r"""
import sys, time
idx = int(sys.argv[1])
names = []
try:
    import pynvml as nv
    nv.nvmlInit()
    h = nv.nvmlDeviceGetHandleByIndex(idx)
    smax = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
    bits = [(nv.nvmlClocksEventReasonHwSlowdown, 'hw_slowdown'), (nv.nvmlClocksEventReasonHwThermalSlowdown, 'hw_thermal_slowdown'),
            (nv.nvmlClocksEventReasonSwThermalSlowdown, 'sw_thermal_slowdown'), (nv.nvmlClocksEventReasonSwPowerCap, 'sw_power_cap'),
            (nv.nvmlClocksEventReasonHwPowerBrakeSlowdown, 'hw_power_brake_slowdown')]
    print('ready', smax, flush=True)
    while True:
        sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
        try:
            pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
            mask = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
        except Exception:
            pw, mask = -1.0, 0
        r = '|'.join(n for b, n in bits if mask & b)
        print('%.6f,%d,%.1f,%s' % (time.time(), sm, pw, r), flush=True)
        time.sleep(0.001)
except Exception as e:
    print('error', repr(e), flush=True)
"""


def grid_wide_batch(n, trk, seed):
    """Rows spread over the WHOLE populated track grid (every [Fe/H], masses 0.1-100, EEPs up to the end of the
    shorter of the bracketing tracks): the gathers touch ~200 MB of model-grid nodes, more than the 126 MB L2 — the
    HBM-bound extreme of the kernel.  Many rows fall outside the age prior (old low-mass stars) and stop after the
    model-grid gather."""
    from isochrones_b200 import synthetic as syn

    rng = np.random.RandomState(seed)
    fehs, masses, eeps = trk["axes"]
    last = np.array([[syn.max_eep_table(m, f) for m in masses] for f in fehs], dtype=float)
    last = np.minimum(last, len(eeps))
    cell_last = np.minimum(np.minimum(last[:-1, :-1], last[1:, :-1]), np.minimum(last[:-1, 1:], last[1:, 1:]))
    m_hi = np.searchsorted(masses, 100.0) - 1
    a = rng.randint(0, len(fehs) - 1, n)
    b = rng.randint(0, m_hi, n)
    p = np.empty((n, 5))
    p[:, 0] = masses[b] + (masses[b + 1] - masses[b]) * rng.random_sample(n)
    p[:, 1] = 1.0 + (cell_last[a, b] - 2.0) * rng.random_sample(n)
    p[:, 2] = fehs[a] + (fehs[a + 1] - fehs[a]) * rng.random_sample(n)
    p[:, 3] = rng.uniform(20.0, 199.0, n)
    p[:, 4] = rng.uniform(0.0, 0.99, n)
    return p


class ClockSampler(object):
    """SM clock / throttle reasons sampled DURING the timed regions by a separate NVML polling process (~1 kHz), so
    that the sampling neither holds this process's GIL nor delays its kernel launches.  `mark()` brackets the timed
    regions; only samples inside a bracket are summarised."""

    def __init__(self, device):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = device
        if vis:
            try:
                idx = int(vis.split(",")[device])
            except (ValueError, IndexError):
                idx = device
        self.windows = []
        self.sm_max = None
        self.proc = None
        try:
            self.proc = subprocess.Popen([sys.executable, "-c", _POLLER, str(idx)], stdout=subprocess.PIPE, text=True)
            first = self.proc.stdout.readline().split()
            if first and first[0] == "ready":
                self.sm_max = float(first[1])
            else:
                self.proc.kill()
                self.proc = None
        except Exception:
            self.proc = None

    def start(self):
        pass

    def mark(self, t0, t1):
        self.windows.append((t0, t1))

    def summary(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": "unavailable"}
        time.sleep(0.005)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, power, reasons = [], [], set()
        for ln in out.splitlines():
            parts = ln.split(",")
            if len(parts) != 4:
                continue
            try:
                t = float(parts[0])
            except ValueError:
                continue
            if not any(a <= t <= b for a, b in self.windows):
                continue
            sm.append(float(parts[1]))
            power.append(float(parts[2]))
            reasons.update(x for x in parts[3].split("|") if x)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": [], "samples": 0, "source": "nvml"}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_min_mhz": sm[0], "sm_max_mhz": self.sm_max, "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power), "source": "nvml polled at ~1 kHz by a side process",
                "window": "samples inside the device-timed loop and the end-to-end loop only"}


def cpu_baseline(trk, bc, ic_host_model, seed, budget_s=12.0):
    """The C port of the reference's CPU path (oracle/) on the host cores: bounded sample of the same workload."""
    from isochrones_b200 import synthetic as syn
    from oracle import oracle

    mg, bg = oracle_grids(trk, bc)
    om = oracle.StarModel(ic_host_model, model_grid=mg, bc_grid=bg)
    threads = host_threads()   # not omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1
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
    mags, mg, bg = oracle_truth_mags(trk, bc, truth)
    mod = make_model(ic, mags)
    om = oracle.StarModel(mod, model_grid=mg, bc_grid=bg)
    threads = host_threads()   # not omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1
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


def interp_workloads(ctx, ic, truth, n_eep, args, mod=None):
    """BASELINE.json configs[0] at bench size: the standalone interpolation entry points (DFInterpolator /
    interp_value, interp_mag) on 1e6 points through the host-pointer C ABI (pinned buffers; H2D + kernel + D2H), and
    the latency of the reference-shaped small calls (scalar lnpost, one 128-row half-step)."""
    from isochrones_b200 import synthetic as syn

    out = {}
    pts = syn.posterior_like_batch("track", BATCH, truth, n_eep=n_eep, seed=321)
    xs = [ctx.pinned_empty((BATCH,)) for _ in range(5)]
    for j in range(5):
        xs[j][:] = pts[:, j]
    props = ["Teff", "logg", "nu_max"]
    vout = ctx.pinned_empty((BATCH, len(props)))
    for _ in range(3):
        ic.interp_value(xs[:3], props, out=vout)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        v = ic.interp_value(xs[:3], props, out=vout)
    dt = (time.perf_counter() - t0) / args.steps
    out["interp_value_3_props"] = {
        "value": BATCH / dt, "unit": "points/s", "ms_per_step": dt * 1e3, "finite_frac": float(np.isfinite(v).mean()),
        "config": "configs[0] shape at 1e6 points: ModelGridInterpolator.interp_value((mass, eep, feh), 3 props, out=pinned) "
                  "-> iso_interp_values, host arrays in / out (48 B per point over PCIe)"}
    # the reference's scalar interface, as emcee / MultiNest call it: one lnpost(p) per Python call, and a
    # 128-row batch (one emcee half-step of 256 walkers) through lnpost_batch on pageable arrays
    if mod is not None:
        p1 = list(pts[0])
        for _ in range(200):
            mod.lnpost(p1)
        reps = 2000
        t0 = time.perf_counter()
        for _ in range(reps):
            mod.lnpost(p1)
        dt = (time.perf_counter() - t0) / reps
        out["scalar_lnpost_call"] = {
            "value": 1.0 / dt, "unit": UNIT, "us_per_call": dt * 1e6,
            "config": "BasicStarModel.lnpost(p): one row per call through the C ABI (H2D, one-row launch, D2H, sync); "
                      "the reference's own scalar call takes 68-93 us in the build container, 369 us published"}
        half = np.ascontiguousarray(pts[:128])
        for _ in range(200):
            mod.lnpost_batch(half)
        t0 = time.perf_counter()
        for _ in range(reps):
            mod.lnpost_batch(half)
        dt = (time.perf_counter() - t0) / reps
        out["emcee_half_step_128_rows_from_host"] = {
            "value": 128.0 / dt, "unit": UNIT, "us_per_call": dt * 1e6,
            "config": "BasicStarModel.lnpost_batch on a pageable [128, 5] array: what a host-driven emcee (vectorize=True) "
                      "pays per half-step of a 256-walker ensemble"}
    # the two interpolation kernels alone (device-resident inputs and outputs, CUDA events)
    d_x = []
    for j in range(5):
        d = ctx.dev_alloc(BATCH * 8)
        ctx.h2d(d, np.ascontiguousarray(pts[:, j]))
        d_x.append(d)
    grid = ic.model_grid.interp
    icols = [grid.column_index[c] for c in props]
    d_v = ctx.dev_alloc(BATCH * 8 * len(props))
    d_m = [ctx.dev_alloc(BATCH * 8) for _ in range(3)] + [ctx.dev_alloc(BATCH * 8 * len(BANDS))]
    for name, fn, b_alg in (
            ("interp_value_3_props_device", lambda: grid.device_grid.interp_values_device([d_x[2], d_x[0], d_x[1]], BATCH, icols, d_v),
             8.0 * (3 + 8 * 3 + 3)),
            ("interp_mag_4_bands_device", lambda: ic.interp_mag_device(d_x, BATCH, list(BANDS), *d_m),
             8.0 * (5 + 8 * 4 + 16 * len(BANDS) + 3 + len(BANDS)))):
        for _ in range(3):
            fn()
        ctx.sync()
        ctx.timer_start()
        for _ in range(args.steps):
            fn()
        k_ms = ctx.timer_stop() / args.steps
        out[name] = {"value": BATCH / (k_ms * 1e-3), "unit": "points/s", "ms_per_step": k_ms,
                     "algorithmic_bytes_per_point": b_alg, "achieved_gb_s": b_alg * BATCH / (k_ms * 1e-3) / 1e9,
                     "config": "kernel only: inputs and outputs in HBM (iso_%s)" % ("interp_values_device" if "value" in name else "interp_mags_device")}
    for d in d_x + d_m + [d_v]:
        ctx.dev_free(d)
    outs = (ctx.pinned_empty((BATCH,)), ctx.pinned_empty((BATCH,)), ctx.pinned_empty((BATCH,)),
            ctx.pinned_empty((BATCH, len(BANDS))))
    for _ in range(3):
        ic.interp_mag(xs, list(BANDS), out=outs)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r = ic.interp_mag(xs, list(BANDS), out=outs)
    dt = (time.perf_counter() - t0) / args.steps
    out["interp_mag_4_bands"] = {
        "value": BATCH / dt, "unit": "points/s", "ms_per_step": dt * 1e3, "finite_frac": float(np.isfinite(r[3]).mean()),
        "config": "ModelGridInterpolator.interp_mag(5 parameter arrays, VJHK, out=pinned arrays) -> iso_interp_mags, "
                  "host arrays in / out (96 B per point over PCIe + one host-side pack of the [5, N] block)"}
    return out


def extra_workloads(ctx, bc, args, peak):
    """BASELINE.json configs 3 and 5 on one GPU (informational "alt" entries; parity for them lives in tests/):
    binary-star lnpost batch on the isochrone grid, and the on-device ensemble sampler (256 walkers x 2000 steps as
    one chain, and 1184 such chains x 100 steps filling the GPU)."""
    import isochrones_b200 as ib
    from isochrones_b200 import synthetic as syn
    from isochrones_b200.sampler import DeviceEnsembleSampler

    out = {}
    iso = syn.make_iso_grid(columns=("Teff", "logg", "feh", "Mbol", "mass", "dm_deep", "nu_max", "delta_nu"))
    ic = ib.ichrone_from_arrays("iso", iso, bc, ctx=ctx)
    truth = syn.default_truth("iso", n_stars=2)
    _, _, _, mags = ic.interp_mag([truth[0]] + list(truth[2:]), list(BANDS))
    obs = {b: (float(np.round(m, 3)) - 0.35, 0.02) for b, m in zip(BANDS, mags)}
    binary = ib.BinaryStarModel(ic, Teff=(5772.0, 80.0), logg=(4.44, 0.1), feh=(0.0, 0.1), parallax=(10.0, 0.1), **obs)
    batches = []
    for s in range(4):
        p = syn.posterior_like_batch("iso", BATCH, truth, seed=70 + s)
        p[:, :2] = -np.sort(-p[:, :2], axis=1)
        batches.append(p)
    d_b = []
    for b in batches:
        d = ctx.dev_alloc(b.nbytes)
        ctx.h2d(d, b)
        d_b.append(d)
    d_out = ctx.dev_alloc(BATCH * 8)
    ms, _ = timed_device_loop(ctx, binary.compiled, d_b, d_out, args.steps, args.warmup)
    res = np.empty(BATCH)
    ctx.d2h(res, d_out)
    k_ms = ms / args.steps
    out["binary_iso_posterior_like"] = {
        "value": BATCH / (k_ms * 1e-3), "unit": UNIT, "ms_per_step": k_ms, "finite_frac": float(np.isfinite(res).mean()),
        "algorithmic_bytes_per_eval": 1848.0, "roofline_frac": 1848.0 * BATCH / (k_ms * 1e-3) / 1e9 / peak,
        "config": "configs[4] on one GPU: BinaryStarModel, iso grid 107x15x1710, 4 bands + parallax, 1e6 rows"}
    for d in d_b:
        ctx.dev_free(d)

    # single-star model on the iso grid drives the sampler runs (starfit's default grid)
    truth1 = syn.default_truth("iso", n_stars=1)
    _, _, _, mags1 = ic.interp_mag(list(truth1), list(BANDS))
    single = ib.SingleStarModel(ic, Teff=(5772.0, 80.0), logg=(4.44, 0.1), feh=(0.0, 0.1), parallax=(10.0, 0.1),
                                **{b: (float(np.round(m, 3)), 0.02) for b, m in zip(BANDS, mags1)})
    nw, nsteps = 256, 2000
    p0 = syn.posterior_like_batch("iso", nw, truth1, seed=4)
    smp = DeviceEnsembleSampler(single.compiled, nw, p0, seed=4)
    smp.run_mcmc(50, store=False)
    t0 = time.perf_counter()
    smp.run_mcmc(nsteps, thin=10)
    dt = time.perf_counter() - t0
    out["emcee_256x2000_one_chain"] = {
        "value": nw * nsteps / dt, "unit": UNIT, "seconds": dt, "acceptance_fraction": float(smp.acceptance_fraction[0]),
        "config": "configs[2]: stretch move a=2, 256 walkers x 2000 steps, one persistent CTA, one launch "
                  "(latency-bound: 4000 dependent half-steps)"}
    smp.close()
    n_chains, steps_c = 1184, 100
    p0c = np.stack([syn.posterior_like_batch("iso", nw, truth1, seed=1000 + c) for c in range(8)])
    p0c = np.ascontiguousarray(np.tile(p0c, (n_chains // 8, 1, 1)))
    smc = DeviceEnsembleSampler(single.compiled, nw, p0c, seed=5, n_chains=n_chains)
    smc.run_mcmc(10, store=False)
    t0 = time.perf_counter()
    smc.run_mcmc(steps_c, store=False)
    dt = time.perf_counter() - t0
    out["emcee_256_walkers_x_1184_chains"] = {
        "value": n_chains * nw * steps_c / dt, "unit": UNIT, "seconds": dt,
        "acceptance_fraction": float(np.mean(smc.acceptance_fraction)),
        "config": "1184 independent 256-walker ensembles x 100 steps in one launch (8 CTAs per SM)"}
    smc.close()

    # configs[3] in miniature on one GPU: a catalog of 10k independent star models, rows of many stars in one launch
    import pandas as pd

    from isochrones_b200.catalog import StarCatalog

    n_stars, rows_per_star = 10_000, 100
    rng = np.random.RandomState(8)
    t = np.tile(truth1, (n_stars, 1))
    t[:, 0] = rng.uniform(300.0, 900.0, n_stars)
    t[:, 1] = rng.uniform(9.0, 9.9, n_stars)
    t[:, 2] = rng.uniform(-0.5, 0.3, n_stars)
    t[:, 3] = rng.uniform(50.0, 400.0, n_stars)
    t[:, 4] = rng.uniform(0.0, 0.5, n_stars)
    _, _, _, mg = ic.interp_mag([t[:, j] for j in range(5)], list(BANDS))
    table = {"parallax": 1000.0 / t[:, 3], "parallax_unc": np.full(n_stars, 0.1)}
    for j, b in enumerate(BANDS):
        table[b + "_mag"] = mg[:, j]
        table[b + "_mag_unc"] = np.full(n_stars, 0.02)
    t0 = time.perf_counter()
    cat = StarCatalog(pd.DataFrame(table), props=["parallax"])
    compiled = cat.compile(ic)
    t_compile = time.perf_counter() - t0
    mor = np.repeat(np.arange(n_stars, dtype=np.int32), rows_per_star)
    pars = np.repeat(t, rows_per_star, axis=0)
    pars *= 1 + 0.002 * rng.standard_normal(pars.shape)
    n_rows = len(pars)
    d_p, d_m, d_o = ctx.dev_alloc(pars.nbytes), ctx.dev_alloc(mor.nbytes), ctx.dev_alloc(n_rows * 8)
    ctx.h2d(d_p, pars)
    ctx.h2d(d_m, mor)
    for _ in range(3):
        compiled.lnpost_device(d_p, n_rows, d_o, d_model_of_row=d_m)
    ctx.sync()
    ctx.timer_start()
    for _ in range(args.steps):
        compiled.lnpost_device(d_p, n_rows, d_o, d_model_of_row=d_m)
    k_ms = ctx.timer_stop() / args.steps
    res = np.empty(n_rows)
    ctx.d2h(res, d_o)
    out["catalog_10k_stars"] = {
        "value": n_rows / (k_ms * 1e-3), "unit": UNIT, "ms_per_step": k_ms, "finite_frac": float(np.isfinite(res).mean()),
        "compile_seconds": t_compile,
        "config": "configs[3] shape on one GPU: 10 000 star models (iso grid, 4 bands + parallax) in HBM, 1e6 rows "
                  "(100 per star) in one launch through model_of_row"}
    return out


def sharded_workloads(ctx, bc, args, rank, world, comm, barrier, max_over_ranks, dist):
    """BASELINE.json configs 4 and 5 on N GPUs (every rank runs this; rank 0 reports):
    configs[4]: BinaryStarModel lnpost batch of 1e6 rows IN TOTAL, rows sharded in contiguous blocks (RowSharder), the
    per-row results all-gathered so that every rank holds the full lnpost vector (what a sampler's acceptance step
    needs); configs[3]: a catalog of 10 000 independent star models, the STARS sharded across ranks (each rank stages
    only its own models), 100 rows per star per step, results all-gathered."""
    import pandas as pd

    import isochrones_b200 as ib
    from isochrones_b200 import parallel, synthetic as syn
    from isochrones_b200.catalog import StarCatalog

    out = {}
    iso = syn.make_iso_grid(columns=("Teff", "logg", "feh", "Mbol", "mass", "dm_deep", "nu_max", "delta_nu"))
    ic = ib.ichrone_from_arrays("iso", iso, bc, ctx=ctx)

    def timed(fn):
        for _ in range(3):
            fn()
        ctx.sync()
        barrier()
        ctx.timer_start()
        for _ in range(args.steps):
            fn()
        ms = max_over_ranks(ctx.timer_stop()) / args.steps
        barrier()
        return ms

    # ---- configs[4]: binary star, 1e6 rows in total ---------------------------------------------------------------
    truth = syn.default_truth("iso", n_stars=2)
    _, _, _, mags = ic.interp_mag([truth[0]] + list(truth[2:]), list(BANDS))
    obs = {b: (float(np.round(m, 3)) - 0.35, 0.02) for b, m in zip(BANDS, mags)}
    binary = ib.BinaryStarModel(ic, Teff=(5772.0, 80.0), logg=(4.44, 0.1), feh=(0.0, 0.1), parallax=(10.0, 0.1), **obs)
    rows = syn.posterior_like_batch("iso", BATCH, truth, seed=70)       # same seed on every rank: one global batch
    rows[:, :2] = -np.sort(-rows[:, :2], axis=1)
    sh = parallel.RowSharder(BATCH, world, rank)
    mine = np.ascontiguousarray(sh.local(rows))
    d_p = ctx.dev_alloc(max(mine.nbytes, 8))
    ctx.h2d(d_p, mine)
    d_o, d_all = ctx.dev_alloc(sh.pad * 8), ctx.dev_alloc(world * sh.pad * 8)
    n_mine = len(mine)
    ms_k = timed(lambda: binary.compiled.lnpost_device(d_p, n_mine, d_o))
    ms_g = timed(lambda: (binary.compiled.lnpost_device(d_p, n_mine, d_o), comm.allgather(d_o, sh.pad, d_all)))
    got = np.empty((world, sh.pad))
    ctx.d2h(got, d_all)
    full = sh.assemble(got)
    ms_f, same = None, None
    # the same step with the gather fused into the kernel (stores over NVLink peer mappings, iso_peer_*)
    peer, _ = make_peer_group(ctx, rank, world, sh.pad, dist, max_over_ranks) if world <= 8 else (None, "")
    if peer is not None:
        got_f = np.empty((world, sh.pad))
        ctx.d2h(got_f, peer.lnpost(binary.compiled, d_p, n_mine))
        same = bool(np.array_equal(sh.assemble(got_f), full, equal_nan=True))
        ms_f = timed(lambda: peer.lnpost(binary.compiled, d_p, n_mine))
        peer.close()
    out["binary_1e6_rows_sharded"] = {
        "value": BATCH / (ms_g * 1e-3), "unit": UNIT, "ms_per_step": ms_g, "ms_per_step_without_gather": ms_k,
        "ms_per_step_fused_peer_store": ms_f, "value_fused_peer_store": (BATCH / (ms_f * 1e-3)) if ms_f else None,
        "fused_identical_to_nccl": same,
        "rows_per_gpu": sh.pad, "finite_frac": float(np.isfinite(full).mean()), "scaling": "strong",
        "config": "configs[4]: BinaryStarModel, iso grid 107x15x1710, 4 bands + parallax, 1e6 rows in total sharded over "
                  "%d GPUs, ncclAllGather of lnpost (%d B per rank) inside the timed step" % (world, sh.pad * 8)}
    for d in (d_p, d_o, d_all):
        ctx.dev_free(d)

    # ---- configs[3]: catalog, stars sharded ---------------------------------------------------------------------
    n_stars, rows_per_star = 10_000, 100
    truth1 = syn.default_truth("iso", n_stars=1)
    rng = np.random.RandomState(8)
    t = np.tile(truth1, (n_stars, 1))
    t[:, 0] = rng.uniform(300.0, 900.0, n_stars)
    t[:, 1] = rng.uniform(9.0, 9.9, n_stars)
    t[:, 2] = rng.uniform(-0.5, 0.3, n_stars)
    t[:, 3] = rng.uniform(50.0, 400.0, n_stars)
    t[:, 4] = rng.uniform(0.0, 0.5, n_stars)
    ssh = parallel.RowSharder(n_stars, world, rank)
    a, b = ssh.bounds()
    tm = t[a:b]
    _, _, _, mg = ic.interp_mag([tm[:, j] for j in range(5)], list(BANDS))
    table = {"parallax": 1000.0 / tm[:, 3], "parallax_unc": np.full(b - a, 0.1)}
    for j, bn in enumerate(BANDS):
        table[bn + "_mag"] = mg[:, j]
        table[bn + "_mag_unc"] = np.full(b - a, 0.02)
    compiled = StarCatalog(pd.DataFrame(table), props=["parallax"]).compile(ic)
    mor = np.repeat(np.arange(b - a, dtype=np.int32), rows_per_star)
    pars = np.repeat(tm, rows_per_star, axis=0)
    pars *= 1 + 0.002 * np.random.RandomState(80 + rank).standard_normal(pars.shape)
    n_rows, pad = len(pars), ssh.pad * rows_per_star
    d_p, d_m = ctx.dev_alloc(pars.nbytes), ctx.dev_alloc(mor.nbytes)
    d_o, d_all = ctx.dev_alloc(pad * 8), ctx.dev_alloc(world * pad * 8)
    ctx.h2d(d_p, pars)
    ctx.h2d(d_m, mor)
    ms_k = timed(lambda: compiled.lnpost_device(d_p, n_rows, d_o, d_model_of_row=d_m))
    ms_g = timed(lambda: (compiled.lnpost_device(d_p, n_rows, d_o, d_model_of_row=d_m), comm.allgather(d_o, pad, d_all)))
    got = np.empty((world, pad))
    ctx.d2h(got, d_all)
    total_rows = n_stars * rows_per_star
    ms_f, same = None, None
    peer, _ = make_peer_group(ctx, rank, world, pad, dist, max_over_ranks) if world <= 8 else (None, "")
    if peer is not None:
        got_f = np.empty((world, pad))
        ctx.d2h(got_f, peer.lnpost(compiled, d_p, n_rows, d_model_of_row=d_m))
        same = bool(all(np.array_equal(got_f[r, :ssh.counts[r] * rows_per_star], got[r, :ssh.counts[r] * rows_per_star],
                                       equal_nan=True) for r in range(world)))
        ms_f = timed(lambda: peer.lnpost(compiled, d_p, n_rows, d_model_of_row=d_m))
        peer.close()
    out["catalog_10k_stars_sharded"] = {
        "value": total_rows / (ms_g * 1e-3), "unit": UNIT, "ms_per_step": ms_g, "ms_per_step_without_gather": ms_k,
        "ms_per_step_fused_peer_store": ms_f, "value_fused_peer_store": (total_rows / (ms_f * 1e-3)) if ms_f else None,
        "fused_identical_to_nccl": same,
        "stars_per_gpu": ssh.pad, "rows_per_gpu": pad,
        "finite_frac": float(np.isfinite(np.concatenate([got[r, :ssh.counts[r] * rows_per_star] for r in range(world)])).mean()),
        "scaling": "strong",
        "config": "configs[3] shape: 10 000 star models sharded over %d GPUs (each rank stages only its own stars), "
                  "100 rows per star per step (1e6 rows in total), ncclAllGather of lnpost inside the timed step" % world}
    for d in (d_p, d_m, d_o, d_all):
        ctx.dev_free(d)
    return out


def make_peer_group(ctx, rank, world, rows_per_rank, dist, max_over_ranks):
    """PeerGather on every rank, or None on every rank (the ranks agree, so that nobody waits in a barrier alone)."""
    from isochrones_b200 import parallel

    peer, err = None, ""
    try:
        peer = parallel.PeerGather(ctx, rank, world, rows_per_rank, parallel.torch_allgather_bytes(dist))
    except Exception as e:   # e.g. CUDA IPC not permitted in this container
        err = repr(e)[:300]
    if max_over_ranks(0.0 if peer is not None else 1.0) > 0.0:
        if peer is not None:
            peer.close()
        return None, err or "peer setup failed on another rank"
    return peer, ""


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
    ap.add_argument("--no-extras", action="store_true", help="skip the binary-star / sampler 'alt' workloads")
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

    numa_cpus = parallel.bind_to_gpu_numa(local_rank) if world > 1 else None   # before any pinned allocation
    ctx = _lib.default_context(local_rank)
    trk, bc, ic, truth, n_eep = build_workload(ctx=ctx, small=args.small)
    mags = truth_mags(ic, truth)
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
    tw0 = time.time()
    ms, launches = timed_device_loop(ctx, compiled, d_post, d_out, args.steps, args.warmup, barrier)
    clocks.mark(tw0, time.time())
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
    tw0 = time.time()
    t0 = time.perf_counter()
    for s in range(args.steps):
        mod.lnpost_batch(h_in[s % 2], out=h_out)
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    clocks.mark(tw0, time.time())
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
        # the same exchange fused into the lnpost kernel: every rank's kernel stores its rows into every rank's
        # receive buffer over NVLink peer mappings (no collective launch); checked once against the NCCL result
        if world <= 8:
            peer, err = make_peer_group(ctx, rank, world, BATCH, dist, max_over_ranks)
            if peer is None:
                allgather["fused_peer_store"] = {"unavailable": err}
            else:
                compiled.lnpost_device(d_post[0], BATCH, d_out)
                comm.allgather(d_out, BATCH, d_all)
                want = np.empty(world * BATCH)
                ctx.d2h(want, d_all)
                got = np.empty(world * BATCH)
                ctx.d2h(got, peer.lnpost(compiled, d_post[0], BATCH))
                same = bool(np.array_equal(got, want, equal_nan=True))
                for s in range(3):
                    peer.lnpost(compiled, d_post[s % N_BATCHES], BATCH)
                ctx.sync()
                barrier()
                ctx.timer_start()
                for s in range(args.steps):
                    peer.lnpost(compiled, d_post[s % N_BATCHES], BATCH)
                ms_p = max_over_ranks(ctx.timer_stop())
                barrier()
                allgather["fused_peer_store"] = {
                    "ms_per_step": ms_p / args.steps, "value": world * BATCH * args.steps / (ms_p * 1e-3),
                    "identical_to_nccl": same,
                    "how": "iso_lnpost_allgather_device: the lnpost kernel writes each row to all ranks' buffers through "
                           "CUDA-IPC peer mappings (NVLink), then a flag exchange; no NCCL call in the step"}
                peer.close()
        ctx.dev_free(d_all)
        sharded = None if args.no_extras else sharded_workloads(ctx, bc, args, rank, world, comm, barrier, max_over_ranks, dist)

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
                      ("scattered_valid", lambda s: scattered_batch(BATCH, seed=50 + s)),
                      ("grid_wide", lambda s: grid_wide_batch(BATCH, trk, seed=90 + s))):
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

    if world == 1 and not args.no_extras:
        alt.update(interp_workloads(ctx, ic, truth, n_eep, args, mod=mod))
        alt.update(extra_workloads(ctx, bc, args, peak))
    if world > 1 and sharded:
        alt.update(sharded)

    achieved = B_ALG * BATCH / (kernel_ms * 1e-3) / 1e9
    # The unit that actually limits the kernel: every corner record is a scattered 32-byte sector, and
    # the L1 data pipe of an SM retires one such sector ("wavefront") per clock, whatever level of the hierarchy
    # holds it.  28 sectors per row: 4 EEP-pair records x 3 (model pack, 96-byte records of 2 x 6 columns) + 16 corners
    # x 1 (BC pack, one chunk of 4 bands) = 896 B, i.e. exactly the algorithmic 8 x 6 x 8 + 16 x 4 x 8 bytes.
    sectors_per_row = 4 * 3 + 16 * 1
    info = ctx.info()
    sm_hz = float(clock_summary.get("sm_mhz") or peaks.get("sm_max_mhz") or 1965.0) * 1e6
    l1_peak = info["sm_count"] * sm_hz
    l1_ach = sectors_per_row * BATCH / (kernel_ms * 1e-3)
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
                   "sharding": "rows sharded across ranks, no data-path collective" if world > 1 else "single GPU",
                   "host_numa_binding": ("rank 0 bound to %d GPU-local cores" % len(numa_cpus)) if numa_cpus else "none"},
        "clocks": clock_summary,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": BATCH * 40, "d2h_bytes_per_step": BATCH * 8,
                "api": "BasicStarModel.lnpost_batch(pinned host array) -> iso_lnpost_batch (C ABI), chunked "
                       "H2D/kernel/D2H pipeline on two streams"},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "kernel": "iso_lnpost_kernel<1,false>", "algorithmic_bytes_per_eval": B_ALG,
                     "peak_source": peak_src, "kernel_ms": kernel_ms,
                     "l1_gather_bound": {"sectors_per_eval": sectors_per_row, "achieved_sectors_per_s": l1_ach,
                                         "peak_sectors_per_s": l1_peak, "frac": l1_ach / l1_peak,
                                         "note": "scattered 32-byte sector gathers at 1 sector per clock per SM "
                                                 "(%d SMs x %.0f MHz): the bound of a thread-per-row gather kernel; "
                                                 "lanes that share a 128-byte line are counted once by the hardware, "
                                                 "so the counter-based figure (ncu l1tex__data_pipe_lsu_wavefronts) is "
                                                 "lower" % (info["sm_count"], sm_hz / 1e6)}},
        "alt": alt,
    }
    if allgather:
        line["allgather"] = allgather
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(trk, bc, mod, seed=900)
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
