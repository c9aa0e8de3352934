#!/usr/bin/env python
"""Benchmark of the lnpost hot path (BASELINE.json: "lnpost evals/sec (batched walkers)").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1]): StarModel lnpost batch 1e6 — Sun-like single star (Teff, logg, feh + V, J, H, K
+ parallax) on the full MIST-shaped evolution-track grid (15 x 196 x 1710 nodes) and BC grid (70 x 26 x 18 x 13).
A "step" is one pass of the fused lnpost kernel over one batch of 1e6 parameter vectors; the timed steps rotate over
8 distinct batches (8 x 40 MB = 320 MB of inputs, larger than the 126 MB L2).  With N > 1 (one rank per GPU, launched by
torchrun or any launcher that exports RANK / LOCAL_RANK / WORLD_SIZE) every rank owns its own 1e6-row shard per step
(weak scaling, no data-path collective).  The ranks rendezvous through files (isochrones_b200.parallel.FileRendezvous):
no torch anywhere in this file.

Beyond the base contract the line carries
  roofline      the bound that binds the headline batch (L1 sector gathers) with the algorithmic / HBM figures next to it
                and the HBM-bound batch (grid_wide) as a second block; DRAM traffic from ncu captures of THIS kernel source
                (profiles/traffic.json carries a hash of isochrones_b200/csrc; a stale hash reports no traffic);
  cpu_baseline  the C port of the reference's CPU path (oracle/) on all host threads, bounded sample;
  alt           the other distributions / grids / BASELINE configs: prior-like, scattered, grid-wide, isochrone grid,
                11 bands, binary (configs[4]), one 256 x 2000 chain (configs[2]) with its CPU arm, the 10 000-star catalog
                FIT at its stated size (configs[3]: >= 1e5 lnpost per star) with its CPU arm, MultiNest-shaped entries;
  multi_gpu     (N > 1) fused peer-store gather vs ncclAllGather, sharded binary, catalog fit sharded by stars, one
                ensemble sharded over the ranks, per-rank host-link rates.
`--impl reference` times the CPU path (the C port under oracle/, all host threads) on a bounded sample of the workload.
"""
import argparse
import hashlib
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
BANDS11 = ("J", "H", "K", "G", "BP", "RP", "W1", "W2", "W3", "TESS", "Kepler")     # mist/bc.py:159
PACK_COLUMNS = ("Teff", "logg", "feh", "Mbol", "age", "dt_deep", "nu_max", "delta_nu")
ISO_PACK_COLUMNS = ("Teff", "logg", "feh", "Mbol", "mass", "dm_deep", "nu_max", "delta_nu")
METRIC = "lnpost evals/sec (batched walkers)"
UNIT = "evals/s"


def b_alg(n_stars, n_bands):
    """Algorithmic bytes per lnpost evaluation: 8 [(S + 4) + 1 + S (8 * 6 + 16 n_b)]."""
    return 8.0 * ((n_stars + 4) + 1 + n_stars * (8 * 6 + 16 * n_bands))


B_ALG = b_alg(1, 4)      # 944 B: single star, 4 bands


def host_threads():
    """Host cores this process may use (the CPU arm uses all of them through OpenMP's num_threads clause)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


KERNEL_SOURCES = ("iso_common.cuh", "iso_prior.cuh", "iso_philox.cuh", "iso_lnpost_row.cuh", "iso_lnpost_kernel.cuh",
                  "iso_lnpost_k1.cu")


def kernel_source_hash():
    """SHA-256 over the sources the fused single-star lnpost kernel is compiled from: ties an ncu capture
    (profiles/traffic.json) to the kernel being timed."""
    d = os.path.join(ROOT, "isochrones_b200", "csrc")
    h = hashlib.sha256()
    for name in KERNEL_SOURCES:
        with open(os.path.join(d, name), "rb") as f:
            h.update(name.encode() + b"\0" + f.read())
    return h.hexdigest()[:16]


# ---------------------------------------------------------------------------------------------------------------------
# workload
# ---------------------------------------------------------------------------------------------------------------------
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


def truth_mags(ic, truth, bands=BANDS):
    """Observed magnitudes of the b200 arm = model magnitudes at the truth, rounded to 1 mmag (the product's own
    ``interp_mag``; the reference arm derives the same numbers from the oracle, see ``oracle_truth_mags``)."""
    _, _, _, mags = ic.interp_mag(list(truth), list(bands))
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


def make_model(ic, mags, bands=BANDS):
    import isochrones_b200 as ib

    obs = {b: (m, 0.02) for b, m in zip(bands, mags)}
    return ib.BasicStarModel(ic, Teff=(5772.0, 80.0), logg=(4.44, 0.1), feh=(0.0, 0.1), parallax=(10.0, 0.1), **obs)


def scattered_batch(n, seed):
    """Rows spread uniformly over the part of the track grid where every EEP is populated and the BC lookup is in
    range: all rows do the full work and the gathers scatter over the grid (L2-bound case)."""
    rng = np.random.RandomState(seed)
    p = np.empty((n, 5))
    p[:, 0] = np.exp(rng.uniform(np.log(0.7), np.log(5.9), n))     # mass
    p[:, 1] = rng.uniform(1.0, 1700.0, n)                           # eep
    p[:, 2] = rng.uniform(-2.0, 0.45, n)                            # feh
    p[:, 3] = rng.uniform(20.0, 199.0, n)                           # distance
    p[:, 4] = rng.uniform(0.0, 0.99, n)                             # AV
    return p


def grid_wide_batch(n, trk, seed):
    """Rows spread over the WHOLE populated track grid (every [Fe/H], masses 0.1-100, EEPs up to the end of the
    shorter of the bracketing tracks): the gathers touch more model-grid nodes than the 126 MB L2 holds — the
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


def catalog_truths(n_stars, wide, seed=8):
    """Truth points of a synthetic catalog on the isochrone grid: ``wide`` spreads the stars over the whole populated
    grid (what a real catalog does), otherwise over a Sun-like corner of it."""
    from isochrones_b200 import synthetic as syn

    rng = np.random.RandomState(seed)
    t = np.tile(syn.default_truth("iso", n_stars=1), (n_stars, 1))
    if wide:
        t[:, 0] = rng.uniform(210.0, 1400.0, n_stars)
        t[:, 1] = rng.uniform(7.5, 10.1, n_stars)
        t[:, 2] = rng.uniform(-3.5, 0.45, n_stars)
    else:
        t[:, 0] = rng.uniform(300.0, 900.0, n_stars)
        t[:, 1] = rng.uniform(9.0, 9.9, n_stars)
        t[:, 2] = rng.uniform(-0.5, 0.3, n_stars)
    t[:, 3] = rng.uniform(50.0, 400.0, n_stars)
    t[:, 4] = rng.uniform(0.0, 0.5, n_stars)
    return t


def catalog_table(ic, t, narrow=None):
    """Measurement table (VJHK magnitudes + parallax) of the stars at truths ``t``; stars whose truth falls off the
    grids take the matching row of ``narrow`` instead.  Returns ``(DataFrame, truths actually used)``."""
    import pandas as pd

    _, _, _, mg = ic.interp_mag([np.ascontiguousarray(t[:, j]) for j in range(5)], list(BANDS))
    bad = ~np.isfinite(mg).all(axis=1)
    if bad.any():
        if narrow is None:
            raise ValueError("catalog truth outside the grids")
        t = t.copy()
        t[bad] = narrow[bad]
        _, _, _, mg2 = ic.interp_mag([np.ascontiguousarray(t[:, j]) for j in range(5)], list(BANDS))
        mg[bad] = mg2[bad]
    table = {"parallax": 1000.0 / t[:, 3], "parallax_unc": np.full(len(t), 0.1)}
    for j, b in enumerate(BANDS):
        table[b + "_mag"] = mg[:, j]
        table[b + "_mag_unc"] = np.full(len(t), 0.02)
    return pd.DataFrame(table), t


# ---------------------------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------------------------
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


# ---------------------------------------------------------------------------------------------------------------------
# rank plumbing (no torch): files in a directory shared by the ranks of the node
# ---------------------------------------------------------------------------------------------------------------------
class Ranks(object):
    def __init__(self):
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", str(self.rank)))
        self.rz = None
        if self.world > 1:
            from isochrones_b200 import parallel

            self.rz = parallel.FileRendezvous.from_env(timeout=600.0)

    def barrier(self):
        if self.rz:
            self.rz.barrier()

    def max(self, x):
        return self.rz.max(x) if self.rz else float(x)

    def all(self, flag):
        return self.rz.all(flag) if self.rz else bool(flag)

    def allgather_bytes(self, payload):
        return self.rz.allgather_bytes(payload) if self.rz else [payload]

    def broadcast(self, payload):
        return self.rz.broadcast(payload) if self.rz else payload

    def close(self):
        if self.rz:
            self.rz.close()


# ---------------------------------------------------------------------------------------------------------------------
# CPU legs (the only places that touch oracle/)
# ---------------------------------------------------------------------------------------------------------------------
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


def reference_as_shipped(budget_calls=4000):
    """The UNMODIFIED reference (pure Python + numba) on this box's host cores: its sources travel as an archive under
    oracle/_ref/ (oracle/build_ref.py, git-ignored) and are timed by oracle/time_reference.py in a subprocess — the scalar
    `BasicStarModel.lnpost(p)` loop on one core, and the reference's own batch recipe (a process pool over all cores)."""
    import shutil
    import tarfile
    import tempfile

    archive = os.path.join(ROOT, "oracle", "_ref", "isochrones_reference.tar.gz")
    if not os.path.exists(archive):
        return {"unavailable": "oracle/_ref/isochrones_reference.tar.gz not built (run __graft_entry__.build() where /root/reference exists)"}
    tmp = tempfile.mkdtemp(prefix="iso_ref_")
    try:
        with tarfile.open(archive) as tar:
            tar.extractall(tmp)
        env = dict(os.environ, NUMBA_CACHE_DIR=os.path.join(tmp, "numba_cache"))
        env.pop("OMP_NUM_THREADS", None)
        r = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "time_reference.py"), "--root", tmp, "--pool", "--json",
                            "--calls", str(budget_calls)], capture_output=True, text=True, timeout=600, env=env)
        lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
        if r.returncode != 0 or not lines:
            return {"unavailable": "time_reference.py failed: %s" % (r.stderr[-300:] or r.stdout[-300:])}
        t = json.loads(lines[-1])
        out = {"kind": "reference", "unit": UNIT, "value": t["lnpost_scalar"]["evals_per_s"], "cores": 1,
               "us_per_call": t["lnpost_scalar"]["us_per_call"],
               "sample": "%d calls of the unmodified reference's BasicStarModel.lnpost(p) (pure Python + numba), posterior-like "
                         "rows of the bench workload, one core" % t["lnpost_scalar"]["calls"],
               "interp_mag_points_per_s": t["interp_mag_arrays"]["points_per_s"],
               "interp_value_points_per_s": t["interp_value_arrays"]["points_per_s"]}
        if "lnpost_scalar_pool" in t:
            out["pool"] = {"value": t["lnpost_scalar_pool"]["evals_per_s"], "cores": t["lnpost_scalar_pool"]["processes"],
                           "sample": "%d calls over multiprocessing.Pool(%d): the reference's own batch recipe"
                                     % (t["lnpost_scalar_pool"]["calls"], t["lnpost_scalar_pool"]["processes"])}
        return out
    except Exception as e:
        return {"unavailable": repr(e)[:300]}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def cpu_chain(model, p0, n_steps, seed):
    """configs[2] on the host cores: the oracle's lnpost driven by the same stretch move (oracle.stretch_move; the
    walkers of a half-step fan out over all threads).  Returns evals/s and the wall time."""
    from oracle import oracle

    om = oracle.StarModel(model)
    threads = host_threads()
    oracle.stretch_move(om, p0, 20, seed, n_threads=threads, store=False)
    t0 = time.perf_counter()
    _, _, _, _, acc = oracle.stretch_move(om, p0, n_steps, seed, n_threads=threads, store=False)
    dt = time.perf_counter() - t0
    n = len(p0) * n_steps
    return {"value": n / dt, "unit": UNIT, "seconds": dt, "cores": threads, "kind": "port",
            "acceptance_fraction": float(acc[0]) / n,
            "sample": "the whole %d x %d run: oracle lnpost + the same stretch move (oracle.stretch_move), half-step walkers "
                      "over %d OpenMP threads" % (len(p0), n_steps, threads)}


def cpu_catalog_fit(cat, ic, truths, n_sample, nw, n_steps, seed):
    """configs[3] on the host cores, bounded sample: ``n_sample`` stars of the catalog, one chain each, chains spread over
    all threads (the reference's own batch recipe is one process per star, notebooks/batch-demo.ipynb:122-124)."""
    from oracle import oracle

    threads = host_threads()
    models = []
    for i, m in enumerate(cat.iter_models(ic)):
        if i >= n_sample:
            break
        models.append(oracle.StarModel(m))
    p0 = walkers_around(truths[:len(models)], nw, seed)
    oracle.stretch_move(models, p0, 2, seed, n_threads=threads, store=False)
    t0 = time.perf_counter()
    oracle.stretch_move(models, p0, n_steps, seed, n_threads=threads, store=False)
    dt = time.perf_counter() - t0
    n = len(models) * nw * n_steps
    return {"value": n / dt, "unit": UNIT, "seconds": dt, "cores": threads, "kind": "port", "stars_per_s": len(models) / dt,
            "sample": "%d of the catalog's stars x %d walkers x %d steps (%d evals), one chain per star over %d OpenMP threads"
                      % (len(models), nw, n_steps, n, threads)}


def walkers_around(truths, nw, seed):
    """Initial ensembles ``[n_stars, nw, 5]``: a tight ball around every star's truth (relative 1e-3)."""
    rng = np.random.RandomState(seed)
    p0 = np.repeat(truths[:, None, :], nw, axis=1)
    return np.ascontiguousarray(p0 * (1.0 + 1e-3 * rng.standard_normal(p0.shape)))


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


# ---------------------------------------------------------------------------------------------------------------------
# device loops
# ---------------------------------------------------------------------------------------------------------------------
def stage(ctx, batches):
    ptrs = []
    for b in batches:
        b = np.ascontiguousarray(b)
        d = ctx.dev_alloc(b.nbytes)
        ctx.h2d(d, b)
        ptrs.append(d)
    return ptrs


def timed_device_loop(ctx, compiled, d_batches, d_out, steps, warmup, n_rows=BATCH, barrier=None, d_mor=None):
    for s in range(warmup):
        compiled.lnpost_device(d_batches[s % len(d_batches)], n_rows, d_out, d_model_of_row=d_mor)
    ctx.sync()
    if barrier:
        barrier()
    l0 = ctx.launch_count()
    ctx.timer_start()
    for s in range(steps):
        compiled.lnpost_device(d_batches[s % len(d_batches)], n_rows, d_out, d_model_of_row=d_mor)
    ms = ctx.timer_stop()
    ctx.sync()
    if barrier:
        barrier()
    return ms, ctx.launch_count() - l0


def device_workload(ctx, compiled, gen, args, peak, bytes_per_eval, n_batches=4, n_rows=BATCH):
    """One ``alt`` entry: ``gen(s) -> rows`` staged in HBM, the fused kernel timed with CUDA events."""
    d_b = stage(ctx, [gen(s) for s in range(n_batches)])
    d_out = ctx.dev_alloc(n_rows * 8)
    ms, _ = timed_device_loop(ctx, compiled, d_b, d_out, args.steps, args.warmup, n_rows=n_rows)
    out = np.empty(n_rows)
    ctx.d2h(out, d_out)
    for d in d_b + [d_out]:
        ctx.dev_free(d)
    k_ms = ms / args.steps
    return {"value": n_rows / (k_ms * 1e-3), "unit": UNIT, "ms_per_step": k_ms, "finite_frac": float(np.isfinite(out).mean()),
            "algorithmic_bytes_per_eval": bytes_per_eval,
            "algorithmic_frac_of_hbm_peak": bytes_per_eval * n_rows / (k_ms * 1e-3) / 1e9 / peak}


def interp_workloads(ctx, ic, truth, n_eep, args, mod):
    """BASELINE.json configs[0] at bench size: the standalone interpolation entry points (DFInterpolator /
    interp_value, interp_mag) on 1e6 points through the host-pointer C ABI (pinned buffers; H2D + kernel + D2H), the
    interpolation kernels alone, and the latency of the reference-shaped small calls (scalar lnpost, one 128-row
    half-step, MultiNest's per-live-point mnest_prior + mnest_loglike)."""
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
    # the reference's scalar interface, as emcee / MultiNest call it
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
        "config": "BasicStarModel.lnpost(p): one row per call through the C ABI (kernel reads / writes page-locked memory, "
                  "one launch + one sync); the reference's own scalar call takes 68-108 us in the build container"}
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
    cube = [0.31, 0.52, 0.44, 0.05, 0.1]
    for _ in range(200):
        c = list(cube)
        mod.mnest_prior(c, 5, 5)
        mod.mnest_loglike(c, 5, 5)
    t0 = time.perf_counter()
    for _ in range(reps):
        c = list(cube)
        mod.mnest_prior(c, 5, 5)
    dt_p = (time.perf_counter() - t0) / reps
    t0 = time.perf_counter()
    for _ in range(reps):
        mod.mnest_loglike(c, 5, 5)
    dt_l = (time.perf_counter() - t0) / reps
    live = np.ascontiguousarray(np.random.RandomState(1).random_sample((400, 5)))
    work = live.copy()
    for _ in range(50):
        work[:] = live
        mod.mnest_lnpost_batch(work)
    t0 = time.perf_counter()
    for _ in range(500):
        work[:] = live
        mod.mnest_lnpost_batch(work)
    dt_b = (time.perf_counter() - t0) / 500
    out["multinest_calls"] = {
        "mnest_prior_us_per_call": dt_p * 1e6, "mnest_loglike_us_per_call": dt_l * 1e6,
        "fused_400_live_points_us_per_call": dt_b * 1e6, "value": 400.0 / dt_b, "unit": UNIT,
        "config": "MultiNest's calling sequence (starmodel.py:797, 1637-1645): per live point mnest_prior(cube) [no device "
                  "allocation: bounds in the kernel parameter block] then mnest_loglike(cube); and "
                  "BasicStarModel.mnest_lnpost_batch on 400 live points (cube -> parameters -> lnpost in one launch)"}
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
    for name, fn, bpp in (
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
                     "algorithmic_bytes_per_point": bpp, "achieved_gb_s": bpp * BATCH / (k_ms * 1e-3) / 1e9,
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
        "config": "ModelGridInterpolator.interp_mag(5 parameter arrays, VJHK, out=pinned arrays) -> iso_interp_mags_cols, "
                  "host arrays in / out (96 B per point over PCIe)"}
    return out


def iso_world(ctx, bc):
    """Isochrone-grid interpolator, single / binary models and truths shared by the alt and multi-GPU workloads."""
    import isochrones_b200 as ib
    from isochrones_b200 import synthetic as syn

    iso = syn.make_iso_grid(columns=ISO_PACK_COLUMNS)
    ic = ib.ichrone_from_arrays("iso", iso, bc, ctx=ctx)
    t2 = syn.default_truth("iso", n_stars=2)
    _, _, _, m2 = ic.interp_mag([t2[0]] + list(t2[2:]), list(BANDS))
    obs = {b: (float(np.round(m, 3)) - 0.35, 0.02) for b, m in zip(BANDS, m2)}
    binary = ib.BinaryStarModel(ic, Teff=(5772.0, 80.0), logg=(4.44, 0.1), feh=(0.0, 0.1), parallax=(10.0, 0.1), **obs)
    t1 = syn.default_truth("iso", n_stars=1)
    _, _, _, m1 = ic.interp_mag(list(t1), list(BANDS))
    single = ib.SingleStarModel(ic, Teff=(5772.0, 80.0), logg=(4.44, 0.1), feh=(0.0, 0.1), parallax=(10.0, 0.1),
                                **{b: (float(np.round(m, 3)), 0.02) for b, m in zip(BANDS, m1)})
    return ic, single, binary, t1, t2


def binary_batch(n, t2, seed):
    from isochrones_b200 import synthetic as syn

    p = syn.posterior_like_batch("iso", n, t2, seed=seed)
    p[:, :2] = -np.sort(-p[:, :2], axis=1)
    return p


def catalog_fit(ctx, ic, rk, comm, args, n_stars=10_000, nw=256, n_steps=400, thin=10, with_cpu=True):
    """BASELINE configs[3] at its stated size, as a FIT: 10 000 independent star models x >= 1e5 lnpost each (256 walkers
    x 400 steps = 102 400; the reference's own recipe is 300 walkers x 300 steps, starmodel.py:889), stars sharded over
    the ranks (1250 per GPU at N = 8), one on-device chain per star (iso_sampler_*, catalog mode; more chains than
    resident CTAs are worked through in 16-step segments from a queue), walkers and samples never leave HBM (the running
    moments take every 10th ensemble); the per-star posterior means / spreads come from
    the kernel's running moments and are all-gathered (ncclAllGather) so that every rank holds the catalog's summary."""
    from isochrones_b200 import parallel
    from isochrones_b200.catalog import StarCatalog
    from isochrones_b200.sampler import DeviceEnsembleSampler

    narrow = catalog_truths(n_stars, wide=False)
    sh = parallel.RowSharder(n_stars, rk.world, rk.rank)
    a, b = sh.bounds()
    t0 = time.perf_counter()
    df, truths = catalog_table(ic, catalog_truths(n_stars, wide=True)[a:b], narrow=narrow[a:b])
    cat = StarCatalog(df, props=["parallax"])
    compiled = cat.compile(ic)
    t_compile = time.perf_counter() - t0
    p0 = walkers_around(truths, nw, seed=31 + rk.rank)
    # walkers that fall off the grids start at the truth itself (a NaN start is refused, as in emcee)
    flat = p0.reshape(-1, 5)
    mor = np.repeat(np.arange(b - a, dtype=np.int32), nw)
    bad = ~np.isfinite(compiled.lnpost(flat, model_of_row=mor))
    flat[bad] = np.repeat(truths, nw, axis=0)[bad]
    smp = DeviceEnsembleSampler(compiled, nw, p0, seed=5, n_chains=b - a, moments=True)
    smp.run_mcmc(20, store=False, fetch=False)          # burn-in + warm-up of the launch path
    smp.reset()
    ctx.sync()
    rk.barrier()
    ctx.timer_start()
    smp.run_mcmc(n_steps, thin=thin, store=False, fetch=False)
    ms = rk.max(ctx.timer_stop())
    rk.barrier()
    mean, std, cnt = smp.moments()
    acc = smp.acceptance_fraction
    # gather of the per-star summaries: every rank ends up with [n_stars, 2 ndim + 1] running sums
    n_mom = 2 * 5 + 1
    ms_gather = None
    if comm is not None:
        d_all = ctx.dev_alloc(rk.world * sh.pad * n_mom * 8)
        d_send = smp.moments_device()
        if b - a < sh.pad:                 # ragged last blocks: pad the send buffer
            d_pad = ctx.dev_alloc(sh.pad * n_mom * 8)
            ctx.memset(d_pad, 0, sh.pad * n_mom * 8)
            tmp = np.empty((b - a, n_mom))
            ctx.d2h(tmp, d_send)
            ctx.h2d(d_pad, tmp)
            d_send = d_pad
        for _ in range(3):
            comm.allgather(d_send, sh.pad * n_mom, d_all)
        ctx.sync()
        rk.barrier()
        ctx.timer_start()
        for _ in range(10):
            comm.allgather(d_send, sh.pad * n_mom, d_all)
        ms_gather = rk.max(ctx.timer_stop()) / 10
        got = np.empty((rk.world, sh.pad, n_mom))
        ctx.d2h(got, d_all)
        total_samples = float(sum(got[r, :sh.counts[r], -1].sum() for r in range(rk.world)))
        ctx.dev_free(d_all)
    else:
        total_samples = float(cnt.sum())
    smp.close()
    evals = n_stars * nw * n_steps
    out = {
        "value": evals / (ms * 1e-3), "unit": UNIT, "seconds": ms * 1e-3, "stars_per_s": n_stars / (ms * 1e-3),
        "evals_per_star": nw * n_steps, "stars": n_stars, "stars_per_gpu": sh.pad, "walkers": nw, "steps": n_steps,
        "acceptance_fraction": float(np.mean(acc)), "posterior_sigma_eep_median": float(np.median(std[:, 0])),
        "samples_in_gathered_summaries": total_samples, "compile_seconds": t_compile,
        "summary_gather_ms": ms_gather, "summary_gather_bytes_per_rank": sh.pad * n_mom * 8, "scaling": "strong",
        "pcie_bytes_per_eval": 0,
        "config": "configs[3] at size: %d star models (iso grid, VJHK + parallax, truths spread over the whole populated "
                  "grid) x %d walkers x %d steps = %d lnpost per star, stars sharded over %d GPU(s), one chain per star, all "
                  "steps in one launch (resident CTAs claim 16-step segments of chains from a queue); summaries = running "
                  "moments gathered with ncclAllGather"
                  % (n_stars, nw, n_steps, nw * n_steps, rk.world)}
    if with_cpu and rk.rank == 0:
        out["cpu"] = cpu_catalog_fit(cat, ic, truths, n_sample=min(96, b - a), nw=nw, n_steps=50, seed=5)
        out["vs_cpu"] = out["value"] / out["cpu"]["value"]
    return out


def sharded_ensemble(ctx, single, t1, rk, args, n_walkers=1 << 20, n_steps=50):
    """ONE ensemble of 2^20 walkers sharded over the ranks (iso_ensemble_*): per half-step a rank evaluates its block of
    the active half, and each proposal gathers its partner walker from the HBM of the rank that owns it over NVLink peer
    mappings (the multi-GPU consumer of the fused peer exchange); a run ends with one replication of every rank's blocks.
    Checked: all ranks end with identical copies, and identical to a single-rank run."""
    from isochrones_b200 import synthetic as syn
    from isochrones_b200.sampler import ShardedEnsembleSampler

    p0 = syn.posterior_like_batch("iso", n_walkers, t1, seed=77)
    bad = ~np.isfinite(single.lnpost_batch(p0))
    p0[bad] = t1
    ens = ShardedEnsembleSampler(single.compiled, n_walkers, p0, seed=13, rank=rk.rank, world=rk.world,
                                 allgather_bytes=rk.allgather_bytes)
    ens.run_mcmc(2, store=False, fetch=False)
    n_short = 10

    def timed_run(n):
        # the ranks are aligned on the DEVICE: a one-step run ends with every rank waiting for every rank's replication,
        # so all ranks leave it within microseconds (a file barrier would leave them milliseconds apart, and the first
        # half-step of the timed run would wait for the slowest rank to arrive)
        rk.barrier()
        ens.run_mcmc(1, store=False, fetch=False)
        ctx.timer_start()
        ens.run_mcmc(n, store=False, fetch=False)
        return rk.max(ctx.timer_stop())

    ms_short = timed_run(n_short)
    ms = timed_run(n_steps)
    pos, lnp, acc, prop = ens.state()
    digest = hashlib.sha256(pos.tobytes() + lnp.tobytes()).hexdigest()
    same_copies = len(set(rk.allgather_bytes(digest.encode()))) == 1
    accepted = sum(int(v) for v in rk.allgather_bytes(str(acc).encode()))
    rk.barrier()
    ens.close()
    same_single = None
    total_steps = 2 + (1 + n_short) + (1 + n_steps)
    if rk.rank == 0:
        one = ShardedEnsembleSampler(single.compiled, n_walkers, p0, seed=13)
        one.run_mcmc(total_steps, store=False, fetch=False)
        p1, l1, _, _ = one.state()
        same_single = hashlib.sha256(p1.tobytes() + l1.tobytes()).hexdigest() == digest
        one.close()
    rk.barrier()
    evals = n_walkers * n_steps
    marginal = (ms - ms_short) / (2 * (n_steps - n_short))
    return {"value": evals / (ms * 1e-3), "unit": UNIT, "ms_per_half_step": ms / (2 * n_steps), "walkers": n_walkers,
            "steps": n_steps, "ms_per_half_step_marginal": marginal,
            "run_overhead_ms": ms_short - 2 * n_short * marginal,
            "acceptance_fraction": accepted / float(n_walkers * total_steps),
            "copies_identical_across_ranks": same_copies, "identical_to_single_rank_run": same_single, "scaling": "strong",
            "config": "one ensemble of %d walkers on the iso grid sharded over %d rank(s), %d steps in one run; each "
                      "proposal gathers its partner walker (40 B) from the owning rank's HBM over NVLink peer mappings inside "
                      "the evaluation kernel, flag exchange per half-step, no collective launch; the run ends with one "
                      "replication of every rank's blocks into every copy (inside the timed region: run_overhead_ms = "
                      "launch + replication, from a 10-step run timed beside it)" % (n_walkers, rk.world, n_steps)}


def extra_workloads(ctx, bc, args, peak, rk):
    """BASELINE.json configs 2, 3 and 4 on one GPU, the isochrone grid and the 11 default bands."""
    import isochrones_b200 as ib
    from isochrones_b200 import synthetic as syn
    from isochrones_b200.catalog import StarCatalog
    from isochrones_b200.sampler import DeviceEnsembleSampler

    out = {}
    ic, single, binary, t1, t2 = iso_world(ctx, bc)
    e = device_workload(ctx, binary.compiled, lambda s: binary_batch(BATCH, t2, 70 + s), args, peak, b_alg(2, 4))
    e["config"] = "configs[4] on one GPU: BinaryStarModel, iso grid 107x15x1710, 4 bands + parallax, 1e6 rows"
    out["binary_iso_posterior_like"] = e
    # starfit's default grid (starfit.py:210-211): single star on the isochrone grid, both distributions
    e = device_workload(ctx, single.compiled, lambda s: syn.posterior_like_batch("iso", BATCH, t1, seed=170 + s), args, peak, B_ALG)
    e["config"] = "single star on the isochrone grid (starfit's default), posterior-like rows"
    out["iso_single_posterior_like"] = e
    b1 = [single.bounds(p) for p in single.param_names]
    e = device_workload(ctx, single.compiled, lambda s: syn.prior_like_batch("iso", BATCH, b1, seed=270 + s), args, peak, B_ALG)
    e["config"] = "single star on the isochrone grid, prior-like rows (uniform over the bounds of every parameter)"
    out["iso_single_prior_like"] = e

    # configs[2]: one 256 x 2000 chain, five repeats (the run is latency-bound: one CTA, 4000 dependent half-steps)
    nw, nsteps = 256, 2000
    p0 = syn.posterior_like_batch("iso", nw, t1, seed=4)
    p0[~np.isfinite(single.lnpost_batch(p0))] = t1
    smp = DeviceEnsembleSampler(single.compiled, nw, p0, seed=4)
    smp.run_mcmc(50, store=False)
    times = []
    for _ in range(5):
        t0 = time.perf_counter()
        smp.run_mcmc(nsteps, thin=10)
        times.append(time.perf_counter() - t0)
    times.sort()
    entry = {"value": nw * nsteps / times[2], "unit": UNIT, "seconds": times[2], "seconds_min": times[0], "seconds_max": times[-1],
             "repeats": 5, "acceptance_fraction": float(smp.acceptance_fraction[0]),
             "config": "configs[2]: stretch move a=2, 256 walkers x 2000 steps, one persistent CTA, one launch per run "
                       "(latency-bound: 4000 dependent half-steps); median of 5 runs incl. the D2H of the thinned chain"}
    smp.close()
    entry["cpu"] = cpu_chain(single, p0, nsteps, seed=4)
    entry["vs_cpu"] = entry["value"] / entry["cpu"]["value"]
    out["emcee_256x2000_one_chain"] = entry

    n_chains, steps_c = 1184, 100
    p0c = np.stack([syn.posterior_like_batch("iso", nw, t1, seed=1000 + c) for c in range(8)])
    for c in range(8):
        p0c[c][~np.isfinite(single.lnpost_batch(p0c[c]))] = t1
    p0c = np.ascontiguousarray(np.tile(p0c, (n_chains // 8, 1, 1)))
    smc = DeviceEnsembleSampler(single.compiled, nw, p0c, seed=5, n_chains=n_chains)
    smc.run_mcmc(10, store=False, fetch=False)
    t0 = time.perf_counter()
    smc.run_mcmc(steps_c, store=False, fetch=False)
    dt = time.perf_counter() - t0
    out["emcee_256_walkers_x_1184_chains"] = {
        "value": n_chains * nw * steps_c / dt, "unit": UNIT, "seconds": dt,
        "acceptance_fraction": float(np.mean(smc.acceptance_fraction)),
        "config": "1184 independent 256-walker ensembles x 100 steps in one launch (8 CTAs per SM)"}
    smc.close()

    # configs[3] as a batch: rows of 10 000 star models in one launch through model_of_row (stars over the whole grid)
    n_stars, rows_per_star = 10_000, 100
    df, t = catalog_table(ic, catalog_truths(n_stars, wide=True), narrow=catalog_truths(n_stars, wide=False))
    t0 = time.perf_counter()
    compiled = StarCatalog(df, props=["parallax"]).compile(ic)
    t_compile = time.perf_counter() - t0
    mor = np.repeat(np.arange(n_stars, dtype=np.int32), rows_per_star)
    rng = np.random.RandomState(8)
    pars = np.repeat(t, rows_per_star, axis=0)
    pars *= 1 + 0.002 * rng.standard_normal(pars.shape)
    d_m = ctx.dev_alloc(mor.nbytes)
    ctx.h2d(d_m, mor)
    d_p, d_o = stage(ctx, [pars])[0], ctx.dev_alloc(len(pars) * 8)
    ms, _ = timed_device_loop(ctx, compiled, [d_p], d_o, args.steps, args.warmup, n_rows=len(pars), d_mor=d_m)
    res = np.empty(len(pars))
    ctx.d2h(res, d_o)
    for d in (d_p, d_m, d_o):
        ctx.dev_free(d)
    k_ms = ms / args.steps
    out["catalog_10k_stars_batch"] = {
        "value": len(pars) / (k_ms * 1e-3), "unit": UNIT, "ms_per_step": k_ms, "finite_frac": float(np.isfinite(res).mean()),
        "compile_seconds": t_compile,
        "config": "configs[3] shape as ONE batch: 10 000 star models (iso grid, 4 bands + parallax, stars spread over the whole "
                  "grid) in HBM, 1e6 rows (100 per star) in one launch through model_of_row"}
    compiled.close()
    # configs[3] at its stated size, as a fit (one GPU here; sharded by stars under multi_gpu)
    if rk.world == 1:
        out["catalog_10k_stars_fit"] = catalog_fit(ctx, ic, rk, None, args)
        out["sharded_ensemble_1M_walkers"] = sharded_ensemble(ctx, single, t1, rk, args)
        # the reference's DEFAULT fit: nested sampling, 1000 live points, evidence tolerance 0.5 (starmodel.py:667-671,
        # 717-802; published: "about 14 minutes on my laptop" for a binary, docs/multiple.ipynb) over the full prior box
        t0 = time.perf_counter()
        res = single.fit_nested(n_live_points=1000, seed=1, max_batches=4000)
        dt = time.perf_counter() - t0
        out["nested_fit_1000_live_points"] = {
            "value": res.n_evals / dt, "unit": UNIT, "seconds": dt, "n_evals": res.n_evals, "n_iter": res.n_iter,
            "logZ": res.logZ, "logZ_err": res.logZ_err, "information_nats": res.information, "efficiency": res.efficiency,
            "finite_fraction_of_prior_box": res.finite_fraction, "converged": res.converged,
            "n_launches": res.n_batches + 1, "n_ellipsoids_max": res.n_ellipsoids_max,
            "posterior_mean": [float(v) for v in res.mean()], "posterior_std": [float(v) for v in res.std()],
            "config": "BasicStarModel.fit_nested on the isochrone-grid single star over the full prior box: host-driven nested "
                      "sampling (MultiNest-style union of bounding ellipsoids), every batch of 8192 candidate live points = one "
                      "iso_mnest_lnpost_batch launch (cube -> parameters -> lnpost)"}
    return out


def eleven_band_workload(ctx, trk, truth, n_eep, args, peak):
    """Single star with the reference's 11 default bands (mist/bc.py:159): three 4-band chunks of the BC pack."""
    import isochrones_b200 as ib
    from isochrones_b200 import synthetic as syn

    bc11 = syn.make_bc_grid(bands=BANDS11)
    ic11 = ib.ichrone_from_arrays("track", trk, bc11, ctx=ctx)
    mod11 = make_model(ic11, truth_mags(ic11, truth, BANDS11), BANDS11)
    e = device_workload(ctx, mod11.compiled, lambda s: syn.posterior_like_batch("track", BATCH, truth, n_eep=n_eep, seed=400 + s),
                        args, peak, b_alg(1, 11))
    e["config"] = "single star, the 11 default bands of the reference (J H K G BP RP W1 W2 W3 TESS Kepler) + Teff/logg/feh + " \
                  "parallax, track grid, posterior-like rows; BC pack of 12 columns = 3 chunks of 4 bands"
    mod11.compiled.close()
    return e


# ---------------------------------------------------------------------------------------------------------------------
# multi-GPU
# ---------------------------------------------------------------------------------------------------------------------
def make_peer_group(ctx, rk, rows_per_rank, comm):
    """The library's own choice of the per-step exchange (parallel.row_gather: all ranks decide together): the fused
    PeerGather, or (None, why) when it handed out the NCCL form — which the bench times separately anyway."""
    from isochrones_b200 import parallel

    g, why = parallel.row_gather(ctx, rk.rank, rk.world, rows_per_rank, rk.allgather_bytes, comm=comm)
    if isinstance(g, parallel.PeerGather):
        return g, ""
    g.close()
    return None, why or "peer setup failed"


def multi_gpu(ctx, compiled, d_post, d_out, bc, args, rk, h_in, h_out, mod):
    """Everything that needs N > 1: the all-gather a sampler's acceptance step needs (ncclAllGather vs the fused
    peer-store kernel, which must agree bit for bit), configs[4] sharded by rows, configs[3] sharded by stars as a fit,
    one ensemble sharded over the ranks, and the per-rank host-link rates behind the end-to-end number."""
    from isochrones_b200 import parallel

    out = {}
    world = rk.world
    comm = parallel.NcclGather(ctx, rk.rank, world, exchange=rk.broadcast)

    def timed(fn, align):
        """`align` runs a few untimed steps that contain the exchange, so the ranks enter the timed region together."""
        rk.barrier()
        for _ in range(3):
            align()
        ctx.timer_start()
        for _ in range(args.steps):
            fn()
        ms = rk.max(ctx.timer_stop()) / args.steps
        rk.barrier()
        return ms

    # ---- weak scaling with the gather: 1e6 rows per GPU, every rank receives all results ---------------------------
    d_all = ctx.dev_alloc(world * BATCH * 8)
    step_i = [0]

    def nccl_step():
        compiled.lnpost_device(d_post[step_i[0] % N_BATCHES], BATCH, d_out)
        comm.allgather(d_out, BATCH, d_all)
        step_i[0] += 1

    ms_g = timed(nccl_step, nccl_step)
    ag = {"ms_per_step_with_gather": ms_g, "value_with_gather": world * BATCH / (ms_g * 1e-3),
          "bytes_per_rank_per_step": BATCH * 8, "collective": "ncclAllGather f64 on the compute stream"}
    peer, err = make_peer_group(ctx, rk, BATCH, comm)
    identical = None
    if peer is None:
        ag["fused_peer_store"] = {"unavailable": err}
    else:
        compiled.lnpost_device(d_post[0], BATCH, d_out)
        comm.allgather(d_out, BATCH, d_all)
        want = np.empty(world * BATCH)
        ctx.d2h(want, d_all)
        got = np.empty(world * BATCH)
        ctx.d2h(got, peer.lnpost(compiled, d_post[0], BATCH))
        identical = rk.all(bool(np.array_equal(got, want, equal_nan=True)))

        def peer_step():
            peer.lnpost(compiled, d_post[step_i[0] % N_BATCHES], BATCH)
            step_i[0] += 1

        ms_p = timed(peer_step, peer_step)
        peer.check()
        ag["fused_peer_store"] = {
            "ms_per_step": ms_p, "value": world * BATCH / (ms_p * 1e-3), "identical_to_nccl": identical,
            "nvlink_gb_s_out_per_rank": (world - 1) * BATCH * 8 / (ms_p * 1e-3) / 1e9,
            "how": "iso_lnpost_allgather_device: the lnpost kernel writes each row to all ranks' buffers through "
                   "CUDA-IPC peer mappings (NVLink), then a flag exchange with a bounded wait; no NCCL call in the step"}
        peer.close()
    ctx.dev_free(d_all)
    out["allgather"] = ag

    # ---- configs[4]: binary star, 1e6 rows in total, sharded by rows ----------------------------------------------
    ic, single, binary, t1, t2 = iso_world(ctx, bc)
    rows = binary_batch(BATCH, t2, 70)                  # same seed on every rank: one global batch
    sh = parallel.RowSharder(BATCH, world, rk.rank)
    mine = np.ascontiguousarray(sh.local(rows))
    d_p = stage(ctx, [mine])[0]
    d_o, d_all = ctx.dev_alloc(sh.pad * 8), ctx.dev_alloc(world * sh.pad * 8)
    n_mine = len(mine)
    k_step = lambda: binary.compiled.lnpost_device(d_p, n_mine, d_o)
    g_step = lambda: (binary.compiled.lnpost_device(d_p, n_mine, d_o), comm.allgather(d_o, sh.pad, d_all))
    ms_k = timed(k_step, g_step)
    ms_g = timed(g_step, g_step)
    got = np.empty((world, sh.pad))
    ctx.d2h(got, d_all)
    full = sh.assemble(got)
    ms_f, same = None, None
    peer, _ = make_peer_group(ctx, rk, sh.pad, comm)
    if peer is not None:
        got_f = np.empty((world, sh.pad))
        ctx.d2h(got_f, peer.lnpost(binary.compiled, d_p, n_mine))
        same = rk.all(bool(np.array_equal(sh.assemble(got_f), full, equal_nan=True)))
        f_step = lambda: peer.lnpost(binary.compiled, d_p, n_mine)
        ms_f = timed(f_step, f_step)
        peer.check()
        peer.close()
    out["binary_1e6_rows_sharded"] = {
        "value": BATCH / ((ms_f or ms_g) * 1e-3), "unit": UNIT, "ms_per_step_fused_peer_store": ms_f,
        "ms_per_step_nccl_allgather": ms_g, "ms_per_step_without_gather": ms_k, "fused_identical_to_nccl": same,
        "rows_per_gpu": sh.pad, "finite_frac": float(np.isfinite(full).mean()), "scaling": "strong",
        "config": "configs[4]: BinaryStarModel, iso grid 107x15x1710, 4 bands + parallax, 1e6 rows in total sharded over "
                  "%d GPUs; every rank receives the full lnpost vector (%d B per rank): fused peer stores, ncclAllGather "
                  "beside it" % (world, sh.pad * 8)}
    for d in (d_p, d_o, d_all):
        ctx.dev_free(d)

    # ---- configs[3] at size, stars sharded; one ensemble sharded -----------------------------------------------------
    out["catalog_10k_stars_fit"] = catalog_fit(ctx, ic, rk, comm, args)
    out["sharded_ensemble_1M_walkers"] = sharded_ensemble(ctx, single, t1, rk, args)

    # ---- host link of every rank (what bounds the end-to-end number at N >= 4) ---------------------------------------
    d_buf = ctx.dev_alloc(BATCH * 40)
    for _ in range(2):
        ctx.h2d(d_buf, h_in)
    rk.barrier()
    t0 = time.perf_counter()
    for _ in range(5):
        ctx.h2d(d_buf, h_in)
    dt_all = time.perf_counter() - t0
    rk.barrier()
    rates = [float(v) for v in rk.allgather_bytes(repr(5 * BATCH * 40 / dt_all / 1e9).encode())]
    solo = None
    for r in range(world):                  # one rank at a time: the link without its neighbours
        rk.barrier()
        if r == rk.rank:
            t0 = time.perf_counter()
            for _ in range(5):
                ctx.h2d(d_buf, h_in)
            solo = 5 * BATCH * 40 / (time.perf_counter() - t0) / 1e9
    rk.barrier()
    solos = [float(v) for v in rk.allgather_bytes(repr(solo).encode())]
    ctx.dev_free(d_buf)
    out["host_link"] = {"h2d_gb_s_per_rank_all_ranks_copying": rates, "h2d_gb_s_per_rank_alone": solos,
                        "sum_gb_s_concurrent": float(sum(rates)),
                        "note": "pinned 40 MB H2D copies; concurrent rates far below the solo rates = the ranks share the "
                                "host side of the box (PCIe switch uplinks / host memory), which is what caps e2e at N >= 4"}
    comm.close()
    failed = [k for k, v in (("allgather", identical), ("binary", same)) if v is False]
    if not out["sharded_ensemble_1M_walkers"]["copies_identical_across_ranks"]:
        failed.append("sharded_ensemble")
    if rk.rank == 0 and out["sharded_ensemble_1M_walkers"]["identical_to_single_rank_run"] is False:
        failed.append("sharded_ensemble_vs_single")
    return out, failed


# ---------------------------------------------------------------------------------------------------------------------
def roofline_block(info, clock_summary, peaks, peak, peak_src, kernel_ms, alt):
    """The bound that binds the headline batch, with the algorithmic (HBM) figure and the HBM-bound batch next to it."""
    # every gathered record is a set of scattered 32-byte sectors and an SM's L1 data pipe retires one such sector per
    # clock whatever level of the hierarchy holds the line: 4 corner pairs x 3.5 sectors on average (48-byte nodes: a
    # pair of EEP-adjacent nodes is 96 bytes at a 16-byte-aligned offset = 3 or 4 sectors) + 16 BC corners x 1 sector
    sectors = 4 * 3.5 + 16 * 1
    sm_hz = float(clock_summary.get("sm_mhz") or peaks.get("sm_max_mhz") or 1965.0) * 1e6
    l1_peak = info["sm_count"] * sm_hz * 32 / 1e9
    l1_ach = sectors * 32 * BATCH / (kernel_ms * 1e-3) / 1e9
    traffic, counters, stale = {}, {}, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            with open(tpath) as f:
                tj = json.load(f)
            stale = tj.get("kernel_source_hash") != kernel_source_hash()
            if not stale:
                traffic = {k: v.get("dram_bytes_per_launch") for k, v in tj.get("batches", {}).items()}
                counters = {k: v.get("counters") for k, v in tj.get("batches", {}).items()}
        except Exception:
            pass
    achieved = B_ALG * BATCH / (kernel_ms * 1e-3) / 1e9
    gw = alt.get("grid_wide", {})
    gw_ms = gw.get("ms_per_step")
    block = {
        "bound": "l1_sector", "achieved": l1_ach, "peak": l1_peak, "unit": "GB/s", "frac": l1_ach / l1_peak,
        "traffic": traffic.get("posterior_like"),
        "kernel": "iso_lnpost_kernel<1 star, single model, default priors, track grid>", "kernel_ms": kernel_ms,
        "peak_source": "148 SMs x SM clock x one 32-byte sector per clock (L1 data-pipe wavefront rate, B300_MICROARCH.md); "
                       "clock = median sampled during the timed loop",
        "sectors_per_eval": sectors, "counters": counters.get("posterior_like"),
        "traffic_source": "profiles/traffic.json (ncu --set full of this kernel source)" if traffic else
                          ("stale: profiles/traffic.json was captured from other kernel sources" if stale else "no capture"),
        "note": "posterior-like rows are served by L1 / L2 (DRAM sees only the row stream: `traffic`), so the kernel is "
                "bounded by the rate at which an SM retires scattered sectors, not by HBM; `algorithmic` is the "
                "SURVEY.md §8d figure (944 B per evaluation) against the measured HBM copy peak for reference",
        "algorithmic": {"bytes_per_eval": B_ALG, "achieved_gb_s": achieved, "hbm_peak_gb_s": peak, "frac_of_hbm_peak": achieved / peak,
                        "peak_source": peak_src},
    }
    if gw_ms:
        gw_ach = B_ALG * BATCH / (gw_ms * 1e-3) / 1e9
        t_gw = traffic.get("grid_wide")
        block["hbm_bound_batch"] = {
            "workload": "grid_wide (gathers over the whole track grid, more than the L2 holds)", "bound": "hbm",
            "achieved": gw_ach, "peak": peak, "unit": "GB/s", "frac": gw_ach / peak, "kernel_ms": gw_ms, "traffic": t_gw,
            "dram_gb_s": (t_gw / (gw_ms * 1e-3) / 1e9) if t_gw else None,
            "dram_frac_of_peak": (t_gw / (gw_ms * 1e-3) / 1e9 / peak) if t_gw else None,
            "counters": counters.get("grid_wide"),
            "note": "achieved = 944 algorithmic bytes x rows / kernel time; traffic = measured DRAM bytes of THIS batch per "
                    "launch (less than the algorithmic bytes: the 48-byte-node grid partly stays in the L2)"}
    return block


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the 'alt' workloads beyond the three track-grid distributions")
    ap.add_argument("--small", action="store_true", help="small grids (debugging only; not a valid bench)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        run_reference(args, int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")))
        return

    rk = Ranks()
    rank, world, local_rank = rk.rank, rk.world, rk.local_rank
    from isochrones_b200 import _lib, parallel, synthetic as syn

    numa_cpus = parallel.bind_to_gpu_numa(local_rank) if world > 1 else None   # before any pinned allocation
    ctx = _lib.default_context(local_rank)
    trk, bc, ic, truth, n_eep = build_workload(ctx=ctx, small=args.small)
    mags = truth_mags(ic, truth)
    mod = make_model(ic, mags)
    compiled = mod.compiled
    bounds = [mod.bounds(p) for p in mod.param_names]

    # ---- device-resident batches: this rank's shard of every step --------------------------------------------
    post = [syn.posterior_like_batch("track", BATCH, truth, n_eep=n_eep, seed=2 + 1000 * rank + s) for s in range(N_BATCHES)]
    d_post = stage(ctx, post)
    d_out = ctx.dev_alloc(BATCH * 8)

    clocks = ClockSampler(local_rank)
    tw0 = time.time()
    ms, launches = timed_device_loop(ctx, compiled, d_post, d_out, args.steps, args.warmup, barrier=rk.barrier)
    clocks.mark(tw0, time.time())
    ms = rk.max(ms)
    value = world * BATCH * args.steps / (ms * 1e-3)
    kernel_ms = ms / args.steps

    # ---- end to end through the public API: pinned host buffers, H2D + kernel + D2H inside the timed region ----
    h_in = [ctx.pinned_empty((BATCH, 5)) for _ in range(2)]
    h_out = ctx.pinned_empty((BATCH,))
    for i in range(2):
        h_in[i][:] = post[i]
    for s in range(max(args.warmup, 3)):
        mod.lnpost_batch(h_in[s % 2], out=h_out)
    rk.barrier()
    tw0 = time.time()
    t0 = time.perf_counter()
    for s in range(args.steps):
        mod.lnpost_batch(h_in[s % 2], out=h_out)
    e2e_s = rk.max(time.perf_counter() - t0)
    clocks.mark(tw0, time.time())
    rk.barrier()
    e2e_value = world * BATCH * args.steps / e2e_s
    finite_frac = float(np.isfinite(h_out).mean())
    # the same call shape when the rows are MADE on the device (prior-box draws: nothing shipped per row, 8 B back)
    for s in range(3):
        mod.prior_box_draws(BATCH, seed=99, row0=(rank * 64 + s) * BATCH, return_pars=False, out=h_out)
    rk.barrier()
    t0 = time.perf_counter()
    for s in range(args.steps):
        mod.prior_box_draws(BATCH, seed=99, row0=(rank * 64 + s) * BATCH, return_pars=False, out=h_out)
    draws_s = rk.max(time.perf_counter() - t0)
    rk.barrier()
    clock_summary = clocks.summary()

    multi, failed = None, []
    if world > 1:
        multi, failed = multi_gpu(ctx, compiled, d_post, d_out, bc, args, rk, h_in[0], h_out, mod)
    rk.close()      # the ranks part here: what follows is rank 0's own (single-GPU) reporting work
    if rank != 0:
        if failed:
            sys.exit(3)
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
        alt[name] = device_workload(ctx, compiled, gen, args, peak, B_ALG, n_batches=N_BATCHES)
    if not args.no_extras and not args.small:
        alt["eleven_bands_posterior_like"] = eleven_band_workload(ctx, trk, truth, n_eep, args, peak)
        if world == 1:
            alt.update(interp_workloads(ctx, ic, truth, n_eep, args, mod))
        alt.update(extra_workloads(ctx, bc, args, peak, rk))

    info = ctx.info()
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": kernel_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "StarModel lnpost batch 1e6 (configs[1]): single Sun-like star, Teff/logg/feh + VJHK + "
                               "parallax, MIST-shaped track grid 15x196x1710 (48-byte nodes: %d MB in HBM), BC 70x26x18x13"
                               % (trk["grid"].shape[0] * trk["grid"].shape[1] * trk["grid"].shape[2] * 48 // 2 ** 20),
                   "batch_rows": BATCH, "rows_per_gpu_per_step": BATCH, "distribution": "posterior-like (sigma: mass .05, "
                   "eep 15, feh .1, d 2 pc, AV .05)", "l2": "inputs larger than L2: steps rotate over %d distinct 1e6-row "
                   "batches (%d MB)" % (N_BATCHES, N_BATCHES * BATCH * 40 // 2 ** 20), "finite_frac": finite_frac,
                   "sharding": "rows sharded across ranks, no data-path collective" if world > 1 else "single GPU",
                   "rank_plumbing": "files (isochrones_b200.parallel.FileRendezvous); no torch" if world > 1 else "none",
                   "host_numa_binding": ("rank 0 bound to %d GPU-local cores" % len(numa_cpus)) if numa_cpus else "none",
                   "kernel_source_hash": kernel_source_hash()},
        "clocks": clock_summary,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": BATCH * 40, "d2h_bytes_per_step": BATCH * 8,
                "api": "BasicStarModel.lnpost_batch(pinned host array) -> iso_lnpost_batch (C ABI), chunked "
                       "H2D/kernel/D2H pipeline on two streams",
                "device_made_rows": {"value": world * BATCH * args.steps / draws_s, "unit": UNIT, "h2d_bytes_per_step": 0,
                                     "d2h_bytes_per_step": BATCH * 8,
                                     "api": "BasicStarModel.prior_box_draws(n, seed) -> iso_lnpost_prior_draws: rows drawn by "
                                            "Philox on the device (prior-like: most rows end at the grid-free priors), only lnpost "
                                            "returns"}},
        "gpu_launches": launches,
        "alt": alt,
    }
    line["roofline"] = roofline_block(info, clock_summary, peaks, peak, peak_src, kernel_ms, alt)
    if multi:
        line["multi_gpu"] = multi
    if failed:
        line["consistency_failures"] = failed
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(trk, bc, mod, seed=900)
        line["cpu_baseline"]["reference_as_shipped"] = reference_as_shipped()
    print(json.dumps(line), flush=True)
    if failed:
        sys.exit(3)


if __name__ == "__main__":
    main()
