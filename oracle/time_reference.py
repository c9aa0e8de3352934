#!/usr/bin/env python
"""TEST INFRASTRUCTURE — times the UNMODIFIED reference (pure Python + numba) on the bench workload.

    python oracle/time_reference.py                         # build container: writes profiles/reference_python_timing.json
    python oracle/time_reference.py --root DIR --pool --json # any box: reference sources under DIR/isochrones (the archive
                                                             # oracle/build_ref.py makes travels to the GPU box under
                                                             # oracle/_ref/), one JSON line on stdout; `--pool` adds the
                                                             # reference's own batch recipe: the scalar loop fanned out over
                                                             # a multiprocessing.Pool of all cores
                                                             # (notebooks/batch-demo.ipynb:122-124)

What is timed (BASELINE.json configs[1] shapes: full MIST-shaped track grid + BC grid, Sun-like star, VJHK + parallax):
  * `BasicStarModel.lnpost(p)` — the call emcee / MultiNest make — over posterior-like rows, one Python call per row;
  * `ModelGridInterpolator.interp_mag` on arrays (numba `interp_mags`, the reference's own batched kernel);
  * `DFInterpolator.__call__` on arrays (numba `interp_values_3d`), 3 properties.
Single thread (the reference has no threaded path).  The numbers are container-CPU numbers: context for the GPU box's
`cpu_baseline`, which runs the C port of the same arithmetic on all host threads.
"""
import json
import os
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

_MOD = None


def _pool_chunk(rows):
    """Worker of the process pool (forked after the model was built: every worker holds its own copy, as the
    reference's per-star process pools do)."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return [_MOD.lnpost(r) for r in rows]


def main():
    import argparse

    ap = argparse.ArgumentParser()
    ap.add_argument("--root", default=None, help="directory that holds the reference's `isochrones` package")
    ap.add_argument("--pool", action="store_true")
    ap.add_argument("--json", action="store_true", help="print one JSON line instead of writing profiles/")
    ap.add_argument("--calls", type=int, default=5000)
    args = ap.parse_args()
    if args.root:
        os.environ["ISOCHRONES_REFERENCE"] = args.root
    import bench  # noqa: E402
    from isochrones_b200 import synthetic as syn  # noqa: E402
    from oracle import ref_shim  # noqa: E402

    global _MOD
    ref = ref_shim.load()
    trk = syn.make_track_grid(columns=bench.PACK_COLUMNS)
    bc = syn.make_bc_grid(bands=bench.BANDS)
    ic = ref_shim.make_ref_ic("track", trk, bc, eep_bounds=(0, 1710))
    truth = syn.default_truth("track")
    _, _, _, mags = ic.interp_mag(list(truth), list(bench.BANDS))
    obs = {b: (float(np.round(m, 3)), 0.02) for b, m in zip(bench.BANDS, mags)}
    mod = ref.starmodel.BasicStarModel(ic, Teff=(5772.0, 80.0), logg=(4.44, 0.1), feh=(0.0, 0.1), parallax=(10.0, 0.1), **obs)
    _MOD = mod
    rows = syn.posterior_like_batch("track", 200_000, truth, seed=2)
    out = {"where": "build container (not the GPU box)" if not args.root else "this box", "threads": 1,
           "grid": "track 15x196x1710, BC 70x26x18x13"}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for r in rows[:200]:
            mod.lnpost(r)                                  # numba compilation + warm-up
        n = args.calls
        t0 = time.perf_counter()
        for r in rows[:n]:
            mod.lnpost(r)
        dt = time.perf_counter() - t0
    out["lnpost_scalar"] = {"calls": n, "us_per_call": dt / n * 1e6, "evals_per_s": n / dt,
                            "what": "reference BasicStarModel.lnpost(p), one Python call per row (starmodel.py:538-542)"}
    if args.pool:
        import multiprocessing as mp

        try:
            cores = max(1, len(os.sched_getaffinity(0)))
        except AttributeError:
            cores = os.cpu_count() or 1
        per = max(200, args.calls // 2)
        chunks = [rows[1000 + i * per:1000 + (i + 1) * per] for i in range(cores)]
        with mp.get_context("fork").Pool(cores) as pool:
            pool.map(_pool_chunk, [c[:20] for c in chunks])              # warm the workers
            t0 = time.perf_counter()
            res = pool.map(_pool_chunk, chunks)
            dt = time.perf_counter() - t0
        out["lnpost_scalar_pool"] = {"calls": cores * per, "processes": cores, "evals_per_s": cores * per / dt,
                                     "finite": int(np.isfinite(np.concatenate(res)).sum()),
                                     "what": "the same scalar loop fanned out over multiprocessing.Pool(all cores) — the "
                                             "reference's own batch recipe (notebooks/batch-demo.ipynb:122-124)"}
    pars = [np.ascontiguousarray(rows[:, j]) for j in range(5)]
    ic.interp_mag(pars, list(bench.BANDS))
    t0 = time.perf_counter()
    ic.interp_mag(pars, list(bench.BANDS))
    dt = time.perf_counter() - t0
    out["interp_mag_arrays"] = {"points": len(rows), "us_per_point": dt / len(rows) * 1e6, "points_per_s": len(rows) / dt,
                                "what": "reference ModelGridInterpolator.interp_mag on arrays -> numba interp_mags (mags.py:64-124)"}
    interp = ic.model_grid.interp
    xx = [pars[2], pars[0], pars[1]]
    interp(xx, ["Teff", "logg", "nu_max"])
    t0 = time.perf_counter()
    interp(xx, ["Teff", "logg", "nu_max"])
    dt = time.perf_counter() - t0
    out["interp_value_arrays"] = {"points": len(rows), "us_per_point": dt / len(rows) * 1e6, "points_per_s": len(rows) / dt,
                                  "what": "reference DFInterpolator.__call__ on arrays -> numba interp_values_3d (interp.py:359-374), 3 columns"}
    if args.json:
        print(json.dumps(out), flush=True)
        return
    path = os.path.join(ROOT, "profiles", "reference_python_timing.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
