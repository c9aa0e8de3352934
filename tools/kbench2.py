#!/usr/bin/env python
"""Kernel-only timing of the fused lnpost kernel for tuning variants (development tool, not the bench):

    [ISO_B200_LIB=path/to/variant.so] python tools/kbench2.py [--steps K] [--only a,b] [--rows N]

Workloads: the four single-star distributions of bench.py (posterior-like, prior-like, scattered_valid, grid_wide)
on the track grid, the binary star on the isochrone grid, and the single star on the isochrone grid.  Each line also
prints a checksum of the finite results (sum and count) so that variants can be compared for identical outputs."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--rows", type=int, default=bench.BATCH)
    ap.add_argument("--only", default="")
    ap.add_argument("--tag", default=os.path.basename(os.environ.get("ISO_B200_LIB", "default")))
    args = ap.parse_args()
    import isochrones_b200 as ib
    from isochrones_b200 import _lib, synthetic as syn

    only = set(args.only.split(",")) if args.only else None
    ctx = _lib.default_context(0)
    trk, bc, ic, truth, n_eep = bench.build_workload(ctx=ctx)
    mags = bench.truth_mags(ic, truth)
    mod = bench.make_model(ic, mags)
    bounds = [mod.bounds(p) for p in mod.param_names]
    n = args.rows
    work = [
        ("posterior", mod, lambda s: syn.posterior_like_batch("track", n, truth, n_eep=n_eep, seed=2 + s)),
        ("prior", mod, lambda s: syn.prior_like_batch("track", n, bounds, seed=3 + s)),
        ("scattered", mod, lambda s: bench.scattered_batch(n, seed=50 + s)),
        ("grid_wide", mod, lambda s: bench.grid_wide_batch(n, trk, seed=90 + s)),
    ]
    if only is None or only & {"binary", "iso_single", "iso_prior"}:
        iso = syn.make_iso_grid(columns=("Teff", "logg", "feh", "Mbol", "mass", "dm_deep", "nu_max", "delta_nu"))
        ici = ib.ichrone_from_arrays("iso", iso, bc, ctx=ctx)
        t2 = syn.default_truth("iso", n_stars=2)
        _, _, _, m2 = ici.interp_mag([t2[0]] + list(t2[2:]), list(bench.BANDS))
        obs = {b: (float(np.round(m, 3)) - 0.35, 0.02) for b, m in zip(bench.BANDS, m2)}
        binary = ib.BinaryStarModel(ici, Teff=(5772.0, 80.0), logg=(4.44, 0.1), feh=(0.0, 0.1), parallax=(10.0, 0.1), **obs)

        def bin_batch(s):
            p = syn.posterior_like_batch("iso", n, t2, seed=70 + s)
            p[:, :2] = -np.sort(-p[:, :2], axis=1)
            return p

        t1 = syn.default_truth("iso", n_stars=1)
        _, _, _, m1 = ici.interp_mag(list(t1), list(bench.BANDS))
        single = ib.SingleStarModel(ici, Teff=(5772.0, 80.0), logg=(4.44, 0.1), feh=(0.0, 0.1), parallax=(10.0, 0.1),
                                    **{b: (float(np.round(m, 3)), 0.02) for b, m in zip(bench.BANDS, m1)})
        b1 = [single.bounds(p) for p in single.param_names]
        work += [("binary", binary, bin_batch),
                 ("iso_single", single, lambda s: syn.posterior_like_batch("iso", n, t1, seed=170 + s)),
                 ("iso_prior", single, lambda s: syn.prior_like_batch("iso", n, b1, seed=270 + s))]
    d_out = ctx.dev_alloc(n * 8)
    out = np.empty(n)
    parts = []
    for name, model, gen in work:
        if only is not None and name not in only:
            continue
        compiled = model.compiled
        ptrs = []
        for s in range(4):
            b = np.ascontiguousarray(gen(s))
            d = ctx.dev_alloc(b.nbytes)
            ctx.h2d(d, b)
            ptrs.append(d)
        for s in range(5):
            compiled.lnpost_device(ptrs[s % 4], n, d_out)
        ctx.sync()
        ctx.timer_start()
        for s in range(args.steps):
            compiled.lnpost_device(ptrs[s % 4], n, d_out)
        ms = ctx.timer_stop() / args.steps
        compiled.lnpost_device(ptrs[0], n, d_out)
        ctx.d2h(out, d_out)
        fin = np.isfinite(out)
        parts.append("%s %.4f ms %.2fe9/s [%d %.9e %d]" % (name, ms, n / ms / 1e6, int(fin.sum()), float(out[fin].sum()),
                                                         int(np.isnan(out).sum())))
        for d in ptrs:
            ctx.dev_free(d)
    print("%-28s " % args.tag + " | ".join(parts), flush=True)


if __name__ == "__main__":
    main()
