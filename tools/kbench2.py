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
    if only is not None and only & {"posterior_sorted", "grid_wide_sorted", "prior_sorted", "scattered_sorted", "posterior_warpsorted"}:
        ax = trk["axes"]                                   # (feh, mass, eep) axes of the track grid

        def cell_sorted(gen, block=None):
            def make(s):
                b = gen(s)
                key = (np.searchsorted(ax[0], b[:, 2]) * 4096 + np.searchsorted(ax[1], b[:, 0])) * 4096 + np.searchsorted(ax[2], b[:, 1])
                if block is None:
                    return b[np.argsort(key, kind="stable")]
                out = b.copy()                             # sorted inside blocks of `block` rows only (what a CTA could do)
                for i in range(0, len(b), block):
                    out[i:i + block] = b[i:i + block][np.argsort(key[i:i + block], kind="stable")]
                return out
            return make
        base = {n_: g_ for n_, _, g_ in work}
        work += [("posterior_sorted", mod, cell_sorted(base["posterior"])), ("grid_wide_sorted", mod, cell_sorted(base["grid_wide"])),
                 ("prior_sorted", mod, cell_sorted(base["prior"])), ("scattered_sorted", mod, cell_sorted(base["scattered"])),
                 ("posterior_warpsorted", mod, cell_sorted(base["posterior"], block=8192))]
    extra = {"binary", "binary8", "binary11", "iso_single", "iso_prior", "catalog", "chains", "one_chain"}
    if only is None or only & extra:
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
        if only is not None and only & {"binary8", "binary11"}:
            many = ("J", "H", "K", "G", "BP", "RP", "W1", "W2", "W3", "TESS", "Kepler")
            for nb in (8, 11):
                bcn = syn.make_bc_grid(bands=many[:nb])
                icn = ib.ichrone_from_arrays("iso", iso, bcn, ctx=ctx)
                _, _, _, mn = icn.interp_mag([t2[0]] + list(t2[2:]), list(many[:nb]))
                obsn = {b: (float(np.round(m, 3)) - 0.35, 0.02) for b, m in zip(many[:nb], mn)}
                work.append(("binary%d" % nb, ib.BinaryStarModel(icn, Teff=(5772.0, 80.0), logg=(4.44, 0.1), feh=(0.0, 0.1),
                                                                parallax=(10.0, 0.1), **obsn), bin_batch))
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
    if only is None or only & {"catalog", "chains", "one_chain"}:
        parts += sampler_and_catalog(ctx, ici, single, t1, only, args)
    print("%-28s " % args.tag + " | ".join(parts), flush=True)


def sampler_and_catalog(ctx, ic, single, truth1, only, args):
    """bench.py's catalog (10 000 star models, 100 rows each through model_of_row) and on-device sampler workloads."""
    import time

    import pandas as pd

    from isochrones_b200 import synthetic as syn
    from isochrones_b200.catalog import StarCatalog
    from isochrones_b200.sampler import DeviceEnsembleSampler

    parts = []
    if only is None or "catalog" in only:
        n_stars, rows_per_star = 10_000, 100
        rng = np.random.RandomState(8)
        t = np.tile(truth1, (n_stars, 1))
        t[:, 0] = rng.uniform(300.0, 900.0, n_stars)
        t[:, 1] = rng.uniform(9.0, 9.9, n_stars)
        t[:, 2] = rng.uniform(-0.5, 0.3, n_stars)
        t[:, 3] = rng.uniform(50.0, 400.0, n_stars)
        t[:, 4] = rng.uniform(0.0, 0.5, n_stars)
        _, _, _, mg = ic.interp_mag([t[:, j] for j in range(5)], list(bench.BANDS))
        table = {"parallax": 1000.0 / t[:, 3], "parallax_unc": np.full(n_stars, 0.1)}
        for j, b in enumerate(bench.BANDS):
            table[b + "_mag"] = mg[:, j]
            table[b + "_mag_unc"] = np.full(n_stars, 0.02)
        compiled = StarCatalog(pd.DataFrame(table), props=["parallax"]).compile(ic)
        mor = np.repeat(np.arange(n_stars, dtype=np.int32), rows_per_star)
        pars = np.repeat(t, rows_per_star, axis=0)
        pars *= 1 + 0.002 * rng.standard_normal(pars.shape)
        perm0 = np.random.RandomState(9).permutation(len(pars))
        # the same catalog with the stars spread over the whole populated isochrone grid (what a real catalog does)
        tw = t.copy()
        tw[:, 0] = rng.uniform(210.0, 1400.0, n_stars)
        tw[:, 1] = rng.uniform(7.5, 10.1, n_stars)
        tw[:, 2] = rng.uniform(-3.5, 0.45, n_stars)
        _, _, _, mgw = ic.interp_mag([tw[:, j] for j in range(5)], list(bench.BANDS))
        bad = ~np.isfinite(mgw).all(axis=1)
        tw[bad], mgw[bad] = t[bad], mg[bad]
        tablew = {"parallax": 1000.0 / tw[:, 3], "parallax_unc": np.full(n_stars, 0.1)}
        for j, b in enumerate(bench.BANDS):
            tablew[b + "_mag"] = mgw[:, j]
            tablew[b + "_mag_unc"] = np.full(n_stars, 0.02)
        compiled_w = StarCatalog(pd.DataFrame(tablew), props=["parallax"]).compile(ic)
        parsw = np.repeat(tw, rows_per_star, axis=0)
        parsw *= 1 + 0.002 * rng.standard_normal(parsw.shape)
        permw = np.random.RandomState(10).permutation(len(parsw))
        cases = (("catalog", compiled, pars, mor), ("catalog_shuffled", compiled, pars[perm0], mor[perm0]),
                 ("catalog_wide", compiled_w, parsw, mor), ("catalog_wide_shuffled", compiled_w, parsw[permw], mor[permw]))
        for tag, compiled, pp, mm in cases:
            pp, mm = np.ascontiguousarray(pp), np.ascontiguousarray(mm)
            n_rows = len(pp)
            d_p, d_m, d_o = ctx.dev_alloc(pp.nbytes), ctx.dev_alloc(mm.nbytes), ctx.dev_alloc(n_rows * 8)
            ctx.h2d(d_p, pp)
            ctx.h2d(d_m, mm)
            for _ in range(3):
                compiled.lnpost_device(d_p, n_rows, d_o, d_model_of_row=d_m)
            ctx.sync()
            ctx.timer_start()
            for _ in range(args.steps):
                compiled.lnpost_device(d_p, n_rows, d_o, d_model_of_row=d_m)
            ms = ctx.timer_stop() / args.steps
            res = np.empty(n_rows)
            ctx.d2h(res, d_o)
            fin = np.isfinite(res)
            parts.append("%s %.4f ms %.2fe9/s [%d %.9e]" % (tag, ms, n_rows / ms / 1e6, int(fin.sum()), float(res[fin].sum())))
            for d in (d_p, d_m, d_o):
                ctx.dev_free(d)
    nw = 256
    if only is None or "one_chain" in only:
        p0 = syn.posterior_like_batch("iso", nw, truth1, seed=4)
        smp = DeviceEnsembleSampler(single.compiled, nw, p0, seed=4)
        smp.run_mcmc(50, store=False, fetch=False)
        ts = []
        for _ in range(5):
            t0 = time.perf_counter()
            smp.run_mcmc(2000, store=False, fetch=False)
            ts.append(time.perf_counter() - t0)
        parts.append("one_chain 256x2000 min %.2f med %.2f max %.2f ms" % (1e3 * min(ts), 1e3 * sorted(ts)[2], 1e3 * max(ts)))
        smp.close()
    if only is None or "chains" in only:
        n_chains, steps_c = 1184, 100
        p0c = np.stack([syn.posterior_like_batch("iso", nw, truth1, seed=1000 + c) for c in range(8)])
        p0c = np.ascontiguousarray(np.tile(p0c, (n_chains // 8, 1, 1)))
        smc = DeviceEnsembleSampler(single.compiled, nw, p0c, seed=5, n_chains=n_chains)
        smc.run_mcmc(10, store=False, fetch=False)
        ts = []
        for _ in range(3):
            t0 = time.perf_counter()
            smc.run_mcmc(steps_c, store=False, fetch=False)
            ts.append(time.perf_counter() - t0)
        parts.append("chains 1184x256x100 %.2f ms %.2fe9/s" % (1e3 * min(ts), n_chains * nw * steps_c / min(ts) / 1e9))
        smc.close()
    return parts


if __name__ == "__main__":
    main()
