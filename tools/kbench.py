#!/usr/bin/env python
"""Kernel-only timing of the fused lnpost kernel on the bench workload (development tool, not the bench):
    [ISO_B200_LIB=path/to/variant.so] python tools/kbench.py [--steps K] [--models track,iso,binary]"""
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
    ap.add_argument("--tag", default=os.environ.get("ISO_B200_LIB", "default"))
    args = ap.parse_args()
    from isochrones_b200 import _lib, synthetic as syn

    ctx = _lib.default_context(0)
    trk, bc, ic, truth, n_eep = bench.build_workload(ctx=ctx)
    mags = bench.truth_mags(ic, truth)
    mod = bench.make_model(ic, mags)
    compiled = mod.compiled
    bounds = [mod.bounds(p) for p in mod.param_names]
    n = args.rows
    gens = {
        "posterior": lambda s: syn.posterior_like_batch("track", n, truth, n_eep=n_eep, seed=2 + s),
        "prior": lambda s: syn.prior_like_batch("track", n, bounds, seed=3 + s),
        "scattered": lambda s: bench.scattered_batch(n, seed=50 + s),
        "grid_wide": lambda s: bench.grid_wide_batch(n, trk, seed=90 + s),
        # bound experiments: every row identical (all gathers broadcast: the L1 data pipe costs nothing) and rows
        # sharing one cell (distinct weights, same 24 nodes)
        "identical": lambda s: np.tile(truth, (n, 1)),
        # groups of 2 / 4 / 8 consecutive rows identical: every gather instruction touches 16 / 8 / 4 distinct lines
        # instead of 32 — the L1 wavefront count lane-cooperative loads would have, without their overhead
        "pair_same": lambda s: np.repeat(syn.posterior_like_batch("track", n // 2, truth, n_eep=n_eep, seed=2 + s), 2, axis=0),
        "quad_same": lambda s: np.repeat(syn.posterior_like_batch("track", n // 4, truth, n_eep=n_eep, seed=2 + s), 4, axis=0),
        "oct_same": lambda s: np.repeat(syn.posterior_like_batch("track", n // 8, truth, n_eep=n_eep, seed=2 + s), 8, axis=0),
        "one_cell": lambda s: truth + np.array([0.004, 0.4, 0.01, 2.0, 0.02]) * np.random.RandomState(s).random_sample((n, 5)),
    }
    d_out = ctx.dev_alloc(n * 8)
    res = {}
    for name, gen in gens.items():
        if args.only and name not in args.only.split(","):
            continue
        ptrs = []
        for s in range(4):
            b = gen(s)
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
        res[name] = ms
        for d in ptrs:
            ctx.dev_free(d)
    print("%-40s " % os.path.basename(args.tag) + "  ".join("%s %.4f ms (%.2fe9/s)" % (k, v, n / v / 1e6) for k, v in res.items()),
          flush=True)


if __name__ == "__main__":
    main()
