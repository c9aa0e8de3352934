#!/usr/bin/env python
"""End-to-end lnpost_batch on pinned host arrays for the current ISO_PIPE_CHUNK_ROWS (development probe)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    from isochrones_b200 import _lib, synthetic as syn

    ctx = _lib.default_context(0)
    trk, bc, ic, truth, n_eep = bench.build_workload(ctx=ctx)
    mod = bench.make_model(ic, bench.truth_mags(ic, truth))
    n = bench.BATCH
    h_in = [ctx.pinned_empty((n, 5)) for _ in range(2)]
    h_out = ctx.pinned_empty((n,))
    for i in range(2):
        h_in[i][:] = syn.posterior_like_batch("track", n, truth, n_eep=n_eep, seed=2 + i)
    for s in range(5):
        mod.lnpost_batch(h_in[s % 2], out=h_out)
    best = []
    for rep in range(3):
        t0 = time.perf_counter()
        for s in range(20):
            mod.lnpost_batch(h_in[s % 2], out=h_out)
        best.append((time.perf_counter() - t0) / 20)
    # plain copies for reference
    d = ctx.dev_alloc(n * 40)
    ctx.h2d(d, h_in[0])
    t0 = time.perf_counter()
    for s in range(20):
        ctx.h2d(d, h_in[s % 2])
    h2d = (time.perf_counter() - t0) / 20
    print("chunk %s: e2e %s ms (%.3e rows/s); bare pinned H2D of the batch %.3f ms (%.1f GB/s)" % (
        os.environ.get("ISO_PIPE_CHUNK_ROWS", "default"), " ".join("%.3f" % (b * 1e3) for b in best), n / min(best), h2d * 1e3,
        n * 40 / h2d / 1e9), flush=True)


if __name__ == "__main__":
    main()
