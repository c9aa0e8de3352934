#!/usr/bin/env python
"""End-to-end (pinned host -> lnpost -> pinned host) timing of BasicStarModel.lnpost_batch for pipeline tuning:
    ISO_PIPE_CHUNK_ROWS=131072 python tools/e2e_bench.py"""
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
    mags = bench.truth_mags(ic, truth)
    mod = bench.make_model(ic, mags)
    n = bench.BATCH
    h_in = [ctx.pinned_empty((n, 5)) for _ in range(2)]
    h_out = ctx.pinned_empty((n,))
    for i in range(2):
        h_in[i][:] = syn.posterior_like_batch("track", n, truth, n_eep=n_eep, seed=2 + i)
    for s in range(5):
        mod.lnpost_batch(h_in[s % 2], out=h_out)
    best = 1e9
    for rep in range(3):
        t0 = time.perf_counter()
        for s in range(20):
            mod.lnpost_batch(h_in[s % 2], out=h_out)
        best = min(best, (time.perf_counter() - t0) / 20)
    pageable = syn.posterior_like_batch("track", n, truth, n_eep=n_eep, seed=9)
    for s in range(3):
        mod.lnpost_batch(pageable)
    t0 = time.perf_counter()
    for s in range(10):
        mod.lnpost_batch(pageable)
    pg = (time.perf_counter() - t0) / 10
    print("chunk %-8s pinned %.4f ms (%.3fe9/s, %.1f GB/s)   pageable %.4f ms (%.3fe9/s)" % (
        os.environ.get("ISO_PIPE_CHUNK_ROWS", "default"), best * 1e3, n / best / 1e9, 48e-3 * n / 1e6 / best / 1e0 / 1e0 * 1e-3 * 1e3,
        pg * 1e3, n / pg / 1e9), flush=True)


if __name__ == "__main__":
    main()
