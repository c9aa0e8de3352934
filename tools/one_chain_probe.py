#!/usr/bin/env python
"""Why does the 256 x 2000 single-chain run take 26 ms in one place and 40 ms in another?  (development probe)"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def clock():
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(0)
        return pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0
    except Exception as e:
        return repr(e)


def main():
    from isochrones_b200 import _lib, synthetic as syn
    from isochrones_b200.sampler import DeviceEnsembleSampler

    ctx = _lib.default_context(0)
    trk, bc, ic, truth, n_eep = bench.build_workload(ctx=ctx)
    ici, single, binary, t1, t2 = bench.iso_world(ctx, bc)
    nw, nsteps = 256, 2000
    p0 = syn.posterior_like_batch("iso", nw, t1, seed=4)
    p0[~np.isfinite(single.lnpost_batch(p0))] = t1
    smp = DeviceEnsembleSampler(single.compiled, nw, p0, seed=4)
    smp.run_mcmc(50, store=False)

    def runs(label, n=5, **kw):
        ts = []
        for _ in range(n):
            t0 = time.perf_counter()
            ctx.timer_start()
            smp.run_mcmc(nsteps, **kw)
            dev = ctx.timer_stop()
            ts.append((time.perf_counter() - t0, dev))
        print(label, " ".join("%.1f/%.1f" % (a * 1e3, b) for a, b in ts), "ms wall/device; clock", clock(), flush=True)

    print("clock at start", clock())
    runs("back to back, thin=10 store", thin=10)
    runs("back to back, no store    ", store=False, fetch=False)
    time.sleep(3.0)
    print("after 3 s idle", clock())
    runs("after idle, thin=10 store ", thin=10)
    rows = syn.posterior_like_batch("iso", 1_000_000, t1, seed=5)
    for _ in range(30):
        single.lnpost_batch(rows)
    print("after 30 big batches", clock())
    runs("after heavy work          ", thin=10)


if __name__ == "__main__":
    main()
