#!/usr/bin/env python
"""bench.py's nested-sampling workload alone (development tool)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    from isochrones_b200 import _lib

    ctx = _lib.default_context(0)
    trk, bc, ic, truth, n_eep = bench.build_workload(ctx=ctx)
    ici, single, binary, t1, t2 = bench.iso_world(ctx, bc)
    print("truth", t1, "lnpost(truth)", single.lnpost(t1))
    for multi in (True, False):
        t0 = time.perf_counter()
        res = single.fit_nested(n_live_points=1000, seed=1, max_batches=4000, multi=multi)
        dt = time.perf_counter() - t0
        print(json.dumps({"multi": multi, "seconds": dt, "n_evals": res.n_evals, "n_iter": res.n_iter, "logZ": res.logZ,
                          "logZ_err": res.logZ_err, "H": res.information, "eff": res.efficiency, "converged": res.converged,
                          "n_batches": res.n_batches, "nell": res.n_ellipsoids_max, "finite": res.finite_fraction,
                          "mean": [float(v) for v in res.mean()], "std": [float(v) for v in res.std()],
                          "lnpost_max": float(np.max(res.lnpost))}), flush=True)


if __name__ == "__main__":
    main()
