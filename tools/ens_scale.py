#!/usr/bin/env python
"""Only the sharded-ensemble workload of bench.py (development tool): run under torchrun with N ranks, prints one JSON line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import argparse

    ap = argparse.ArgumentParser()
    ap.add_argument("--walkers", type=int, default=1 << 20)
    ap.add_argument("--steps", type=int, default=50)
    args = ap.parse_args()
    rk = bench.Ranks()
    from isochrones_b200 import _lib, parallel

    if rk.world > 1:
        parallel.bind_to_gpu_numa(rk.local_rank)
    ctx = _lib.default_context(rk.local_rank)
    trk, bc, ic, truth, n_eep = bench.build_workload(ctx=ctx)
    ici, single, binary, t1, t2 = bench.iso_world(ctx, bc)
    out = bench.sharded_ensemble(ctx, single, t1, rk, args, n_walkers=args.walkers, n_steps=args.steps)
    rk.close()
    if rk.rank == 0:
        out.pop("config", None)
        out["world"] = rk.world
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
