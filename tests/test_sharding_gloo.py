"""CPU, world_size 2 over gloo: the N > 1 host logic of the path — row sharding, padded all-gather, reassembly
(isochrones_b200/parallel.py).  The per-shard evaluator here is the oracle (test infrastructure standing in for
the GPU); on GPUs the same `sharded_lnpost` runs the fused kernel and `NcclGather`."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_rows, ret):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from isochrones_b200 import synthetic as syn
    from isochrones_b200.parallel import RowSharder, sharded_lnpost
    from oracle import oracle
    from tests.helpers import product_ic

    import isochrones_b200 as ib

    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        trk = syn.make_track_grid(n_feh=4, n_mass=12, n_eep=60)
        bc = syn.make_bc_grid(bands=("V", "K"), n_teff=14, n_logg=8, n_feh=6, n_av=5)
        ic = product_ic("track", trk, bc)
        mod = ib.BasicStarModel(ic, Teff=(5772.0, 80.0), V=(10.0, 0.02), K=(8.5, 0.02), parallax=(10.0, 0.1))
        om = oracle.StarModel(mod)
        truth = syn.default_truth("track", n_eep=60)
        pars = syn.posterior_like_batch("track", n_rows, truth, n_eep=60, seed=5)     # same on every rank

        class Shard:            # stands in for CompiledModel.lnpost on this rank's GPU
            calls = 0

            def lnpost(self, rows):
                Shard.calls += len(rows)
                return om.lnpost_batch(rows)

        def gather(send):
            out = [torch.empty(len(send), dtype=torch.float64) for _ in range(world)]
            dist.all_gather(out, torch.from_numpy(send))
            return np.stack([t.numpy() for t in out])

        sh = RowSharder(n_rows, world, rank)
        got = sharded_lnpost(Shard(), pars, sh, gather)
        want = om.lnpost_batch(pars)
        ok = np.array_equal(got, want, equal_nan=True) and Shard.calls == sh.counts[rank]
        ret.put((rank, bool(ok), int(Shard.calls)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_rows", [257, 64])
def test_sharded_lnpost_world2(n_rows):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_rows, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    res = sorted(ret.get(timeout=5) for _ in range(2))
    assert [r[1] for r in res] == [True, True]
    assert sum(r[2] for r in res) == n_rows          # every row evaluated exactly once across the ranks
