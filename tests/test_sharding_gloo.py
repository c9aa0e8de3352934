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


def test_row_gather_decision_is_collective():
    """`parallel.row_gather`: every rank ends up with the same kind of gather; one rank failing its peer setup moves
    ALL ranks to the NCCL form (stand-ins for the device objects: the decision logic is host code)."""
    import threading

    from isochrones_b200 import parallel

    world = 3
    barrier = threading.Barrier(world)
    slots = {}
    lock = threading.Lock()

    def make_allgather(rank):
        state = {"round": 0}

        def allgather(payload):
            r = state["round"]
            state["round"] += 1
            with lock:
                slots.setdefault(r, {})[rank] = payload
            barrier.wait()
            return [slots[r][k] for k in range(world)]
        return allgather

    class FakePeer(object):
        closed = False

        def close(self):
            self.closed = True

    class FakeCtx(object):
        def dev_alloc(self, n):
            return n

        def dev_free(self, p):
            pass

        def memset(self, *a):
            pass

    def run(fail_rank, results):
        def worker(rank):
            def make_peer():
                if rank == fail_rank:
                    raise RuntimeError("cudaIpcOpenMemHandle: peer access not supported")
                return FakePeer()
            g, why = parallel.row_gather(FakeCtx(), rank, world, 128, make_allgather(rank), comm=object(), make_peer=make_peer)
            results[rank] = (type(g).__name__, why)
        ts = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
        [t.start() for t in ts]
        [t.join(30) for t in ts]

    res = {}
    run(None, res)
    assert all(res[r] == ("FakePeer", "") for r in range(world)), res
    slots.clear()
    res = {}
    run(1, res)
    assert all(res[r][0] == "NcclRowGather" for r in range(world)), res
    assert all("peer access not supported" in res[r][1] for r in range(world)), res
