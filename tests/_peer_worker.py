"""Worker of tests/test_gpu_peer.py — one process per rank, launched by the test itself (plain subprocesses; RANK /
WORLD_SIZE / ISO_B200_RDZV in the environment, no torch).  Rank r runs on GPU ``r % device_count``: with one visible
GPU both ranks share it (separate processes, CUDA-IPC mappings of each other's buffers), with two they use NVLink.

    mode "gather":  fused lnpost + peer all-gather over several steps vs a plain evaluation of all rows;
    mode "timeout": rank 1 stops publishing steps; rank 0's bounded wait must turn into ISO_E_TIMEOUT, not a hang;
    mode "fallback": rank 1's peer setup fails: parallel.row_gather moves ALL ranks to kernel + ncclAllGather (needs one
                     GPU per rank), same results."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import isochrones_b200 as ib
    from isochrones_b200 import _lib, parallel, synthetic as syn

    mode = sys.argv[1] if len(sys.argv) > 1 else "gather"
    rdzv = parallel.FileRendezvous.from_env(timeout=120.0)
    rank, world = rdzv.rank, rdzv.world
    ctx = _lib.default_context(rank % _lib.device_count())
    trk = syn.make_track_grid(n_feh=6, n_mass=24, n_eep=171)
    bc = syn.make_bc_grid(bands=("V", "J", "H", "K"), n_teff=24, n_logg=10, n_feh=8, n_av=7)
    ic = ib.ichrone_from_arrays("track", trk, bc, ctx=ctx)
    truth = syn.default_truth("track", n_eep=171)
    _, _, _, mags = ic.interp_mag(list(truth), ["V", "J", "H", "K"])
    mod = ib.BasicStarModel(ic, Teff=(5772.0, 80.0), parallax=(10.0, 0.1), **{b: (float(m), 0.02) for b, m in zip("VJHK", mags)})
    n_total = 10_003                                      # ragged: the last rank's block is shorter than the pad
    sh = parallel.RowSharder(n_total, world, rank)
    if mode == "fallback":
        def failing():
            if rank == 1:
                raise RuntimeError("peer access refused (forced by the test)")
            return parallel.PeerGather(ctx, rank, world, sh.pad, parallel._agreeing(rdzv.allgather_bytes))

        # rank 1 never reaches the handle exchange, so the other ranks' exchange must not wait for it: they ship their
        # handles only after everyone has reported in
        ok = rdzv.allgather_bytes(b"0" if rank == 1 else b"1")
        make_peer = failing if all(v == b"1" for v in ok) else (lambda: (_ for _ in ()).throw(RuntimeError("peer access refused")))
        peer, why = parallel.row_gather(ctx, rank, world, sh.pad, rdzv.allgather_bytes, broadcast=rdzv.broadcast, make_peer=make_peer)
        assert isinstance(peer, parallel.NcclRowGather) and "refused" in why, (type(peer), why)
    else:
        peer, why = parallel.row_gather(ctx, rank, world, sh.pad, rdzv.allgather_bytes, broadcast=rdzv.broadcast)
        assert isinstance(peer, parallel.PeerGather), why
    if mode == "timeout":
        peer.set_timeout(1.0)
    n_steps = 5
    for step in range(n_steps):                           # several steps: both parity buffers, flags advancing
        rows = syn.posterior_like_batch("track", n_total, truth, n_eep=171, seed=100 + step)
        rows[::97, 1] = 1e9                               # some -inf rows
        mine = np.ascontiguousarray(sh.local(rows))
        if mode == "timeout" and step == 2:
            if rank == 1:
                break                                     # this rank never publishes step 3
            d_p = ctx.dev_alloc(mine.nbytes)
            ctx.h2d(d_p, mine)
            peer.lnpost(mod.compiled, d_p, len(mine))
            try:
                peer.check()
            except _lib.IsoError as e:
                assert e.code == -6 and "rank 1" in str(e), str(e)
                try:                                      # sticky: the group refuses further steps
                    peer.lnpost(mod.compiled, d_p, len(mine))
                except _lib.IsoError as e2:
                    assert e2.code == -6
                    print("rank 0 timeout reported", flush=True)
                    break
            raise AssertionError("no timeout reported")
        d_p = ctx.dev_alloc(mine.nbytes)
        ctx.h2d(d_p, mine)
        d_all = peer.lnpost(mod.compiled, d_p, len(mine))
        got = np.empty((world, sh.pad))
        ctx.d2h(got, d_all)
        ctx.dev_free(d_p)
        peer.check()
        full = sh.assemble(got)
        want = mod.lnpost_batch(rows)
        assert np.array_equal(full, want, equal_nan=True), "step %d rank %d" % (step, rank)
        rdzv.barrier()
    rdzv.barrier()                                        # nobody unmaps while a peer may still store
    peer.close()
    rdzv.close()
    print("rank %d ok" % rank, flush=True)


if __name__ == "__main__":
    main()
