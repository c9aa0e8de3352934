"""Worker of tests/test_gpu_peer.py (one process per GPU under torchrun): fused lnpost + peer all-gather vs a plain
evaluation of all rows on this rank's own GPU."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    import isochrones_b200 as ib
    from isochrones_b200 import _lib, parallel, synthetic as syn

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo")
    ctx = _lib.default_context(rank)
    trk = syn.make_track_grid(n_feh=6, n_mass=24, n_eep=171)
    bc = syn.make_bc_grid(bands=("V", "J", "H", "K"), n_teff=24, n_logg=10, n_feh=8, n_av=7)
    ic = ib.ichrone_from_arrays("track", trk, bc, ctx=ctx)
    truth = syn.default_truth("track", n_eep=171)
    _, _, _, mags = ic.interp_mag(list(truth), ["V", "J", "H", "K"])
    mod = ib.BasicStarModel(ic, Teff=(5772.0, 80.0), parallax=(10.0, 0.1), **{b: (float(m), 0.02) for b, m in zip("VJHK", mags)})
    n_total = 10_003                                      # ragged: the last rank's block is shorter than the pad
    sh = parallel.RowSharder(n_total, world, rank)
    peer = parallel.PeerGather(ctx, rank, world, sh.pad, parallel.torch_allgather_bytes(dist))
    for step in range(5):                                 # several steps: both parity buffers, flags advancing
        rows = syn.posterior_like_batch("track", n_total, truth, n_eep=171, seed=100 + step)
        rows[::97, 1] = 1e9                               # some -inf rows
        mine = np.ascontiguousarray(sh.local(rows))
        d_p = ctx.dev_alloc(mine.nbytes)
        ctx.h2d(d_p, mine)
        d_all = peer.lnpost(mod.compiled, d_p, len(mine))
        got = np.empty((world, sh.pad))
        ctx.d2h(got, d_all)
        ctx.dev_free(d_p)
        full = sh.assemble(got)
        want = mod.lnpost_batch(rows)
        assert np.array_equal(full, want, equal_nan=True), "step %d rank %d" % (step, rank)
        dist.barrier()
    peer.close()
    dist.destroy_process_group()
    print("rank %d ok" % rank, flush=True)


if __name__ == "__main__":
    main()
