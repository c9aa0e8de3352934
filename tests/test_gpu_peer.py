"""GPU: the fused lnpost + all-gather over peer memory (iso_peer_*).

On one GPU the group has a single rank: the PEER kernel variant stores into its own receive buffer and the flag
exchange is with itself — the kernel and the step bookkeeping are covered.  With two or more GPUs visible a torchrun
job (world size 2) checks the real exchange against a plain evaluation of all rows, over several steps."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_single_rank_group_matches_batch_kernel():
    import isochrones_b200 as ib
    from isochrones_b200 import _lib, parallel, synthetic as syn

    ctx = _lib.default_context()
    iso = syn.make_iso_grid(n_age=20, n_feh=6, n_eep=171)
    bc = syn.make_bc_grid(bands=("V", "J", "H", "K"), n_teff=24, n_logg=10, n_feh=8, n_av=7)
    ic = ib.ichrone_from_arrays("iso", iso, bc, ctx=ctx)
    for N in (1, 2):
        truth = syn.default_truth("iso", n_eep=171, n_stars=N)
        _, _, _, mags = ic.interp_mag([truth[0]] + list(truth[N:]), ["V", "J", "H", "K"])
        mod = ib.BasicStarModel(ic, N=N, logg=(4.44, 0.1), parallax=(10.0, 0.1),
                                **{b: (float(m) - 0.3 * (N - 1), 0.02) for b, m in zip("VJHK", mags)})
        n = 5000
        peer = parallel.PeerGather(ctx, 0, 1, n, None)
        for step in range(3):
            rows = syn.posterior_like_batch("iso", n, truth, n_eep=171, seed=7 + step)
            rows[:, :N] = -np.sort(-rows[:, :N], axis=1)
            rows[::41, 0] = -5.0
            d_p = ctx.dev_alloc(rows.nbytes)
            ctx.h2d(d_p, rows)
            m = n - 13 * step                      # fewer rows than the pad: the tail of the block is left alone
            d_all = peer.lnpost(mod.compiled, d_p, m)
            got = np.empty(n)
            ctx.d2h(got, d_all)
            ctx.dev_free(d_p)
            assert np.array_equal(got[:m], mod.lnpost_batch(rows[:m]), equal_nan=True)
        peer.lnpost(mod.compiled, None, 0)          # a step without rows still completes (the flags advance)
        ctx.sync()
        peer.close()
    # a model with a non-default prior has no fused variant: a loud error, not a silent fallback
    from isochrones_b200.priors import GaussianPrior
    mod.set_prior(age=GaussianPrior(9.6, 0.2, bounds=(8, 10)))
    peer = parallel.PeerGather(ctx, 0, 1, 16, None)
    d_p = ctx.dev_alloc(16 * 8 * mod.n_params)
    with pytest.raises(_lib.IsoError):
        peer.lnpost(mod.compiled, d_p, 16)
    ctx.dev_free(d_p)
    peer.close()


def test_two_ranks_over_peer_memory():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "_peer_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "rank 0 ok" in r.stdout and "rank 1 ok" in r.stdout
