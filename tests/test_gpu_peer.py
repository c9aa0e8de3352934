"""GPU: the fused lnpost + all-gather over peer memory (iso_peer_*).

Single-rank group: the PEER kernel variant stores into its own receive buffer and the flag exchange is with itself —
the kernel and the step bookkeeping are covered.  Two-rank group: the test launches two worker processes itself (no
torchrun); rank r uses GPU r % device_count, so on a one-GPU box both ranks share the GPU and still exchange through
CUDA-IPC mappings of each other's buffers, and on a multi-GPU box the stores travel over NVLink.  A third test checks
that a rank which stops publishing steps produces ISO_E_TIMEOUT on its peer instead of a hung stream."""
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
    # a model with a non-default prior runs the generic-profile variant of the fused kernel
    from isochrones_b200.priors import GaussianPrior
    mod.set_prior(age=GaussianPrior(9.6, 0.2, bounds=(8, 10)))
    peer = parallel.PeerGather(ctx, 0, 1, 4000, None)
    rows = syn.posterior_like_batch("iso", 4000, truth, n_eep=171, seed=11)
    rows[:, :2] = -np.sort(-rows[:, :2], axis=1)
    d_p = ctx.dev_alloc(rows.nbytes)
    ctx.h2d(d_p, rows)
    got = np.empty(4000)
    ctx.d2h(got, peer.lnpost(mod.compiled, d_p, 4000))
    peer.check()
    assert np.array_equal(got, mod.lnpost_batch(rows), equal_nan=True) and np.isfinite(got).sum() > 1000
    ctx.dev_free(d_p)
    peer.close()


def _run_ranks(mode, world=2, timeout=300):
    """Launch `world` worker processes (rank r on GPU r % device_count) and return their outputs."""
    import tempfile

    procs = []
    with tempfile.TemporaryDirectory() as rdzv:
        for r in range(world):
            env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), ISO_B200_RDZV=rdzv)
            procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_peer_worker.py"), mode],
                                          stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env))
        outs = []
        for p in procs:
            try:
                out, _ = p.communicate(timeout=timeout)
            except subprocess.TimeoutExpired:
                for q in procs:
                    q.kill()
                raise
            outs.append((p.returncode, out))
    return outs


def test_two_ranks_over_peer_memory():
    """Two processes exchange their rows through each other's CUDA-IPC mapped buffers — over NVLink when two GPUs are
    visible, through the one GPU's memory when the box has a single GPU (the driver's test box): same code path."""
    outs = _run_ranks("gather")
    for r, (rc, out) in enumerate(outs):
        assert rc == 0 and ("rank %d ok" % r) in out, out[-3000:]


def test_missing_rank_times_out_instead_of_hanging():
    outs = _run_ranks("timeout")
    assert outs[0][0] == 0 and "rank 0 timeout reported" in outs[0][1], outs[0][1][-3000:]
    assert outs[1][0] == 0, outs[1][1][-3000:]


def test_failed_peer_setup_moves_every_rank_to_nccl():
    """`parallel.row_gather`: one rank cannot map its peers -> ALL ranks use kernel + ncclAllGather, same results.  NCCL
    needs one GPU per rank, so this runs on multi-GPU boxes only (the decision logic itself: tests/test_sharding_gloo.py)."""
    from isochrones_b200 import _lib

    if _lib.device_count() < 2:
        pytest.skip("ncclAllGather needs one GPU per rank")
    outs = _run_ranks("fallback")
    for r, (rc, out) in enumerate(outs):
        assert rc == 0 and ("rank %d ok" % r) in out, out[-3000:]
