"""GPU: ONE ensemble sharded over ranks (iso_ensemble_*) — the consumer of the fused peer exchange.

The sharded sampler must reproduce the one-GPU persistent sampler's chain bit for bit, whatever the number of ranks:
single rank in-process, then two and three worker processes launched by the test itself (rank r on GPU r %
device_count: on a one-GPU box the ranks share the GPU and still exchange through CUDA-IPC mappings)."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_single_rank_matches_persistent_sampler_and_large_ensemble():
    import isochrones_b200 as ib
    from isochrones_b200 import _lib, synthetic as syn
    from isochrones_b200.sampler import DeviceEnsembleSampler, ShardedEnsembleSampler

    ctx = _lib.default_context()
    trk = syn.make_track_grid(n_feh=6, n_mass=24, n_eep=171)
    bc = syn.make_bc_grid(bands=("V", "J", "H", "K"), n_teff=24, n_logg=10, n_feh=8, n_av=7)
    ic = ib.ichrone_from_arrays("track", trk, bc, ctx=ctx)
    truth = syn.default_truth("track", n_eep=171)
    _, _, _, mags = ic.interp_mag(list(truth), ["V", "J", "H", "K"])
    mod = ib.BasicStarModel(ic, Teff=(5772.0, 80.0), logg=(4.44, 0.1), parallax=(10.0, 0.1),
                            **{b: (float(m), 0.02) for b, m in zip("VJHK", mags)})
    nw = 128
    p0 = syn.posterior_like_batch("track", nw, truth, n_eep=171, seed=3)
    ref = DeviceEnsembleSampler(mod.compiled, nw, p0, seed=11)
    ref.run_mcmc(40, thin=4)
    ens = ShardedEnsembleSampler(mod.compiled, nw, p0, seed=11)
    ens.run_mcmc(40, thin=4)
    assert np.array_equal(ens.chains, ref.chains[:, 0]) and np.array_equal(ens.lnprobs, ref.lnprobs[:, 0])
    assert ens.state()[2] == ref.state()[2][0]
    ens.close()
    ref.close()
    # an ensemble far beyond one CTA's shared memory: 20 000 walkers move, stay finite and keep their lnpost consistent
    nw = 20_000
    p0 = syn.posterior_like_batch("track", nw, truth, n_eep=171, seed=4)
    big = ShardedEnsembleSampler(mod.compiled, nw, p0, seed=12)
    pos, lnp = big.run_mcmc(10, store=False)
    assert np.isfinite(lnp).all() and 0.1 < big.state()[2] / big.state()[3] < 0.9
    assert np.array_equal(mod.lnpost_batch(pos), lnp)
    big.close()
    # a NaN starting point is refused, as emcee does
    bad = p0[:64].copy()
    bad[3, 4] = 50.0                                     # AV far beyond the BC grid: lnpost NaN
    bad[3, 3] = 100.0
    mod2 = ib.BasicStarModel(ic, maxAV=100.0, Teff=(5772.0, 80.0), **{b: (float(m), 0.02) for b, m in zip("VJHK", mags)})
    if np.isnan(mod2.lnpost(bad[3])):
        with pytest.raises(_lib.IsoError):
            ShardedEnsembleSampler(mod2.compiled, 64, bad, seed=1)
        with pytest.raises(_lib.IsoError):
            DeviceEnsembleSampler(mod2.compiled, 64, bad, seed=1)


def _run_ranks(world, *args):
    procs = []
    with tempfile.TemporaryDirectory() as rdzv:
        for r in range(world):
            env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), ISO_B200_RDZV=rdzv)
            procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_ensemble_worker.py")] + list(args),
                                          stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env))
        outs = []
        for p in procs:
            try:
                out, _ = p.communicate(timeout=600)
            except subprocess.TimeoutExpired:
                for q in procs:
                    q.kill()
                raise
            outs.append((p.returncode, out))
    return outs


@pytest.mark.parametrize("world", [2, 3])
def test_ranks_reproduce_the_single_gpu_chain(world):
    for r, (rc, out) in enumerate(_run_ranks(world)):
        assert rc == 0 and ("rank %d ok" % r) in out, out[-3000:]


def test_missing_rank_turns_into_a_timeout_error():
    outs = _run_ranks(2, "timeout")
    assert outs[0][0] == 0 and "rank 0 timeout reported" in outs[0][1], outs[0][1][-3000:]
    assert outs[1][0] == 0, outs[1][1][-3000:]
