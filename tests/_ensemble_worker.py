"""Worker of tests/test_gpu_ensemble.py — one process per rank (RANK / WORLD_SIZE / ISO_B200_RDZV in the environment,
no torch; rank r on GPU r % device_count): the sharded ensemble sampler against the one-GPU persistent sampler run on
this rank's own GPU with the same seed.  The two must agree bit for bit on every kept ensemble.

    mode "timeout": rank 1 stops taking part after the first run; rank 0's next run must end with ISO_E_TIMEOUT (bounded
    waits in the half-step kernels), not hang."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import isochrones_b200 as ib
    from isochrones_b200 import _lib, parallel, synthetic as syn
    from isochrones_b200.sampler import DeviceEnsembleSampler, ShardedEnsembleSampler

    rdzv = parallel.FileRendezvous.from_env(timeout=120.0)
    rank, world = rdzv.rank, rdzv.world
    ctx = _lib.default_context(rank % _lib.device_count())
    iso = syn.make_iso_grid(n_age=20, n_feh=6, n_eep=171)
    bc = syn.make_bc_grid(bands=("V", "J", "H", "K"), n_teff=24, n_logg=10, n_feh=8, n_av=7)
    ic = ib.ichrone_from_arrays("iso", iso, bc, ctx=ctx)
    truth = syn.default_truth("iso", n_eep=171)
    _, _, _, mags = ic.interp_mag(list(truth), ["V", "J", "H", "K"])
    mod = ib.BasicStarModel(ic, Teff=(5772.0, 80.0), parallax=(10.0, 0.1), **{b: (float(m), 0.02) for b, m in zip("VJHK", mags)})
    nw, steps, thin = 90, 24, 3                          # 45 walkers per half: ragged blocks over two ranks
    p0 = syn.posterior_like_batch("iso", nw, truth, n_eep=171, seed=5)
    ref = DeviceEnsembleSampler(mod.compiled, nw, p0, seed=77)
    ref.run_mcmc(steps, thin=thin)
    ens = ShardedEnsembleSampler(mod.compiled, nw, p0, seed=77, rank=rank, world=world, allgather_bytes=rdzv.allgather_bytes)
    ens.run_mcmc(steps // 2, thin=thin)                  # two runs: the step counter and flags carry over
    if len(sys.argv) > 1 and sys.argv[1] == "timeout":
        if rank == 0:
            ens.set_timeout(0.5)
            try:
                ens.run_mcmc(1, store=False, fetch=False)
            except _lib.IsoError as e:
                assert e.code == -6 and "rank 1" in str(e), str(e)
                print("rank 0 timeout reported", flush=True)
            else:
                raise AssertionError("no timeout reported")
        rdzv.barrier()                                    # rank 1 keeps its memory mapped until rank 0 has given up
        ens.close()
        ref.close()
        rdzv.close()
        print("rank %d ok" % rank, flush=True)
        return
    ens.run_mcmc(steps // 2, thin=thin)
    assert np.array_equal(ens.chains, ref.chains[:, 0]), "rank %d: chain differs from the one-GPU sampler" % rank
    assert np.array_equal(ens.lnprobs, ref.lnprobs[:, 0]), "rank %d: lnprob differs" % rank
    pos, lnp, acc, prop = ens.state()
    rpos, rlnp, racc, rprop = ref.state()
    assert np.array_equal(pos, rpos[0]) and np.array_equal(lnp, rlnp[0]) and prop == rprop
    total_acc = sum(int(v) for v in rdzv.allgather_bytes(str(acc).encode()))
    assert total_acc == int(racc[0]) and 0 < total_acc < prop, (total_acc, racc, prop)
    rdzv.barrier()                                        # nobody unmaps while a peer may still store
    ens.close()
    ref.close()
    rdzv.close()
    print("rank %d ok (accepted %d of %d)" % (rank, total_acc, prop), flush=True)


if __name__ == "__main__":
    main()
