"""GPU: the on-device ensemble sampler (iso_sampler_*, SURVEY.md §8f-1 / BASELINE config 3).

emcee is third-party and absent here, so sampler parity is pinned two ways: (1) an EXACT replay — the kernel's
counter-based random stream is re-generated on the host and the stretch move is re-run with the oracle's lnpost;
accept/reject decisions, positions and lnprob must coincide step for step; (2) statistics of a longer run (posterior
mean / width / acceptance fraction) against the same host algorithm.
"""
import numpy as np
import pytest

from tests.helpers import golden_grids, product_ic, replay_stretch_move

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup(golden):
    import isochrones_b200 as ib
    from isochrones_b200 import _lib, synthetic as syn
    from oracle import oracle

    ctx = _lib.default_context()
    trk = syn.make_track_grid(n_feh=6, n_mass=24, n_eep=171)
    iso = syn.make_iso_grid(n_age=20, n_feh=6, n_eep=171)
    bc = syn.make_bc_grid(bands=("V", "J", "H", "K"), n_teff=24, n_logg=10, n_feh=8, n_av=7)
    out = {}
    for kind, model, N in (("track", trk, 1), ("iso", iso, 2)):
        ic = ib.ichrone_from_arrays(kind, model, bc, ctx=ctx)
        truth = syn.default_truth(kind, n_eep=171, n_stars=N)
        prim = list(truth) if kind == "track" else [truth[0]] + list(truth[N:])
        _, _, _, mags = ic.interp_mag(prim, ["V", "J", "H", "K"])
        obs = {b: (float(m) - (0.4 if N == 2 else 0.0), 0.05) for b, m in zip("VJHK", mags)}
        mod = ib.BasicStarModel(ic, N=N, Teff=(5772.0, 150.0), logg=(4.44, 0.2), feh=(0.0, 0.2), parallax=(10.0, 0.5), **obs)
        out[kind] = (mod, oracle.StarModel(mod), truth, syn)
    return out


@pytest.mark.parametrize("kind", ["track", "iso"])
def test_exact_replay_against_oracle(setup, kind):
    from isochrones_b200.sampler import DeviceEnsembleSampler

    mod, om, truth, syn = setup[kind]
    n_walkers, n_steps, seed = 48, 60, 1234
    N = mod.N
    p0 = syn.posterior_like_batch(kind, n_walkers, truth, n_eep=171, seed=3) if N == 1 else None
    if p0 is None:
        p0 = syn.posterior_like_batch(kind, n_walkers, truth, n_eep=171, seed=3)
        p0[:, :N] = -np.sort(-p0[:, :N], axis=1)          # eep_0 >= eep_1 (prior ordering)
    smp = DeviceEnsembleSampler(mod.compiled, n_walkers, p0, seed=seed)
    smp.run_mcmc(n_steps)
    chain, lnp, n_acc = replay_stretch_move(lambda r: om.lnpost_batch(np.ascontiguousarray(r)), p0, n_steps, seed)
    got = smp.chains[:, 0]
    assert got.shape == chain.shape
    # identical decisions => identical positions (the proposal arithmetic is unfused on both sides)
    same = np.all(got == chain, axis=(1, 2))
    assert same.all(), "first diverging step: %d" % int(np.argmin(same))
    assert np.allclose(smp.lnprobs[:, 0], lnp, rtol=1e-12, atol=1e-8)
    pos, lnprob, acc, prop = smp.state()
    assert acc[0] == n_acc and prop == n_steps * n_walkers
    assert 0.1 < smp.acceptance_fraction[0] < 0.9


def test_many_chains_and_statistics(setup):
    """64 independent chains in one launch; per-chain streams differ; pooled posterior brackets the truth."""
    from isochrones_b200.sampler import DeviceEnsembleSampler

    mod, om, truth, syn = setup["track"]
    n_chains, n_walkers = 64, 32
    p0 = np.stack([syn.posterior_like_batch("track", n_walkers, truth, n_eep=171, seed=100 + c) for c in range(n_chains)])
    smp = DeviceEnsembleSampler(mod.compiled, n_walkers, p0, seed=7, n_chains=n_chains)
    smp.run_mcmc(300, store=False)          # burn-in
    smp.reset()
    smp.run_mcmc(400, thin=4)
    ch = smp.chains                          # [100, 64, 32, 5]
    assert ch.shape == (100, n_chains, n_walkers, 5)
    assert not np.array_equal(ch[:, 0], ch[:, 1])
    flat = ch.reshape(-1, 5)
    mean, std = flat.mean(0), flat.std(0)
    assert abs(mean[3] - truth[3]) < 3 * std[3] and std[3] < 10.0          # distance pinned by the parallax
    assert abs(mean[0] - truth[0]) < 3 * std[0]
    af = smp.acceptance_fraction
    assert af.shape == (n_chains,) and np.all(af > 0.1) and np.all(af < 0.9)
    # every stored lnprob is the lnpost of the stored position
    idx = np.random.RandomState(0).choice(len(flat), 3000, replace=False)
    lp = smp.lnprobs.reshape(-1)[idx]
    want = om.lnpost_batch(np.ascontiguousarray(flat[idx]))
    assert np.allclose(lp, want, rtol=1e-12, atol=1e-8)
    # continuing a run == one longer run (counter-based stream)
    a = DeviceEnsembleSampler(mod.compiled, n_walkers, p0[0], seed=11)
    a.run_mcmc(20)
    a.run_mcmc(30)
    b = DeviceEnsembleSampler(mod.compiled, n_walkers, p0[0], seed=11)
    b.run_mcmc(50)
    assert np.array_equal(a.chains, b.chains)


def test_catalog_chains(setup):
    """One star per chain (catalog mode): chain c must sample model c."""
    import isochrones_b200 as ib
    from isochrones_b200.sampler import DeviceEnsembleSampler
    from oracle import oracle

    mod, om, truth, syn = setup["track"]
    ic = mod.ic
    models = []
    for s in range(6):
        t = truth.copy()
        t[3] = 60.0 + 25.0 * s
        _, _, _, mags = ic.interp_mag(list(t), ["V", "K"])
        models.append(ib.BasicStarModel(ic, V=(float(mags[0]), 0.05), K=(float(mags[1]), 0.05), parallax=(1000.0 / t[3], 0.3)))
    compiled, _ = ib.compile_catalog(models)
    n_walkers = 32
    p0 = np.stack([syn.posterior_like_batch("track", n_walkers, np.r_[truth[:3], 60.0 + 25.0 * s, truth[4]], n_eep=171, seed=s)
                   for s in range(6)])
    smp = DeviceEnsembleSampler(compiled, n_walkers, p0, seed=5, n_chains=6)
    smp.run_mcmc(200, store=False)
    smp.run_mcmc(200, thin=2)
    d = smp.chains[..., 3].reshape(100, 6, -1).mean(axis=(0, 2))
    assert np.all(np.abs(d - (60.0 + 25.0 * np.arange(6))) < 8.0), d
    for s in (0, 5):
        om_s = oracle.StarModel(models[s])
        chain, lnp, _ = replay_stretch_move(lambda r: om_s.lnpost_batch(np.ascontiguousarray(r)), p0[s], 40, 5, chain=s)
        again = DeviceEnsembleSampler(compiled, n_walkers, p0, seed=5, n_chains=6)
        again.run_mcmc(40)
        assert np.array_equal(again.chains[:, s], chain)


def test_work_queue_of_multi_wave_runs_matches_the_oracle_chains(setup):
    """More chains than resident CTAs: the kernel works through (chain, 16-step segment) items from a queue.  The chains
    must not depend on the schedule — each one is replayed by the oracle's C driver with the same stream."""
    from isochrones_b200.sampler import DeviceEnsembleSampler
    from oracle import oracle

    mod, om, truth, syn = setup["track"]
    n_chains, nw, n_steps, seed = 2900, 8, 40, 99          # 2900 one-warp CTAs on 148 SMs x 16 resident: 1.22 waves
    base = syn.posterior_like_batch("track", 4000, truth, n_eep=171, seed=21)
    base = base[np.isfinite(om.lnpost_batch(base))][:nw * 50]
    p0 = np.ascontiguousarray(np.stack([base[(c % 50) * nw:(c % 50 + 1) * nw] for c in range(n_chains)]))
    smp = DeviceEnsembleSampler(mod.compiled, nw, p0, seed=seed, n_chains=n_chains, moments=True)
    smp.run_mcmc(n_steps, thin=8)
    pos, lnp, acc, prop = smp.state()
    mean, std, cnt = smp.moments()
    chain, _, opos, olp, oacc = oracle.stretch_move(om, p0, n_steps, seed, n_threads=8)
    assert np.array_equal(pos, opos)                       # identical decisions -> identical ensembles, all 2600 chains
    assert np.allclose(lnp, olp, rtol=1e-12, atol=1e-8) and np.array_equal(acc, oacc)
    assert np.array_equal(smp.chains, chain[7::8])         # kept (thinned) ensembles land in the right slots
    assert (cnt == (n_steps // 8) * nw).all()
    assert np.allclose(mean, chain[7::8].mean(axis=(0, 2)), rtol=1e-11, atol=1e-11)
    smp.close()


@pytest.mark.parametrize("kind", ["track", "iso"])
def test_fit_mcmc_samples_and_derived_samples(setup, kind):
    """`BasicStarModel.fit_mcmc` (burn-in, reset, production: starmodel.py:889-972) and what the reference builds from the
    samples afterwards (`samples`, `derived_samples`: starmodel.py:1646-1714) — single star and binary."""
    mod, _, truth, _ = setup[kind]
    nw, niter = 64, 120
    sampler = mod.fit_mcmc(nwalkers=nw, nburn=80, niter=niter, seed=5)
    assert sampler.chain.shape == (nw, niter, mod.n_params) and sampler.lnprobability.shape == (nw, niter)
    acc = float(np.mean(sampler.acceptance_fraction))
    assert 0.05 < acc < 0.95, acc
    sam = mod.samples
    assert list(sam.columns) == list(mod.param_names) + ["lnprob"] and len(sam) == nw * niter
    assert np.isfinite(sam["lnprob"]).all()
    # lnprob IS the lnpost of the stored positions
    again = mod.lnpost_batch(sam[list(mod.param_names)].to_numpy()[:500])
    assert np.allclose(again, sam["lnprob"].to_numpy()[:500], rtol=1e-12, atol=1e-9)
    der = mod.derived_samples
    assert len(der) == len(sam) and list(der.columns)[-1] in ("AV", "parallax")
    assert np.allclose(der["parallax"], 1000.0 / sam["distance"]) and np.array_equal(der["distance"], sam["distance"])
    teff = "Teff" if mod.N == 1 else "Teff_0"
    assert np.isfinite(der[teff]).all() and abs(np.median(der[teff]) - 5772.0) < 600.0
    for b in mod.bands:
        assert np.isfinite(der[b + "_mag"]).all()
        if mod.N == 2:      # the system is brighter than either star
            assert (der[b + "_mag"] < der[b + "_mag_0"]).all() and (der[b + "_mag"] < der[b + "_mag_1"]).all()
    assert mod.derived_samples is der                     # cached until the next fit
