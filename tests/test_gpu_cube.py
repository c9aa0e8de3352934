"""GPU: the MultiNest-shaped entries — `mnest_prior` (starmodel.py:1637-1640), the fused `mnest_prior` + `mnest_loglike`
batch (one launch for a whole set of live points) and the on-device prior-box draws — plus the sampler's `reset` and
running moments.

`mnest_prior` is pinned bit for bit by the reference's own outputs (tests/golden, lp_*_cube_phys); the fused entry must
leave the same numbers in the cube and return the lnpost the plain batch entry gives for them; the device draws are
replayed on the host from the same Philox4x32-10 counters."""
import numpy as np
import pytest

from tests.helpers import golden_grids, load_specs, philox4x32_10, product_ic, product_model_from_spec, u01

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def world(golden):
    from isochrones_b200 import _lib

    ctx = _lib.default_context()
    gi, gl = golden["interp"], golden["lnpost"]
    trk, iso, bc = golden_grids(gi)
    ics = {"track": product_ic("track", trk, bc, ctx=ctx), "iso": product_ic("iso", iso, bc, ctx=ctx)}
    return ics, load_specs(gl), gl


def test_fused_mnest_batch_matches_reference_cube_and_batch_lnpost(world):
    ics, specs, gl = world
    n_checked = 0
    for name, spec in specs.items():
        mod = product_model_from_spec(spec, ics[spec["kind"]])
        cube = np.ascontiguousarray(gl["lp_%s_cube" % name].copy())
        want_phys = gl["lp_%s_cube_phys" % name]                 # the reference's mnest_prior, row by row
        lnpost, lnprior, lnlike = mod.mnest_lnpost_batch(cube, parts=True)
        assert np.array_equal(cube, want_phys), name             # mapped in place, bit-exact (unfused multiply-add)
        ref_post, ref_prior, ref_like = mod.lnpost_batch(want_phys, parts=True)
        assert np.array_equal(lnpost, ref_post, equal_nan=True) and np.array_equal(lnprior, ref_prior, equal_nan=True)
        assert np.array_equal(lnlike, ref_like, equal_nan=True)
        # lnpost only; a pinned cube (zero-copy small-call path) gives the same
        pinned = mod.ic.ctx.pinned_empty(cube.shape)
        pinned[:] = gl["lp_%s_cube" % name]
        assert np.array_equal(mod.mnest_lnpost_batch(pinned), ref_post, equal_nan=True) and np.array_equal(pinned, want_phys)
        # MultiNest's own calling sequence per live point: mnest_prior(cube) then mnest_loglike(cube)
        row = list(gl["lp_%s_cube" % name][0])
        mod.mnest_prior(row, len(row), len(row))
        one = mod.mnest_loglike(row, len(row), len(row))
        assert (one == ref_post[0]) or (np.isnan(one) and np.isnan(ref_post[0]))
        n_checked += 1
        with pytest.raises(ValueError):
            mod.mnest_lnpost_batch(cube[:, :-1])
    assert n_checked >= 9


def test_large_fused_batch_goes_through_the_pipeline(world):
    ics, specs, gl = world
    name = next(n for n, s in specs.items() if s["kind"] == "track" and s["N"] == 1)
    mod = product_model_from_spec(specs[name], ics["track"])
    rng = np.random.RandomState(5)
    cube = rng.random_sample((700_000, mod.n_params))             # several pipeline chunks, ragged tail
    lo = np.array([mod.bounds(p)[0] for p in mod.param_names])
    hi = np.array([mod.bounds(p)[1] for p in mod.param_names])
    phys = (hi - lo) * cube + lo                                 # numpy evaluates this unfused, like the reference
    got = mod.mnest_lnpost_batch(cube)
    assert np.array_equal(cube, phys)
    assert np.array_equal(got, mod.lnpost_batch(phys), equal_nan=True)
    assert np.isfinite(got).sum() > 1000


def _host_draws(seed, row0, n, lo, hi):
    g = np.arange(row0, row0 + n, dtype=np.uint64)
    c0, c1 = (g & np.uint64(0xFFFFFFFF)), (g >> np.uint64(32))
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    us = []
    for blk in range(4):
        r = philox4x32_10(c0, c1, np.uint32(blk), np.uint32(0x43554245), k0, k1)
        us += [u01(r[0], r[1]), u01(r[2], r[3])]
    u = np.stack(us[:len(lo)], axis=1)
    return (hi - lo) * u + lo


def test_prior_box_draws_replay_and_split_independence(world):
    ics, specs, gl = world
    for name, spec in specs.items():
        if spec["N"] == 3:
            continue
        mod = product_model_from_spec(spec, ics[spec["kind"]])
        lo = np.array([mod.bounds(p)[0] for p in mod.param_names], dtype=float)
        hi = np.array([mod.bounds(p)[1] for p in mod.param_names], dtype=float)
        seed, n = 0x1234567890ABCDEF, 5000
        pars, lnpost = mod.prior_box_draws(n, seed=seed)
        assert np.array_equal(pars, _host_draws(seed, 0, n, lo, hi)), name      # same stream, same unfused mapping
        assert np.array_equal(lnpost, mod.lnpost_batch(pars), equal_nan=True)
        # the rows do not depend on how the draw is split (row0 is the global counter)
        a = mod.prior_box_draws(2000, seed=seed, row0=0, return_pars=False)
        b = mod.prior_box_draws(3000, seed=seed, row0=2000, return_pars=False)
        assert np.array_equal(np.concatenate([a, b]), lnpost, equal_nan=True)
        assert (pars >= lo).all() and (pars <= hi).all()
    big = mod.prior_box_draws(600_000, seed=7, return_pars=False)     # pipeline chunks: 8 B per row back, nothing in
    assert np.array_equal(big[:4000], mod.prior_box_draws(4000, seed=7, return_pars=False), equal_nan=True)


def test_sampler_reset_and_running_moments(world):
    from isochrones_b200 import synthetic as syn
    from isochrones_b200.sampler import DeviceEnsembleSampler

    ics, specs, gl = world
    name = next(n for n, s in specs.items() if s["kind"] == "iso" and s["N"] == 1)
    mod = product_model_from_spec(specs[name], ics["iso"])
    pars = gl["lp_%s_pars" % name]
    good = pars[np.isfinite(gl["lp_%s_lnpost" % name])]
    n_chains, nw = 3, 32
    p0 = np.stack([good[c * nw:(c + 1) * nw] for c in range(n_chains)])
    smp = DeviceEnsembleSampler(mod.compiled, nw, p0, seed=9, n_chains=n_chains, moments=True)
    smp.run_mcmc(30, store=False)                                # burn-in
    _, _, acc_burn, prop_burn = smp.state()
    assert prop_burn == 30 * nw
    smp.reset()
    _, _, acc0, prop0 = smp.state()
    assert (acc0 == 0).all() and prop0 == 0                      # emcee's reset() zeroes naccepted
    smp.run_mcmc(40, thin=4)
    _, _, acc, prop = smp.state()
    assert prop == 40 * nw and (acc <= prop).all() and (acc > 0).any()
    mean, std, cnt = smp.moments()
    kept = smp.chains                                            # [10, n_chains, nw, ndim]
    assert (cnt == kept.shape[0] * nw).all()
    want_mean = kept.mean(axis=(0, 2))
    want_std = kept.std(axis=(0, 2))
    assert np.allclose(mean, want_mean, rtol=1e-12, atol=1e-12)
    assert np.allclose(std, want_std, rtol=1e-6, atol=1e-9)
    smp.close()
