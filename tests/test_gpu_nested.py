"""GPU: nested sampling on batched device evaluations (isochrones_b200/nested.py — the reference's default fit path,
pymultinest.run(mnest_loglike, mnest_prior), starmodel.py:717-802).

MultiNest is third-party and absent, so the check is against what it estimates: Z = integral over the unit cube of
exp(lnpost), computed here by brute force — tens of millions of uniform box draws evaluated on the device
(prior_box_draws) — on a model whose box is narrow enough for plain Monte Carlo to converge; and the weighted posterior
against the importance-weighted moments of the same draws."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_evidence_and_posterior_match_brute_force_integration():
    import isochrones_b200 as ib
    from isochrones_b200 import _lib, synthetic as syn

    ctx = _lib.default_context()
    trk = syn.make_track_grid(n_feh=9, n_mass=60, n_eep=513)
    bc = syn.make_bc_grid(bands=("V", "J", "H", "K"), n_teff=40, n_logg=14, n_feh=10, n_av=9)
    ic = ib.ichrone_from_arrays("track", trk, bc, ctx=ctx)
    truth = syn.default_truth("track", n_eep=513)
    _, _, _, mags = ic.interp_mag(list(truth), ["V", "J", "H", "K"])
    mod = ib.BasicStarModel(ic, Teff=(5772.0, 150.0), logg=(4.44, 0.2), feh=(0.0, 0.2), parallax=(10.0, 0.5),
                            **{b: (float(m), 0.05) for b, m in zip("VJHK", mags)})
    mod.set_bounds(mass=(0.7, 1.4), eep=(60.0, 150.0), feh=(-0.5, 0.5), distance=(85.0, 118.0), AV=(0.0, 0.6))
    # brute force: Z = mean over the box of exp(lnpost)
    n_mc, chunk = 40_000_000, 4_000_000
    lmax, s0, s1, s2 = -np.inf, 0.0, np.zeros(5), np.zeros(5)
    parts = []
    for c in range(n_mc // chunk):
        pars, lp = mod.prior_box_draws(chunk, seed=123, row0=c * chunk)
        lp = np.where(np.isfinite(lp), lp, -np.inf)
        parts.append((pars[lp > lp.max() - 40], lp[lp > lp.max() - 40]))
    pars = np.concatenate([p for p, _ in parts])
    lp = np.concatenate([l for _, l in parts])
    lmax = lp.max()
    w = np.exp(lp - lmax)
    logz_mc = lmax + np.log(w.sum() / n_mc)
    ess = w.sum() ** 2 / np.sum(w * w)
    assert ess > 2000, ess                                   # the brute-force estimate itself is converged
    mean_mc = (w @ pars) / w.sum()
    std_mc = np.sqrt((w @ (pars - mean_mc) ** 2) / w.sum())

    res = mod.fit_nested(n_live_points=1500, seed=3)
    assert res.converged and res.n_iter > 5000 and 0.01 < res.efficiency <= 1.0
    assert abs(res.weights.sum() - 1.0) < 1e-12 and res.samples.shape[1] == 5
    # evidence: within 4 sigma of the nested-sampling error estimate (+ the Monte-Carlo error of the brute force)
    tol = 4.0 * np.hypot(res.logZ_err, 1.0 / np.sqrt(ess))
    assert abs(res.logZ - logz_mc) < tol, (res.logZ, logz_mc, tol)
    # posterior moments
    assert np.all(np.abs(res.mean() - mean_mc) < 0.15 * std_mc), (res.mean(), mean_mc, std_mc)
    assert np.all(np.abs(res.std() / std_mc - 1.0) < 0.15), (res.std(), std_mc)
    eq = res.equal_weighted(4000, seed=1)
    assert eq.shape == (4000, 5) and np.all(np.abs(eq.mean(axis=0) - mean_mc) < 0.2 * std_mc)
    # same seed, same run
    again = mod.fit_nested(n_live_points=1500, seed=3)
    assert again.logZ == res.logZ and again.n_evals == res.n_evals
