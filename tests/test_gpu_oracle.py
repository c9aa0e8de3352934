"""GPU: the CUDA path against the oracle (C restatement of the reference, pinned to the golden vectors by
tests/test_oracle_vs_golden.py) on seeded inputs at sizes the oracle finishes in seconds, on the FULL MIST-shaped
grids of the benchmark, plus size-independent properties at BASELINE.json's full batch size (1e6 rows).

Tolerances: 1e-6 relative on interpolated properties, 1e-4 absolute on lnpost (BASELINE.json north_star); asserted
much tighter where the arithmetic allows.  NaN / -inf patterns must be identical.
"""
import numpy as np
import pytest

import bench
from tests.helpers import assert_same_special, max_abs_err, max_rel_err

pytestmark = pytest.mark.gpu

RTOL_PROPS = 1e-6
ATOL_LNPOST = 1e-4


@pytest.fixture(scope="module")
def world():
    """Full-size synthetic MIST grids + the bench's star model, on the GPU and in the oracle."""
    import isochrones_b200 as ib
    from isochrones_b200 import _lib, synthetic as syn
    from oracle import oracle

    ctx = _lib.default_context()
    trk = syn.make_track_grid(columns=bench.PACK_COLUMNS)
    iso = syn.make_iso_grid(columns=("Teff", "logg", "feh", "Mbol", "mass", "dm_deep", "nu_max", "delta_nu"))
    bc = syn.make_bc_grid(bands=("V", "J", "H", "K", "G", "BP", "RP"))
    w = {"ctx": ctx, "trk": trk, "iso": iso, "bc": bc}
    w["ic_track"] = ib.ichrone_from_arrays("track", trk, bc, ctx=ctx)
    w["ic_iso"] = ib.ichrone_from_arrays("iso", iso, bc, ctx=ctx)
    w["og_track"] = oracle.Grid(trk["grid"], trk["axes"])
    w["og_iso"] = oracle.Grid(iso["grid"], iso["axes"])
    w["og_bc"] = oracle.Grid(bc["grid"], bc["axes"])
    return w


def _model(w, kind, N=1, bands=("V", "J", "H", "K"), **extra):
    import isochrones_b200 as ib
    from isochrones_b200 import synthetic as syn
    from oracle import oracle

    ic = w["ic_" + kind]
    truth = syn.default_truth(kind, n_stars=N)
    prim = list(truth) if kind == "track" else [truth[0]] + list(truth[N:])
    _, _, _, mags = ic.interp_mag(prim, list(bands))
    obs = {b: (float(np.round(m, 3)) - (0.3 if N > 1 else 0.0), 0.02) for b, m in zip(bands, mags)}
    kw = dict(Teff=(5772.0, 80.0), logg=(4.44, 0.1), feh=(0.0, 0.1), parallax=(10.0, 0.1))
    kw.update(extra)
    mod = ib.BasicStarModel(ic, N=N, **kw, **obs)
    om = oracle.StarModel(mod, model_grid=w["og_" + kind], bc_grid=w["og_bc"])
    return mod, om, truth


def _batches(kind, mod, truth, axes, n):
    from isochrones_b200 import synthetic as syn

    bounds = [mod.bounds(p) for p in mod.param_names]
    return np.concatenate([syn.posterior_like_batch(kind, n, truth, seed=11),
                           syn.prior_like_batch(kind, n, bounds, seed=12),
                           syn.edge_batch(kind, n // 4, truth, axes, bounds, seed=13)])


def _compare(got, want, atol):
    assert_same_special(got, want)
    m = np.isfinite(want)
    err = max_abs_err(got, want)
    # the budget is absolute; rows with |lnpost| ~ 1e6 carry ~1e-9 relative rounding headroom
    assert np.allclose(got[m], want[m], rtol=1e-12, atol=atol), err
    return err


@pytest.mark.parametrize("kind,N", [("track", 1), ("iso", 1), ("iso", 2), ("iso", 3)])
def test_lnpost_vs_oracle_full_grids(world, kind, N):
    mod, om, truth = _model(world, kind, N)
    pars = _batches(kind, mod, truth, world["trk" if kind == "track" else "iso"]["axes"], 60_000)
    lnpost, lnprior, lnlike = mod.lnpost_batch(pars, parts=True)
    o_post, o_prior, o_like = om.lnpost_batch(pars, n_threads=8, parts=True)
    _compare(lnprior, o_prior, 1e-9)
    _compare(lnlike, o_like, ATOL_LNPOST)
    _compare(lnpost, o_post, ATOL_LNPOST)
    assert np.isfinite(o_post).sum() > 50_000 and np.isneginf(o_post).sum() > 1000
    assert np.array_equal(mod.lnpost_batch(pars), lnpost, equal_nan=True)


MANY_BANDS = ("J", "H", "K", "G", "BP", "RP", "W1", "W2", "W3", "TESS", "Kepler", "V", "B", "u")


@pytest.mark.parametrize("N,n_bands", [(2, 7), (3, 7), (2, 11), (3, 11), (2, 14), (3, 5)])
def test_multi_star_band_chunks_vs_oracle(world, N, n_bands):
    """Binaries / triples whose BC pack holds 2, 3 and 4 chunks of four bands: the star-sequential kernels for up to three
    chunks, the chunk-major form beyond; summed fluxes (utils.py:67-75) in the reference's star order either way."""
    import isochrones_b200 as ib
    from isochrones_b200 import synthetic as syn
    from oracle import oracle

    bands = MANY_BANDS[:n_bands]
    bc = syn.make_bc_grid(bands=bands)
    w = {"ic_iso": ib.ichrone_from_arrays("iso", world["iso"], bc, ctx=world["ctx"]), "og_iso": world["og_iso"],
         "og_bc": oracle.Grid(bc["grid"], bc["axes"])}
    mod, om, truth = _model(w, "iso", N, bands=bands)
    assert len(mod.bands) == n_bands
    pars = _batches("iso", mod, truth, world["iso"]["axes"], 12_000)
    got = mod.lnpost_batch(pars, parts=True)
    want = om.lnpost_batch(pars, n_threads=8, parts=True)
    _compare(got[1], want[1], 1e-9)
    _compare(got[2], want[2], ATOL_LNPOST)
    _compare(got[0], want[0], ATOL_LNPOST)
    assert np.isfinite(want[0]).sum() > 8_000
    assert np.array_equal(mod.lnpost_batch(pars), got[0], equal_nan=True)


def test_seismology_and_bands_vs_oracle(world):
    mod, om, truth = _model(world, "track", 1, bands=("G", "BP", "RP", "J", "H", "K", "V"),
                            nu_max=(3000.0, 60.0), delta_nu=(130.0, 2.0))
    pars = _batches("track", mod, truth, world["trk"]["axes"], 20_000)
    got = mod.lnpost_batch(pars, parts=True)
    want = om.lnpost_batch(pars, n_threads=8, parts=True)
    for g, wv in zip(got, want):
        _compare(g, wv, ATOL_LNPOST)


def test_interp_values_vs_oracle_full_grid(world):
    """Config 1 on the full-size track grid: 1024 points (+ a 200k batch) -> Teff / logg / Mbol, 1e-6 relative."""
    it = world["ic_track"].model_grid.interp
    rng = np.random.RandomState(1)
    axes = world["trk"]["axes"]
    for n in (1024, 200_000):
        pts = [a[0] + (a[-1] - a[0]) * rng.random_sample(n) for a in axes]
        for d in range(3):      # a quarter of the points on nodes / edges / outside
            k = n // 4
            pts[d][:k] = rng.choice(axes[d], k)
            pts[d][k:k + 8] = [axes[d][0], axes[d][-1], axes[d][0] - 1e-9, axes[d][-1] + 1e-9, np.nan, axes[d][1],
                               axes[d][-2], axes[d][0]]
        pts[0][:] = np.where(pts[0] >= axes[0][-1], axes[0][-2], pts[0])      # keep the reference's UB corner out
        got = it(pts, ["Teff", "logg", "Mbol"])
        ci = it.column_index
        want = world["og_track"].interp_values(pts, [ci["Teff"], ci["logg"], ci["Mbol"]])
        assert_same_special(got, want)
        assert max_rel_err(got, want) < RTOL_PROPS
        assert max_rel_err(got, want) < 1e-11      # what the float64 kernel actually achieves (cancellation near 0 incl.)


def test_interp_mags_vs_oracle_full_grid(world):
    from isochrones_b200 import synthetic as syn
    from oracle import oracle

    ic = world["ic_track"]
    mod, om, truth = _model(world, "track", 1)
    pars = _batches("track", mod, truth, world["trk"]["axes"], 50_000)
    bands = ["K", "G", "V"]
    teff, logg, feh, mags = ic.interp_mag([pars[:, j] for j in range(5)], bands)
    ci = ic.model_grid.interp.column_index
    bi = ic.bc_grid.interp.column_index
    o = oracle.interp_mags(pars.T, ic.param_index_order, world["og_track"], ci["Teff"], ci["logg"], ci["feh"], ci["Mbol"],
                           world["og_bc"], [bi[b] for b in bands])
    for g, wv in zip((teff, logg, feh, mags), o):
        assert_same_special(g, wv)
        assert max_rel_err(g, wv) < 1e-9       # budget 1e-6; magnitudes pass through 0, where relative error amplifies


def test_full_size_properties(world):
    """BASELINE config 2 size (1e6 rows): properties that need no oracle at that size."""
    from isochrones_b200 import synthetic as syn

    mod, om, truth = _model(world, "track", 1)
    n = 1_000_000
    bounds = [mod.bounds(p) for p in mod.param_names]
    pars = np.concatenate([syn.posterior_like_batch("track", n // 2, truth, seed=21),
                           syn.prior_like_batch("track", n // 2, bounds, seed=22)])
    lnpost = mod.lnpost_batch(pars)
    # (1) permutation equivariance, bit for bit (no cross-row state, no order-dependent arithmetic)
    perm = np.random.RandomState(0).permutation(n)
    assert np.array_equal(mod.lnpost_batch(pars[perm]), lnpost[perm], equal_nan=True)
    # (2) chunk independence: the pipelined host path (4 chunks) == many small calls == the device-buffer path
    pieces = np.concatenate([mod.lnpost_batch(pars[a:a + 99_991]) for a in range(0, n, 99_991)])
    assert np.array_equal(pieces, lnpost, equal_nan=True)
    # (3) lnpost == lnprior + lnlike where the prior is finite, -inf elsewhere (starmodel.py:538-542)
    post2, prior, like = mod.lnpost_batch(pars, parts=True)
    assert np.array_equal(post2, lnpost, equal_nan=True)
    fin = np.isfinite(prior)
    assert np.array_equal(lnpost[fin], (prior + like)[fin], equal_nan=True)
    assert np.all(np.isneginf(lnpost[~fin]))
    # (4) a 20k-row random sample agrees with the oracle
    idx = np.random.RandomState(1).choice(n, 20_000, replace=False)
    _compare(lnpost[idx], om.lnpost_batch(pars[idx], n_threads=8), ATOL_LNPOST)
    # (5) pinned and pageable host buffers give identical results
    pin = world["ctx"].pinned_empty((n, 5))
    pin[:] = pars
    out = world["ctx"].pinned_empty((n,))
    mod.lnpost_batch(pin, out=out)
    assert np.array_equal(out, lnpost, equal_nan=True)
    # (6) the maximum of lnpost over the posterior-like half sits next to the truth
    best = pars[np.nanargmax(np.where(np.isfinite(lnpost), lnpost, -np.inf))]
    assert abs(best[0] - truth[0]) < 0.2 and abs(best[3] - truth[3]) < 5.0


def test_edge_sizes(world):
    mod, om, truth = _model(world, "track", 1)
    assert mod.lnpost_batch(np.empty((0, 5))).shape == (0,)
    one = mod.lnpost_batch(truth.reshape(1, 5))
    assert one.shape == (1,) and one[0] == mod.lnpost(truth) and np.isfinite(one[0])
    from isochrones_b200 import synthetic as syn

    for n in (1, 31, 255, 257, (1 << 18) - 1, (1 << 18) + 1):       # warp / block / pipeline-chunk boundaries
        pars = syn.posterior_like_batch("track", n, truth, seed=n)
        got = mod.lnpost_batch(pars)
        k = min(n, 2000)
        _compare(got[-k:], om.lnpost_batch(pars[-k:]), ATOL_LNPOST)
    with pytest.raises(ValueError):
        mod.lnpost_batch(np.zeros((4, 6)))
    it = world["ic_track"].model_grid.interp
    assert it([np.array([]), np.array([]), np.array([])], ["Teff"]).shape == (0, 1)


def test_catalog_mode(world):
    """Config 4 in miniature: 64 star models (different observations, band sets and priors) in one launch; row i
    uses model_of_row[i].  Compared with per-model launches and with the oracle's catalog loop."""
    import isochrones_b200 as ib
    from isochrones_b200 import synthetic as syn
    from oracle import oracle

    rng = np.random.RandomState(5)
    ic = world["ic_track"]
    truth = syn.default_truth("track")
    models, oracles = [], []
    all_bands = ["V", "J", "H", "K", "G", "BP", "RP"]
    for s in range(64):
        t = truth * (1 + 0.05 * rng.standard_normal(5))
        t[4] = abs(t[4])
        bands = list(rng.choice(all_bands, size=rng.randint(1, 6), replace=False))
        _, _, _, mags = ic.interp_mag(list(t), bands)
        kw = {b: (float(m), 0.03) for b, m in zip(bands, mags)}
        if s % 2:
            kw["Teff"] = (5800.0 + 10 * s, 90.0)
        if s % 3 == 0:
            kw["parallax"] = (1000.0 / t[3], 0.2)
        mod = ib.BasicStarModel(ic, maxAV=0.5 + 0.01 * s, **kw)
        models.append(mod)
        oracles.append(oracle.StarModel(mod, model_grid=world["og_track"], bc_grid=world["og_bc"]))
    compiled, bands = ib.compile_catalog(models)
    n = 40_000
    mor = rng.randint(0, 64, n).astype(np.int32)
    pars = syn.posterior_like_batch("track", n, truth, seed=6)
    got = compiled.lnpost(pars, model_of_row=mor)
    want = oracle.lnpost_catalog(oracles, mor, pars, n_threads=8)
    _compare(got, want, ATOL_LNPOST)
    for s in (0, 7, 63):
        sel = mor == s
        # same rows through the single-model launch: the band terms are summed in the model's own band order there
        # and in catalog-pack column order here, so agreement is to rounding, not bit for bit
        _compare(models[s].lnpost_batch(pars[sel]), got[sel], 1e-9)


@pytest.mark.parametrize("kind,N", [("track", 1), ("iso", 2)])
def test_custom_priors_generic_profile(world, kind, N):
    """set_prior with non-default classes routes the launch to the generic-prior kernel profile; results must still
    match the oracle (GaussianPrior on feh, bounded GaussianPrior on distance, LogNormal-free Salpeter-like mass)."""
    from isochrones_b200 import priors as P

    mod, _, truth = _model(world, kind, N)
    mod.set_prior(feh=P.GaussianPrior(-0.1, 0.3, bounds=(-2.0, 0.5)), AV=P.FlatPrior((0.0, 0.8)),
                  distance=P.GaussianPrior(100.0, 30.0, bounds=(1.0, 400.0)))
    if kind == "track":
        mod.set_prior(mass=P.SalpeterPrior(bounds=(0.2, 20.0)))
    else:
        mod.set_prior(age=P.GaussianPrior(9.6, 0.5, bounds=(8.0, 10.1)))
    from oracle import oracle

    om = oracle.StarModel(mod, model_grid=world["og_" + kind], bc_grid=world["og_bc"])
    pars = _batches(kind, mod, truth, world["trk" if kind == "track" else "iso"]["axes"], 30_000)
    got = mod.lnpost_batch(pars, parts=True)
    want = om.lnpost_batch(pars, n_threads=8, parts=True)
    _compare(got[1], want[1], 1e-9)
    _compare(got[2], want[2], ATOL_LNPOST)
    _compare(got[0], want[0], ATOL_LNPOST)
    assert np.isfinite(want[0]).sum() > 10_000


def test_star_catalog_compile(world):
    """StarCatalog (table of stars) -> one device model per row; checked against per-row host models + the oracle."""
    import pandas as pd

    from isochrones_b200 import synthetic as syn
    from isochrones_b200.catalog import StarCatalog
    from oracle import oracle

    ic = world["ic_track"]
    rng = np.random.RandomState(3)
    n = 300
    truth = syn.default_truth("track")
    t = np.tile(truth, (n, 1)) * (1 + 0.05 * rng.standard_normal((n, 5)))
    t[:, 4] = np.abs(t[:, 4])
    _, _, _, mags = ic.interp_mag([t[:, j] for j in range(5)], ["V", "J", "K"])
    df = pd.DataFrame({"V_mag": mags[:, 0], "V_mag_unc": 0.02, "J_mag": mags[:, 1], "J_mag_unc": 0.03,
                       "K_mag": mags[:, 2], "K_mag_unc": 0.03, "Teff": 5800.0 + 50 * rng.standard_normal(n), "Teff_unc": 80.0,
                       "parallax": 1000.0 / t[:, 3], "parallax_unc": 0.1})
    df.loc[5, "J_mag"] = np.nan
    df.loc[9, "parallax"] = np.nan
    cat = StarCatalog(df, props=["Teff", "parallax"])
    compiled = cat.compile(ic)
    assert compiled.n_models == n
    rows_per_star = 40
    mor = np.repeat(np.arange(n, dtype=np.int32), rows_per_star)
    pars = np.repeat(t, rows_per_star, axis=0) * (1 + 0.01 * rng.standard_normal((n * rows_per_star, 5)))
    got = compiled.lnpost(pars, model_of_row=mor)
    oms = [oracle.StarModel(m, model_grid=world["og_track"], bc_grid=world["og_bc"]) for m in cat.iter_models(ic)]
    want = oracle.lnpost_catalog(oms, mor, pars, n_threads=8)
    _compare(got, want, ATOL_LNPOST)
    assert np.isfinite(want).mean() > 0.9


def test_concurrent_host_threads_share_a_context(world):
    """ctypes releases the GIL: several host threads may call into one context at once; calls serialise on the
    context's mutex and every thread gets its own correct results."""
    import threading

    from isochrones_b200 import synthetic as syn

    mod, om, truth = _model(world, "track", 1)
    batches = [syn.posterior_like_batch("track", 300_000 + 1000 * t, truth, seed=40 + t) for t in range(4)]
    want = [mod.lnpost_batch(b) for b in batches]
    got = [None] * 4
    errs = []

    def work(t):
        try:
            for _ in range(5):
                got[t] = mod.lnpost_batch(batches[t])
        except Exception as e:      # pragma: no cover
            errs.append(e)

    threads = [threading.Thread(target=work, args=(t,)) for t in range(4)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errs
    for t in range(4):
        assert np.array_equal(got[t], want[t], equal_nan=True)
