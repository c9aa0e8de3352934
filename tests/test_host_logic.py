"""CPU: host-side logic of the product — prior constants computed at construction (quad normalisations), model
compilation into the C struct, parameter names / bounds bookkeeping, row sharding — against the golden specs the
reference produced (tests/golden) and the constants recorded in SURVEY.md §8a."""
import json

import numpy as np
import pytest

from tests.helpers import golden_grids, load_specs, product_ic, product_model_from_spec, product_prior_cases


def _close(a, b):
    return np.isclose(a, b, rtol=1e-12, atol=0.0)


def test_prior_constants_match_reference(golden):
    specs = json.loads(str(golden["priors"]["pr_specs_json"]))
    cases = product_prior_cases()
    assert set(cases) == set(specs)
    for name, pr in cases.items():
        d = specs[name]
        assert type(pr).__name__ == d["cls"]
        b = getattr(pr, "_bounds", None)
        assert (b is None) == (d["_bounds"] is None)
        if b is not None:
            assert [float(b[0]), float(b[1])] == d["_bounds"]
        assert _close(pr._norm, d["_norm"]), name
        for attr in ("alpha", "mean", "sigma", "norm", "lognorm", "mu", "scale", "log_s", "halo_fraction"):
            if attr in d:
                assert _close(float(getattr(pr, attr)), d[attr]), (name, attr)
        if "components" in d:
            assert np.allclose(pr.norms, d["norms"], rtol=1e-10)
            assert np.allclose(pr.lognorms, d["lognorms"], rtol=1e-10, atol=1e-12)
            assert list(map(float, pr.breakpoints)) == d["breakpoints"]


def test_prior_constants_match_survey():
    from isochrones_b200 import priors as P

    cb = P.ChabrierPrior()
    cb.bounds = (0.1, 300)
    assert np.allclose(cb.lognorms, [-0.82605718, 2.13418178], atol=1e-8)
    fb = P.FehPrior()
    fb.bounds = (-4, 0.5)
    assert np.isclose(fb._norm, 0.9991867187106847, rtol=1e-12)


def test_prior_structs():
    from isochrones_b200 import _lib, priors as P

    s = P.ChabrierPrior().to_struct()
    assert s.self.kind == _lib.ISO_PRIOR_BROKEN and s.n_comp == 2
    assert s.comp[0].kind == _lib.ISO_PRIOR_LOGNORMAL and s.comp[1].kind == _lib.ISO_PRIOR_POWERLAW
    assert s.comp[0].flags == _lib.ISO_PF_HAS_BOUNDS                     # plain Prior with _bounds = (0, inf)
    assert s.comp[1].flags == _lib.ISO_PF_BOUNDED | _lib.ISO_PF_HAS_BOUNDS
    assert (s.comp[1].lo, s.comp[1].hi) == (1.0, 100.0)
    f = P.FehPrior().to_struct()
    assert f.self.kind == _lib.ISO_PRIOR_FEH and f.self.flags == _lib.ISO_PF_LOCAL      # unbounded until bounds set
    g = P.GaussianPrior(9.6, 1.0).to_struct()
    assert g.self.flags == _lib.ISO_PF_BOUNDED                            # BoundedPrior with bounds None

    class Custom(P.Prior):
        def _pdf(self, x):
            return 1.0

    with pytest.raises(TypeError):
        Custom().to_struct()           # no device implementation and no CPU fallback


def test_models_compile_to_structs(golden):
    gi, gl = golden["interp"], golden["lnpost"]
    trk, iso, bc = golden_grids(gi)
    specs = load_specs(gl)
    ics = {"track": product_ic("track", trk, bc), "iso": product_ic("iso", iso, bc)}
    for name, spec in specs.items():
        mod = product_model_from_spec(spec, ics[spec["kind"]])
        assert list(mod.param_names) == spec["param_names"]
        assert mod.n_params == 4 + spec["N"]
        assert list(mod.bands) == spec["bands"]
        assert [tuple(float(v) for v in mod.bounds(p)) for p in mod.param_names] == [tuple(b) for b in spec["bounds"]]
        s = mod.to_struct()
        assert s.n_stars == spec["N"] and s.n_bands == len(spec["bands"])
        assert s.eep_replaces_age == (1 if spec["kind"] == "track" else 0)
        assert list(s.index_order) == ([2, 0, 1, 3, 4] if spec["kind"] == "track" else [1, 2, 0, 3, 4])
        assert bool(s.has_plax) == ("parallax" in spec["kwargs"])
        assert bool(s.has_nu_max) == ("nu_max" in spec["kwargs"])
        for i, b in enumerate(spec["bands"]):
            assert (s.mag_val[i], s.mag_unc[i]) == tuple(spec["kwargs"][b])
        for i, k in enumerate(["Teff", "logg", "feh"]):
            if k in spec["kwargs"]:
                assert s.spec_val[i] == spec["kwargs"][k][0]
            else:
                assert np.isnan(s.spec_val[i])
        # prior objects carry the reference's constants
        for key in ("mass", "age", "feh", "distance", "AV"):
            d = spec["priors"][key]
            pr = mod._priors[key]
            assert type(pr).__name__ == d["cls"], (name, key)
            assert _close(pr._norm, d["_norm"]), (name, key)
            if "lognorms" in d:
                assert np.allclose(pr.lognorms, d["lognorms"], rtol=1e-10, atol=1e-12)
        assert (s.eep_lo, s.eep_hi) == tuple(spec["priors"]["eep"]["_bounds"])


def test_model_validation():
    import isochrones_b200 as ib
    from isochrones_b200 import synthetic as syn

    trk = syn.make_track_grid(n_feh=3, n_mass=6, n_eep=20)
    iso = syn.make_iso_grid(n_age=5, n_feh=3, n_eep=20)
    bc = syn.make_bc_grid(bands=("V", "K"), n_teff=6, n_logg=4, n_feh=3, n_av=3)
    ict = ib.ichrone_from_arrays("track", trk, bc)
    with pytest.raises(ValueError):
        ib.BinaryStarModel(ict, V=(10, 0.02))            # starmodel.py:1396-1397
    ici = ib.ichrone_from_arrays("iso", iso, bc)
    m = ib.TripleStarModel(ici, V=(10, 0.02), K=(float("nan"), 0.02), parallax=(5.0, 0.1))
    assert m.param_names == ("eep_0", "eep_1", "eep_2", "age", "feh", "distance", "AV")
    assert m.bands == ["V"]                               # NaN observations are dropped (starmodel.py:1427-1430)
    assert m.bounds("distance") == (0, 400.0)             # 2000 / parallax (starmodel.py:1468-1472)
    assert m.bounds("eep_2") == m.bounds("eep")
    with pytest.raises(ValueError):
        m.set_bounds(AV=(0, 1, 2))


def test_row_sharder():
    from isochrones_b200.parallel import RowSharder

    for n, world in ((10, 3), (7, 8), (0, 2), (1_000_003, 8), (16, 4)):
        shards = [RowSharder(n, world, r) for r in range(world)]
        cover = np.concatenate([np.arange(*s.bounds()) for s in shards]) if n else np.array([])
        assert np.array_equal(cover, np.arange(n))
        assert max(s.counts[s.rank] for s in shards) - min(s.counts[s.rank] for s in shards) <= 1
        rows = np.arange(n, dtype=float)
        gathered = np.stack([s.padded(s.local(rows) * 2.0) for s in shards]) if world else None
        assert np.array_equal(shards[0].assemble(gathered), rows * 2.0)
    with pytest.raises(ValueError):
        RowSharder(4, 2, 2)


def test_catalog_structs_match_per_model_structs():
    """StarCatalog.build_structs (vectorised) == one BasicStarModel.to_struct per row (the reference's iter_models)."""
    import pandas as pd

    import isochrones_b200 as ib
    from isochrones_b200 import synthetic as syn
    from isochrones_b200.catalog import StarCatalog

    trk = syn.make_track_grid(n_feh=3, n_mass=6, n_eep=20)
    bc = syn.make_bc_grid(bands=("V", "J", "K"), n_teff=6, n_logg=4, n_feh=3, n_av=3)
    ic = ib.ichrone_from_arrays("track", trk, bc)
    rng = np.random.RandomState(0)
    n = 12
    df = pd.DataFrame({
        "V_mag": 10 + rng.rand(n), "V_mag_unc": 0.02 + 0 * rng.rand(n),
        "J_mag": 9 + rng.rand(n), "J_mag_unc": 0.03 + 0 * rng.rand(n),
        "K_mag": 8 + rng.rand(n), "K_mag_unc": 0.03 + 0 * rng.rand(n),
        "Teff": 5000 + 1000 * rng.rand(n), "Teff_unc": 80.0 + 0 * rng.rand(n),
        "parallax": 5 + 5 * rng.rand(n), "parallax_unc": 0.1 + 0 * rng.rand(n),
    })
    df.loc[3, "J_mag"] = np.nan            # band missing for one star
    df.loc[5, "Teff"] = np.nan             # spectroscopy missing
    df.loc[7, "parallax"] = -0.4           # negative parallax -> bound from the uncertainty
    df.loc[9, "parallax"] = np.nan         # no parallax -> default distance bounds
    cat = StarCatalog(df, props=["Teff", "parallax"])
    assert cat.bands == ("V", "J", "K")
    arr, bands = cat.build_structs(ic, maxAV=0.7)
    assert bands == ["V", "J", "K"] and len(arr) == n
    col = {b: i for i, b in enumerate(bands)}
    import ctypes as C

    for i, mod in enumerate(cat.iter_models(ic)):
        mod.set_bounds(AV=(0, 0.7))
        want = mod.to_struct(band_columns=col)
        got = arr[i]
        assert got.n_bands == want.n_bands == len(mod.bands)
        assert list(got.band_col)[:got.n_bands] == list(want.band_col)[:want.n_bands]
        for f in ("mag_val", "mag_unc"):
            assert list(getattr(got, f))[:got.n_bands] == list(getattr(want, f))[:want.n_bands]
        for f in ("spec_val", "spec_unc"):
            assert np.array_equal(np.array(getattr(got, f)), np.array(getattr(want, f)), equal_nan=True)
        assert (got.has_plax, got.has_nu_max) == (want.has_plax, want.has_nu_max)
        if want.has_plax:
            assert (got.plax, got.plax_unc) == (want.plax, want.plax_unc)
        assert (got.distance.self.lo, got.distance.self.hi) == (want.distance.self.lo, want.distance.self.hi), i
        assert (got.AV.self.lo, got.AV.self.hi) == (0.0, 0.7)
        assert bytes(got.mass) == bytes(want.mass) and bytes(got.feh) == bytes(want.feh)
        assert bytes(got.eep_orig) == bytes(want.eep_orig)


def test_load_reference_npz_cache(tmp_path):
    """DFInterpolator.from_npz reads the dense-grid cache format the reference writes (interp.py:611-612)."""
    import itertools

    import pandas as pd

    from isochrones_b200 import DFInterpolator

    x, y, z = np.arange(3.0), np.arange(4.0) * 2, np.arange(5.0) + 1
    index = pd.MultiIndex.from_product((x, y, z), names=["a", "b", "c"])
    df = pd.DataFrame(index=index)
    df["s"] = [a + b + c for a, b, c in itertools.product(x, y, z)]
    df["p"] = [a * b * c for a, b, c in itertools.product(x, y, z)]
    fn = str(tmp_path / "full_grid.npz")
    from oracle import ref_shim

    if ref_shim.available():       # written by the reference itself (build container only)
        ref = ref_shim.load()
        ref.interp.DFInterpolator(df, filename=fn, is_full=True)
    else:                           # same format, written by hand
        np.savez(fn, grid=np.array(df.values, dtype=float).reshape(3, 4, 5, 2), columns=list(df.columns))
    it = DFInterpolator.from_npz(fn, (x, y, z), index_names=["a", "b", "c"])
    mine = DFInterpolator(df, is_full=True)
    assert it.columns == mine.columns == ["s", "p"]
    assert np.array_equal(it.grid, mine.grid) and it.grid.shape == (3, 4, 5, 2)
    assert all(np.array_equal(a, b) for a, b in zip(it.index_columns, mine.index_columns))
    # the product's own cache round trip uses the same format
    fn2 = str(tmp_path / "mine.npz")
    DFInterpolator(df, filename=fn2, is_full=True)
    again = DFInterpolator(df, filename=fn2, is_full=True)
    assert np.array_equal(again.grid, mine.grid)
    with pytest.raises(ValueError):
        DFInterpolator.from_npz(fn, (x, y))
