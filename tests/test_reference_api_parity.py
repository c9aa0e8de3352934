"""CPU, build container only (needs /root/reference): the product's host classes expose the reference's interface
for the lnpost path — same class names, constructor / method signatures, class constants and attribute surface — so
code written against the reference's API for this path runs against the product unchanged."""
import inspect

import numpy as np
import pytest

from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present (GPU box)")


def _params(fn):
    return [p for p in inspect.signature(fn).parameters if p != "self"]


def test_prior_classes():
    ref = ref_shim.load()
    from isochrones_b200 import priors as P

    for name in ("Prior", "BoundedPrior", "BrokenPrior", "GaussianPrior", "LogNormalPrior", "FlatPrior", "FlatLogPrior",
                 "PowerLawPrior", "FehPrior", "EEP_prior", "AgePrior", "DistancePrior", "AVPrior", "QPrior", "SalpeterPrior",
                 "ChabrierPrior"):
        r, m = getattr(ref.priors, name), getattr(P, name)
        assert [c.__name__ for c in m.__mro__[:-1]] == [c.__name__ for c in r.__mro__[:-1]], name
        assert _params(m.__init__) == _params(r.__init__), name
        for meth in ("lnpdf", "pdf", "__call__", "sample"):
            assert hasattr(m, meth), (name, meth)
            assert _params(getattr(m, meth))[:1] == _params(getattr(r, meth))[:1], (name, meth)
    # default objects carry the same parameters / bounds
    for name in ("AgePrior", "DistancePrior", "AVPrior", "QPrior", "SalpeterPrior", "ChabrierPrior", "FehPrior"):
        r, m = getattr(ref.priors, name)(), getattr(P, name)()
        assert tuple(r.bounds) == tuple(m.bounds), name
        for attr in ("alpha", "halo_fraction", "local", "breakpoints", "n_components"):
            if hasattr(r, attr):
                assert getattr(r, attr) == getattr(m, attr), (name, attr)
        if hasattr(r, "lognorms"):
            assert np.allclose(r.lognorms, m.lognorms, rtol=1e-10, atol=1e-12)


def test_interpolator_and_models():
    ref = ref_shim.load()
    import isochrones_b200 as ib
    from isochrones_b200 import interp as I, models as M

    assert _params(I.DFInterpolator.__init__)[:4] == _params(ref.interp.DFInterpolator.__init__)
    # the reference's parameters first, in order; the product only appends optional keywords (out=)
    assert _params(I.DFInterpolator.__call__)[:2] == _params(ref.interp.DFInterpolator.__call__)
    for meth in ("add_column", "_make_grid"):
        assert _params(getattr(I.DFInterpolator, meth)) == _params(getattr(ref.interp.DFInterpolator, meth))
    for name in ("EvolutionTrackInterpolator", "IsochroneInterpolator"):
        r, m = getattr(ref.models, name), getattr(M, name)
        assert r.param_names == m.param_names and r.eep_replaces == m.eep_replaces
        assert tuple(r._param_index_order) == tuple(m._param_index_order)
    for meth in ("interp_value", "interp_mag", "initialize", "get_eep", "generate", "__call__", "mass", "radius", "Teff",
                 "logg", "feh", "density", "nu_max", "delta_nu"):
        r, m = getattr(ref.models.ModelGridInterpolator, meth), getattr(M.ModelGridInterpolator, meth)
        assert _params(m)[:len(_params(r))] == _params(r) or _params(m) == _params(r), meth
    assert M.ModelGridInterpolator.eep_bounds == (0, 1710)


def test_star_models():
    ref = ref_shim.load()
    from isochrones_b200 import starmodel as S

    r, m = ref.starmodel.BasicStarModel, S.BasicStarModel
    assert _params(m.__init__) == _params(r.__init__)
    for meth in ("lnlike", "lnprior", "bounds", "set_bounds", "set_prior", "prior", "mnest_loglike", "sample_from_prior"):
        assert _params(getattr(m, meth)) == _params(getattr(r, meth)), meth
    assert _params(m.mnest_prior)[:1] == ["cube"] and _params(r.lnpost)[0] == _params(m.lnpost)[0] == "p"
    for prop in ("param_names", "bands", "props", "spec_props", "n_params", "labelstring", "ic"):
        assert isinstance(getattr(m, prop), property) and isinstance(getattr(r, prop), property), prop
    for prop in ("samples", "derived_samples"):     # defined on the reference's StarModel base, inherited by BasicStarModel
        assert isinstance(getattr(m, prop), property) and isinstance(getattr(r, prop), property), prop
    assert m._not_a_band == r._not_a_band
    for name in ("SingleStarModel", "BinaryStarModel", "TripleStarModel"):
        assert issubclass(getattr(S, name), S.BasicStarModel) and issubclass(getattr(ref.starmodel, name), r)


def test_same_objects_on_the_same_small_grid(golden):
    """Build the reference's and the product's model on the same grids: identical parameter names, bands, bounds,
    prior constants and default prior samples' support — everything up to (not including) the GPU evaluation."""
    ref = ref_shim.load()
    from tests.helpers import golden_grids, product_ic

    import isochrones_b200 as ib

    trk, iso, bc = golden_grids(golden["interp"])
    for kind, model, N in (("track", trk, 1), ("iso", iso, 2)):
        mref = dict(model)
        mref["limits"] = {"age": (5, 10.13), "feh": (-4, 0.5), "eep": (0, 60), "mass": (0.1, 300)}
        ric = ref_shim.make_ref_ic(kind, mref, bc, eep_bounds=(0, 60))
        pic = product_ic(kind, model, bc)
        kw = dict(Teff=(5772.0, 80.0), V=(10.0, 0.02), K=(8.0, 0.02), parallax=(8.0, 0.2))
        rm = ref.starmodel.BasicStarModel(ric, N=N, maxAV=0.8, **kw)
        pm = ib.BasicStarModel(pic, N=N, maxAV=0.8, **kw)
        assert tuple(rm.param_names) == tuple(pm.param_names) and rm.bands == pm.bands
        assert [tuple(map(float, rm.bounds(p))) for p in rm.param_names] == [tuple(map(float, pm.bounds(p))) for p in pm.param_names]
        for k in ("mass", "age", "feh", "distance", "AV"):
            assert type(rm._priors[k]).__name__ == type(pm._priors[k]).__name__
            assert np.isclose(rm._priors[k]._norm, pm._priors[k]._norm, rtol=1e-12)
        assert rm._priors["eep"].bounds == pm._priors["eep"].bounds
        assert rm._priors["eep"].orig_par == pm._priors["eep"].orig_par and rm._priors["eep"].deriv_prop == pm._priors["eep"].deriv_prop
