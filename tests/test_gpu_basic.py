"""GPU: the reference's own smoke / property tests for this path (isochrones/tests/test_basic.py:11-163 and
tests/test_likelihood.py), run against the product on MIST-shaped synthetic grids: every scalar / array broadcast
combination is finite, NaN in -> NaN out, on-grid calls work, `get_eep(accurate=True)` round-trips the mass / age,
spectroscopic likelihoods are finite, and the N = 1 / 2 / 3 models agree with each other where they must."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SHAPE = {"track": dict(n_feh=9, n_mass=60, n_eep=513), "iso": dict(n_age=40, n_feh=9, n_eep=513),
         "bc": dict(n_teff=40, n_logg=14, n_feh=10, n_av=9)}


@pytest.fixture(scope="module")
def ics():
    import isochrones_b200 as ib

    bands = ["J", "H", "K"]
    return {"iso": ib.get_ichrone("mist", bands=bands, synthetic_shape=SHAPE),
            "track": ib.get_ichrone("mist", bands=bands, tracks=True, synthetic_shape=SHAPE)}


def test_get_ichrone(ics):
    import isochrones_b200 as ib

    assert isinstance(ics["iso"], ib.IsochroneInterpolator) and isinstance(ics["track"], ib.EvolutionTrackInterpolator)
    assert ics["iso"].bands == ["J", "H", "K"]
    ics["iso"].initialize([200.0, 9.5, -0.2, 500.0, 0.2])
    ics["track"].initialize([1.0, 150.0, -0.2, 500.0, 0.2])
    with pytest.raises(ValueError):
        ib.get_ichrone("dartmouth")


def test_basic_ic_checks(ics):
    """reference tests/test_basic.py:90-125 (_basic_ic_checks)."""
    ic = ics["iso"]
    age, feh = (9.5, -0.2)
    eep = ic.get_eep(1.0, age, feh, accurate=True)
    o = np.ones(100)
    assert np.isfinite(ic.radius(eep, age, feh))
    for args in ((o * eep, age, feh), (eep, o * age, feh), (eep, age, o * feh), (eep, o * age, o * feh),
                 (o * eep, age, o * feh), (o * eep, o * age, feh), (o * eep, o * age, o * feh)):
        r = ic.radius(*args)
        assert r.shape == (100,) and np.isfinite(r).all()
    assert np.isfinite(ic.Teff(eep, age, feh)) and np.isfinite(ic.Teff(eep, o * age, feh)).all()
    assert np.isfinite(ic.density(eep, age, feh)) and np.isfinite(ic.density(eep, age, o * feh)).all()
    assert np.isfinite(ic.nu_max(eep, age, feh)) and np.isfinite(ic.delta_nu(eep, age, feh))
    _, _, _, mags = ic.interp_mag((eep, age, feh, 500, 0.2), ic.bands)
    assert np.isfinite(mags).all()
    assert len(ic(np.arange(100.0, 200.0), 8.0, 0.0).dropna()) > 0   # on-the-grid call (reference issue #64)
    assert np.isnan(ic.radius(1.0, np.nan, 0.1))                    # NaN in -> NaN out (reference issue #65)
    # the accurate EEP puts the requested initial mass back (test_basic.py:60-76, resid_tol 0.02)
    assert abs(ic.initial_mass(eep, age, feh) - 1.0) < 0.022


def test_basic_ic_checks_tracks(ics):
    """reference tests/test_basic.py:128-157 (_basic_ic_checks_tracks)."""
    ic = ics["track"]
    mass, feh = (1.0, -0.2)
    eep = ic.get_eep(mass, 9.6, feh, accurate=True)
    o = np.ones(100)
    assert np.isfinite(ic.radius(mass, eep, feh))
    for args in ((o * mass, eep, feh), (mass, o * eep, feh), (mass, eep, o * feh), (mass, o * eep, o * feh),
                 (o * mass, eep, o * feh), (o * mass, o * eep, feh), (o * mass, o * eep, o * feh)):
        r = ic.radius(*args)
        assert r.shape == (100,) and np.isfinite(r).all()
    assert np.isfinite(ic.Teff(mass, eep, feh)) and np.isfinite(ic.Teff(mass, o * eep, feh)).all()
    assert np.isfinite(ic.density(mass, eep, feh)) and np.isfinite(ic.density(mass, eep, o * feh)).all()
    assert np.isfinite(ic.nu_max(mass, eep, feh)) and np.isfinite(ic.delta_nu(mass, eep, feh))
    _, _, _, mags = ic.interp_mag((mass, eep, feh, 500, 0.2), ic.bands)
    assert np.isfinite(mags).all()
    assert len(ic(1.0, np.arange(100.0, 200.0), 0.0).dropna()) > 0
    assert np.isnan(ic.radius(1.0, np.nan, 0.1))
    # fast and accurate EEP lookups agree, and both land on the requested age
    fast = ic.get_eep(mass, 9.6, feh)
    assert abs(fast - eep) < 3.0
    assert abs(ic.interp_value([mass, eep, feh], ["age"])[0] - 9.6) < 0.02


def test_closest_eep(ics):
    """reference tests/test_basic.py:60-87 (_check_closest_eep), 60 random stars instead of 10 000."""
    ic = ics["iso"]
    rng = np.random.RandomState(1234)
    n, resid_tol = 60, 0.02
    masses = rng.random_sample(n) * 1.5 + 0.5
    fehs = rng.random_sample(n) * 2.0 - 1.5
    ages = rng.random_sample(n) * 1.2 + 8.6
    n_ok = 0
    for m, a, f in zip(masses, ages, fehs):
        e = ic.get_eep(m, a, f, return_nan=True, accurate=True)
        if not np.isnan(e):
            assert abs(ic.initial_mass(e, a, f) - m) < resid_tol * 1.1
            n_ok += 1
    assert n_ok > n // 3
    # the batched form solves all stars in three launches and agrees with the one-by-one calls
    batch = ic.get_eep(masses, ages, fehs, accurate=True)
    one = np.array([ic.get_eep(m, a, f, accurate=True) for m, a, f in zip(masses, ages, fehs)])
    assert np.array_equal(batch, one, equal_nan=True)
    ok = ~np.isnan(batch)
    assert np.all(np.abs(ic.initial_mass(batch[ok], ages[ok], fehs[ok]) - masses[ok]) < 1e-3)


def test_spec_likelihoods(ics):
    """reference tests/test_basic.py:122-125, 160-163 (_check_spec / _check_spec_tracks) on BasicStarModel."""
    import isochrones_b200 as ib

    for kind, ic in ics.items():
        mod = ib.BasicStarModel(ic, Teff=(5700, 100), logg=(4.5, 0.1), feh=(0.0, 0.2))
        eep = ic.get_eep(1.0, 9.6, 0.1, accurate=True)
        pars = [eep, 9.6, 0.1, 200, 0.2] if kind == "iso" else [1.0, eep, 0.1, 200, 0.2]
        assert np.isfinite(mod.lnlike(pars)) and np.isfinite(mod.lnprior(pars)) and np.isfinite(mod.lnpost(pars))


@pytest.mark.parametrize("props", [dict(Teff=(5800, 100), logg=(4.5, 0.1), J=(3.58, 0.05), K=(3.22, 0.05), parallax=(100, 0.1)),
                                   dict(Teff=(5800, 100), logg=(4.5, 0.1), parallax=(100, 0.1)),
                                   dict(J=(3.58, 0.05), K=(3.22, 0.05), parallax=(100, 0.1))])
def test_multiplicity_consistency(ics, props):
    """reference tests/test_likelihood.py:14-57 compares its two model classes for N = 1, 2, 3; the product has one
    class, so the cross-checks are the ones that must hold between multiplicities: the spectroscopic terms use the primary only,
    lnpost = lnprior + lnlike, and the ordering rule of lnprior is the reference's (starmodel.py:1618-1623)."""
    import isochrones_b200 as ib

    ic = ics["iso"]
    m1, m2, m3 = (ib.BasicStarModel(ic, N=n, **props) for n in (1, 2, 3))
    s = 513.0 / 1710.0
    e0, e1, e2 = 860 * s, 800 * s, 740 * s          # a populated stretch of the synthetic 9.5-dex isochrone
    p1 = [e0, 9.5, 0.01, 10.0, 0.1]
    p2 = [e0, e1, 9.5, 0.01, 10.0, 0.1]
    p3 = [e0, e1, e2, 9.5, 0.01, 10.0, 0.1]
    for m, p in ((m1, p1), (m2, p2), (m3, p3)):
        lp, ll, post = m.lnprior(p), m.lnlike(p), m.lnpost(p)
        assert np.isfinite(lp) and np.isfinite(ll) and np.isclose(post, lp + ll, rtol=0, atol=1e-9)
    assert m2.lnprior([e1, e0, 9.5, 0.01, 10.0, 0.1]) == -np.inf          # eep_1 > eep_0
    assert m2.lnpost([e1, e0, 9.5, 0.01, 10.0, 0.1]) == -np.inf
    # N = 3 rule as written: `not (p0 > p1) and (p1 > p2)`
    assert m3.lnprior([e1, e0, e2, 9.5, 0.01, 10.0, 0.1]) == -np.inf
    assert np.isfinite(m3.lnprior([e0, e2, e1, 9.5, 0.01, 10.0, 0.1]))
    if "J" not in props:          # no photometry: companions cannot change the likelihood at all
        assert m2.lnlike(p2) == m1.lnlike(p1) and m3.lnlike(p3) == m1.lnlike(p1)
    else:                         # an identical twin brightens every band by 2.5 log10(2): the likelihood must change
        assert m2.lnlike([e0, e0, 9.5, 0.01, 10.0, 0.1]) != m1.lnlike(p1)
