"""TEST INFRASTRUCTURE — generate tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python oracle/make_golden.py

Imports the reference in-process (oracle/ref_shim.py), evaluates its own numba/pandas CPU path
(DFInterpolator, interp_mags, BasicStarModel.lnprior/lnlike/lnpost, priors) on small synthetic
MIST-shaped grids and stores inputs, grids and outputs.  The committed vectors pin

  * the C oracle (tests/test_oracle_vs_golden.py, CPU), and
  * the CUDA path (tests/test_gpu_golden.py, GPU)

against the reference itself.  Grids are stored inside the files so the vectors are independent of
the numpy build that regenerates synthetic grids elsewhere.
"""
import itertools
import json
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from isochrones_b200 import synthetic as syn  # noqa: E402
from oracle import ref_shim  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def prior_to_dict(obj):
    """Serialise a reference prior object (attribute names of priors.py)."""
    d = {"cls": type(obj).__name__, "mro": [c.__name__ for c in type(obj).__mro__]}
    b = getattr(obj, "_bounds", None)
    d["_bounds"] = None if b is None else [float(b[0]), float(b[1])]
    d["_norm"] = float(getattr(obj, "_norm", 1.0))
    for attr in ("alpha", "mean", "sigma", "norm", "lognorm", "mu", "scale", "log_s", "halo_fraction"):
        if hasattr(obj, attr):
            d[attr] = float(getattr(obj, attr))
    if hasattr(obj, "local"):
        d["local"] = bool(obj.local)
    if hasattr(obj, "components"):
        d["components"] = [prior_to_dict(c) for c in obj.components]
        d["n_components"] = int(obj.n_components)
        d["breakpoints"] = [float(x) for x in obj.breakpoints]
        d["norms"] = [float(x) for x in obj.norms]
        d["lognorms"] = [float(x) for x in obj.lognorms]
    if hasattr(obj, "orig_prior"):
        d["orig_prior"] = prior_to_dict(obj.orig_prior)
        d["orig_par"] = obj.orig_par
        d["deriv_prop"] = obj.deriv_prop
    return d


def grids_small():
    trk = syn.make_track_grid(n_feh=4, n_mass=12, n_eep=60)
    iso = syn.make_iso_grid(n_age=10, n_feh=4, n_eep=60)
    bc = syn.make_bc_grid(bands=("V", "J", "H", "K", "G"), n_teff=14, n_logg=8, n_feh=6, n_av=5)
    return trk, iso, bc


def pack_grid(prefix, g, out):
    out[prefix + "_grid"] = g["grid"]
    for i, a in enumerate(g["axes"]):
        out["%s_ax%d" % (prefix, i)] = np.asarray(a, dtype=float)
    out[prefix + "_columns"] = np.array(g["columns"])


def ub_mask(axes, coords):
    """Rows for which the reference's unchecked corner indexing reads beyond the END of its array.

    An exact hit on the last node of an axis makes the reference touch ``index + 1 == len(axis)``
    (interp.py:266-291, weight 0).  On a trailing axis that lands on the next row of the flat array
    (deterministic, reproduced by the oracle); when the flat offset passes the end of the whole array
    it is undefined behaviour (heap garbage * 0), so such rows cannot be golden.
    """
    coords = np.asarray(coords, dtype=float)
    n_nodes = int(np.prod([len(a) for a in axes]))
    flat = np.zeros(len(coords), dtype=np.int64)
    inb = np.ones(len(coords), dtype=bool)
    for d, ax in enumerate(axes):
        x = coords[:, d]
        inb &= (x >= ax[0]) & (x <= ax[-1])
        xs = np.where(np.isfinite(x), x, ax[0])
        idx = np.clip(np.searchsorted(ax, xs, side="right") - 1, 0, len(ax) - 1)
        flat = flat * len(ax) + (idx + 1)
    return inb & (flat >= n_nodes)


def fix_ub(axes, coords):
    """Move the leading coordinate of undefined-behaviour rows into the first cell (in place)."""
    m = ub_mask(axes, coords)
    coords[m, 0] = axes[0][0] + 0.37 * (axes[0][1] - axes[0][0])
    assert not ub_mask(axes, coords).any()
    return int(m.sum())


def interp_points(axes, n, seed):
    """50 % uniform, 25 % on nodes (incl. edges), 25 % out-of-bounds / NaN."""
    rng = np.random.RandomState(seed)
    nd = len(axes)
    pts = np.empty((n, nd))
    for d, ax in enumerate(axes):
        pts[:, d] = ax[0] + (ax[-1] - ax[0]) * rng.random_sample(n)
    q = n // 4
    for i in range(2 * q, 3 * q):           # on nodes
        for d, ax in enumerate(axes):
            if rng.rand() < 0.7:
                pts[i, d] = rng.choice(ax)
            if rng.rand() < 0.15:
                pts[i, d] = ax[0]
            elif rng.rand() < 0.15:
                pts[i, d] = ax[-1]
    for i in range(3 * q, n):               # OOB / NaN
        d = rng.randint(nd)
        r = rng.rand()
        span = axes[d][-1] - axes[d][0]
        if r < 0.4:
            pts[i, d] = axes[d][0] - 1e-9 - 0.1 * span * rng.rand()
        elif r < 0.8:
            pts[i, d] = axes[d][-1] + 1e-9 + 0.1 * span * rng.rand()
        else:
            pts[i, d] = np.nan
    fix_ub(axes, pts)
    return pts


def golden_interp(ref, out):
    trk, iso, bc = grids_small()
    pack_grid("trk", trk, out)
    pack_grid("iso", iso, out)
    pack_grid("bc", bc, out)

    # --- 3-D: config 1 (1024 points -> Teff / logg / radius) on the track grid
    it = ref_shim.make_ref_interp(trk["grid"], trk["axes"], trk["columns"])
    pts = interp_points(trk["axes"], 1024, seed=1)
    cols = ["Teff", "logg", "radius"]
    out["i3_pts"] = pts
    out["i3_cols"] = np.array(cols)
    out["i3_vals"] = it([pts[:, 0], pts[:, 1], pts[:, 2]], cols)
    out["i3_vals_all"] = it([pts[:64, 0], pts[:64, 1], pts[:64, 2]])            # cols="all"
    out["i3_scalar"] = np.array([it([float(a), float(b), float(c)], cols) for a, b, c in pts[:32]])

    # --- 3-D on the isochrone grid (different NaN structure)
    ii = ref_shim.make_ref_interp(iso["grid"], iso["axes"], iso["columns"])
    pts = interp_points(iso["axes"], 512, seed=11)
    out["i3iso_pts"] = pts
    out["i3iso_cols"] = np.array(["mass", "Teff", "dm_deep"])
    out["i3iso_vals"] = ii([pts[:, 0], pts[:, 1], pts[:, 2]], ["mass", "Teff", "dm_deep"])

    # --- 4-D: BC grid
    ib = ref_shim.make_ref_interp(bc["grid"], bc["axes"], bc["columns"])
    pts = interp_points(bc["axes"], 512, seed=12)
    out["i4_pts"] = pts
    out["i4_vals"] = ib([pts[:, 0], pts[:, 1], pts[:, 2], pts[:, 3]])
    out["i4_vals_sub"] = ib([pts[:, 0], pts[:, 1], pts[:, 2], pts[:, 3]], ["K", "V"])

    # --- 2-D: the docs/interpolate.ipynb toy frame through the reference's own DataFrame constructor
    import pandas as pd

    x = np.arange(1, 4)
    y = np.arange(1, 6)
    index = pd.MultiIndex.from_product((x, y), names=["x", "y"])
    df = pd.DataFrame(index=index)
    df["sum"] = [a + b for a, b in itertools.product(x, y)]
    df["product"] = [a * b for a, b in itertools.product(x, y)]
    df["power"] = [a ** b for a, b in itertools.product(x, y)]
    i2 = ref.interp.DFInterpolator(df)
    out["i2_grid"] = i2.grid
    out["i2_ax0"], out["i2_ax1"] = i2.index_columns
    out["i2_a"] = i2([1.4, 2.1])
    out["i2_b"] = i2([2.2, 4.6], ["product"])
    i2m = ref.interp.DFInterpolator(df.drop([(3, 3), (3, 4)]))
    out["i2m_grid"] = i2m.grid
    out["i2m_a"] = i2m([1.3, 2.2])
    out["i2m_b"] = i2m([2.3, 3])
    rng = np.random.RandomState(5)
    pts = np.column_stack([1 + 2 * rng.random_sample(64), 1 + 4 * rng.random_sample(64)])
    out["i2_pts"] = pts
    out["i2_vals"] = i2([pts[:, 0], pts[:, 1]])

    # --- the reference's tests/test_interp.py grid (10 x 11 x 12, x^2 cos(y/10) + z)
    xx, yy, zz = [np.arange(10 + np.log10(n)) * n for n in [1, 10, 100]]
    g = np.array([[[a ** 2 * np.cos(b / 10) + c for c in zz] for b in yy] for a in xx])[..., None]
    itst = ref_shim.make_ref_interp(g, (xx, yy, zz), ["val"])
    rng = np.random.RandomState(6)
    pts = rng.random_sample((10, 3)) * 9
    pts[:, 1] *= 10
    pts[:, 2] *= 100
    out["tst_pts"] = pts
    out["tst_vals"] = itst([pts[:, 0], pts[:, 1], pts[:, 2]], ["val"])
    out["tst_node"] = itst([6.0, 50.0, 200.0], ["val"])
    out["tst_pt"] = itst([3.1, 44.0, 503.0], ["val"])


def golden_mags(ref, out):
    trk, iso, bc = grids_small()
    for kind, model in (("track", trk), ("iso", iso)):
        ic = ref_shim.make_ref_ic(kind, model, bc, eep_bounds=(0, 60))
        truth = syn.default_truth(kind, n_eep=60)
        bounds = [(0.1, 300), (0, 60), (-4, 0.5), (0, 2000), (0, 1)] if kind == "track" else \
                 [(0, 60), (5, 10.13), (-4, 0.5), (0, 2000), (0, 1)]
        p = np.concatenate([
            syn.posterior_like_batch(kind, 192, truth, n_eep=60, seed=21),
            syn.prior_like_batch(kind, 192, bounds, seed=22),
            syn.edge_batch(kind, 128, truth, model["axes"], bounds, seed=23),
        ])
        io = ic.param_index_order
        coords = p[:, io[:3]].copy()
        fix_ub(model["axes"], coords)
        p[:, io[0]] = coords[:, 0]
        for bands in (["V", "J", "H", "K"], ["G"], ["K", "V", "G", "J", "H"]):
            teff, logg, feh, mags = ic.interp_mag([p[:, j] for j in range(5)], bands)
            tag = "%s_%s" % (kind, "".join(bands))
            out["m_%s_teff" % tag], out["m_%s_logg" % tag], out["m_%s_feh" % tag] = teff, logg, feh
            out["m_%s_mags" % tag] = mags
        out["m_%s_pars" % kind] = p
        # scalar signature
        out["m_%s_scalar" % kind] = np.array(
            [np.concatenate([[t, g, f], m]) for t, g, f, m in (ic.interp_mag(list(row), ["V", "K"]) for row in p[:16])]
        )


def model_cases():
    """(name, kind, N, obs kwargs builder, init kwargs, post-construction tweaks)."""
    spec = dict(Teff=(5772.0, 80.0), logg=(4.44, 0.1), feh=(0.0, 0.1))
    plax = dict(parallax=(10.0, 0.1))
    return [
        ("track_full", "track", 1, dict(spec=spec, bands=["V", "J", "H", "K"], extra=plax), {}, {}),
        ("track_phot", "track", 1, dict(spec={}, bands=["J", "H", "K"], extra={}), {}, {}),
        ("track_spec", "track", 1, dict(spec=spec, bands=[], extra=plax), {"maxAV": 0.6}, {}),
        ("track_seis", "track", 1, dict(spec=dict(Teff=(5772.0, 80.0)), bands=["G"],
                                        extra=dict(nu_max=(3000.0, 50.0), delta_nu=(130.0, 2.0), parallax=(10.0, 0.1))),
         {"halo_fraction": 0.05}, {}),
        ("track_eepb", "track", 1, dict(spec=spec, bands=["V", "K"], extra=dict(parallax=(-1.0, 0.5))),
         {"eep_bounds": (5, 40), "max_distance": 500.0}, {}),
        ("iso_single", "iso", 1, dict(spec=spec, bands=["V", "J", "H", "K"], extra=plax), {}, {}),
        ("iso_binary", "iso", 2, dict(spec=spec, bands=["V", "J", "H", "K"], extra=plax), {}, {}),
        ("iso_triple", "iso", 3, dict(spec=dict(logg=(4.44, 0.1)), bands=["V", "J", "H", "K", "G"], extra=plax), {}, {}),
        ("iso_binary_prior", "iso", 2, dict(spec={}, bands=["J", "H", "K", "G"], extra=dict(parallax=(2.0, 0.05))),
         {}, {"set_bounds": {"eep": (1, 50), "age": (8, 10)}, "gauss_age": (9.6, 1.0, (8, 10))}),
    ]


def golden_lnpost(ref, out):
    trk, iso, bc = grids_small()
    specs = {}
    for name, kind, N, obs, init_kw, tweaks in model_cases():
        model = trk if kind == "track" else iso
        ic = ref_shim.make_ref_ic(kind, model, bc, eep_bounds=(0, 60))
        truth = syn.default_truth(kind, n_eep=60, n_stars=N)
        # observed magnitudes = model magnitudes of the truth (+ fixed offsets), SURVEY §8d
        kwargs = dict(obs["spec"])
        if obs["bands"]:
            prim = [truth[0]] + list(truth[N:]) if kind == "iso" else list(truth)
            _, _, _, mags = ic.interp_mag(prim, obs["bands"])
            for i, b in enumerate(obs["bands"]):
                kwargs[b] = (float(np.round(mags[i], 3) + 0.01 * (i - 1)), 0.02)
        kwargs.update(obs["extra"])
        mod = ref.starmodel.BasicStarModel(ic, N=N, **init_kw, **kwargs)
        if "set_bounds" in tweaks:
            mod.set_bounds(**tweaks["set_bounds"])
        if "gauss_age" in tweaks:
            mean, sig, b = tweaks["gauss_age"]
            mod.set_prior(age=ref.priors.GaussianPrior(mean, sig, bounds=b))
        bounds = [tuple(float(v) for v in mod.bounds(p)) for p in mod.param_names]
        p = np.concatenate([
            syn.posterior_like_batch(kind, 96, truth, n_eep=60, seed=31),
            syn.prior_like_batch(kind, 96, bounds, seed=32),
            syn.edge_batch(kind, 64, truth, model["axes"], bounds, seed=33),
        ])
        lead = {"track": 2, "iso": N}[kind]          # parameter that maps to the leading model axis
        for _ in range(2):
            for k in range(N):
                cols = [2, 0, 1] if kind == "track" else [N, N + 1, k]
                coords = p[:, cols].copy()
                fix_ub(model["axes"], coords)
                p[:, lead] = coords[:, 0]
        lnprior = np.empty(len(p))
        lnlike = np.empty(len(p))
        lnpost = np.empty(len(p))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for i, row in enumerate(p):
                lnprior[i] = mod.lnprior(row)
                lnlike[i] = mod.lnlike(row)
                lnpost[i] = mod.lnpost(row)
        out["lp_%s_pars" % name] = p
        out["lp_%s_lnprior" % name] = lnprior
        out["lp_%s_lnlike" % name] = lnlike
        out["lp_%s_lnpost" % name] = lnpost
        cube = np.random.RandomState(41).random_sample((8, len(bounds)))
        phys = cube.copy()
        for row in phys:
            mod.mnest_prior(row, len(bounds), len(bounds))
        out["lp_%s_cube" % name] = cube
        out["lp_%s_cube_phys" % name] = phys
        specs[name] = {
            "kind": kind, "N": N, "kwargs": {k: [float(v[0]), float(v[1])] for k, v in kwargs.items()},
            "init_kwargs": {k: (list(v) if isinstance(v, tuple) else v) for k, v in init_kw.items()},
            "tweaks": tweaks, "eep_bounds": [0, 60], "bounds": bounds, "param_names": list(mod.param_names),
            "bands": list(mod.bands),
            "priors": {k: prior_to_dict(v) for k, v in mod._priors.items()},
            "n_finite": int(np.isfinite(lnpost).sum()), "n_nan": int(np.isnan(lnpost).sum()),
        }
        print("%-18s finite %3d  nan %3d  -inf %3d" % (name, specs[name]["n_finite"], specs[name]["n_nan"],
                                                       int(np.isneginf(lnpost).sum())))
    out["lp_specs_json"] = np.array(json.dumps(specs))


def golden_priors(ref, out):
    P = ref.priors
    cases = {
        "flat": P.FlatPrior((0.5, 2.5)),
        "av": P.AVPrior(),
        "flatlog": P.FlatLogPrior((6.0, 10.0)),
        "age": P.AgePrior(),
        "powerlaw": P.PowerLawPrior(-1.7, (0.2, 30.0)),
        "distance": P.DistancePrior(),
        "distance500": P.DistancePrior(max_distance=500),
        "salpeter": P.SalpeterPrior(),
        "q": P.QPrior(),
        "gauss": P.GaussianPrior(9.6, 1.0),
        "gauss_b": P.GaussianPrior(9.6, 1.0, bounds=(8, 10)),
        "lognormal": P.LogNormalPrior(np.log(0.079), 0.69 * np.log(10)),
        "feh": P.FehPrior(),
        "feh_halo": P.FehPrior(halo_fraction=0.3),
        "feh_nonlocal": P.FehPrior(local=False),
        "chabrier": P.ChabrierPrior(),
    }
    fb = P.FehPrior()
    fb.bounds = (-4, 0.5)
    cases["feh_bounded"] = fb
    cb = P.ChabrierPrior()
    cb.bounds = (0.1, 300)
    cases["chabrier_bounded"] = cb
    ab = P.AgePrior()
    ab.bounds = (5, 10.13)
    cases["age_bounded"] = ab
    xs = np.concatenate([
        np.linspace(-5, 12, 69), np.array([0.0, 1.0, 0.1, 100.0, 100.5, 300.0, 1e4, 1e4 + 1, -0.0, np.nan, 0.5, 10.13,
                                            10.15, 5.0, 0.25, -4.0]),
        np.logspace(-3, 4.2, 40),
    ])
    specs = {}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for name, pr in cases.items():
            lnpdf = np.array([pr.lnpdf(float(x)) for x in xs], dtype=float)
            if name == "gauss":          # BoundedPrior with bounds None cannot be called (priors.py:54-56)
                call = np.full(len(xs), np.nan)
            else:
                call = np.array([pr(float(x)) for x in xs], dtype=float)
            out["pr_%s_lnpdf" % name] = lnpdf
            out["pr_%s_call" % name] = call
            specs[name] = prior_to_dict(pr)
    out["pr_x"] = xs
    out["pr_specs_json"] = np.array(json.dumps(specs))



def track_age_arrays(trk):
    """(arrays[n_feh * n_mass, n_eep], weight_arrays, lengths) as StellarModelGrid.get_array_grids builds them
    (models.py:171-205): the leading non-NaN run of each track's age / dt_deep column."""
    g = trk["grid"]
    cols = trk["columns"]
    age = g[..., cols.index("age")].reshape(-1, g.shape[2])
    dt = g[..., cols.index("dt_deep")].reshape(-1, g.shape[2])
    lengths = np.array([int(np.argmax(np.isnan(a))) if np.isnan(a).any() else len(a) for a in age], dtype=np.int64)
    arrays = np.full_like(age, np.nan)
    weights = np.full_like(dt, np.nan)
    for i, n in enumerate(lengths):
        arrays[i, :n] = age[i, :n]
        weights[i, :n] = dt[i, :n]
    return arrays, weights, lengths


def golden_eep(ref, out):
    """interp_eeps (interp.py:488-558): (age, feh, mass) -> EEP on the small and on a mid-size track grid."""
    for tag, trk in (("small", grids_small()[0]), ("mid", syn.make_track_grid(n_feh=7, n_mass=40, n_eep=342))):
        fehs, masses, eeps = trk["axes"]
        arrays, weights, lengths = track_age_arrays(trk)
        rng = np.random.RandomState(71)
        n = 1500
        x0 = fehs[0] + (fehs[-1] - fehs[0]) * rng.random_sample(n)
        x1 = np.exp(np.log(masses[0]) + (np.log(masses[-1]) - np.log(masses[0])) * rng.random_sample(n))
        x = 5.0 + 5.6 * rng.random_sample(n)
        q = n // 5
        x0[:q] = rng.choice(fehs[:-1], q)                 # exact [Fe/H] nodes (not the last: reference UB)
        x1[q:2 * q] = rng.choice(masses, q)               # exact mass nodes, the last one included (flat-offset read)
        x[2 * q:3 * q] = arrays[rng.randint(0, len(arrays), q), rng.randint(0, 5, q)]   # exact ages of some track
        x1[3 * q:3 * q + 20] = masses[-1] + 1.0           # out of bounds
        x0[3 * q + 20:3 * q + 40] = fehs[0] - 0.1
        x[3 * q + 40:3 * q + 60] = np.nan
        x0[x0 >= fehs[-1]] = fehs[-2]
        # mass on its last node makes the reference read track (i0 + 1) * n1 + n1: beyond the array in the last [Fe/H] cell
        x0[(x1 >= masses[-1]) & (x0 >= fehs[-2])] = fehs[0] + 0.3 * (fehs[1] - fehs[0])
        res = ref.interp.interp_eeps(x, x0, x1, fehs, masses, len(masses), arrays, weights, lengths)
        one = np.array([ref.interp.interp_eep(float(a), float(b), float(c), fehs, masses, len(masses), arrays, weights, lengths)
                        for a, b, c in zip(x[:16], x0[:16], x1[:16])])
        assert np.array_equal(one, res[:16], equal_nan=True)
        out[tag + "_age"], out[tag + "_feh"], out[tag + "_mass"], out[tag + "_eep"] = x, x0, x1, res
        out[tag + "_lengths"] = lengths
        if tag == "small":
            pack_grid("trk", trk, out)
        else:
            out["mid_shape"] = np.array([7, 40, 342])
        print("eep", tag, "finite", int(np.isfinite(res).sum()), "nan", int(np.isnan(res).sum()))
def golden_derived(ref, out):
    """The reference's derived-sample table (``BasicStarModel.derived_samples``, starmodel.py:1646-1714): equal-weight
    samples written where the reference expects MultiNest's ``post_equal_weights.dat``, read back and expanded by its own
    ``_make_samples`` — single star on both grid kinds, a binary and a triple."""
    import tempfile

    trk, iso, bc = grids_small()
    cases = {c[0]: c for c in model_cases()}
    specs = {}
    for name in ("track_full", "iso_single", "iso_binary", "iso_triple"):
        _, kind, N, obs, init_kw, _ = cases[name]
        model = trk if kind == "track" else iso
        ic = ref_shim.make_ref_ic(kind, model, bc, eep_bounds=(0, 60))
        truth = syn.default_truth(kind, n_eep=60, n_stars=N)
        kwargs = dict(obs["spec"])
        prim = [truth[0]] + list(truth[N:]) if kind == "iso" else list(truth)
        _, _, _, mags = ic.interp_mag(prim, obs["bands"])
        for i, b in enumerate(obs["bands"]):
            kwargs[b] = (float(np.round(mags[i], 3) + 0.01 * (i - 1)), 0.02)
        kwargs.update(obs["extra"])
        mod = ref.starmodel.BasicStarModel(ic, N=N, **init_kw, **kwargs)
        rows = syn.posterior_like_batch(kind, 48, truth, n_eep=60, seed=51)
        rows[5, 0 if kind == "iso" else 1] = 1e4                      # one sample outside the grid: NaN row
        lnprob = np.linspace(-30.0, -20.0, len(rows))
        with tempfile.TemporaryDirectory() as d, warnings.catch_warnings():
            warnings.simplefilter("ignore")
            mod.mnest_basename = os.path.join(d, "fit-")
            np.savetxt(mod.mnest_basename + "post_equal_weights.dat", np.column_stack([rows, lnprob]), fmt="%.17e")
            # the installed pandas no longer knows read_csv(delim_whitespace=True) (starmodel.py:1656): same meaning, new
            # spelling — a shim of the harness like the stubbed third-party modules, the reference file is untouched
            import pandas as pd
            real_read_csv = pd.read_csv

            def read_csv(*a, **kw):
                if kw.pop("delim_whitespace", False):
                    kw["sep"] = r"\s+"
                return real_read_csv(*a, **kw)
            pd.read_csv = read_csv
            try:
                table = mod.derived_samples
                sam = mod.samples
            finally:
                pd.read_csv = real_read_csv
        # pandas' default float parser is not round-trip exact: the samples the reference actually expanded are the ones it
        # read back
        assert np.allclose(sam[list(mod.param_names)].values, rows, rtol=1e-14, atol=0)
        rows = sam[list(mod.param_names)].values.astype(float)
        lnprob = sam["lnprob"].values.astype(float)
        out["dv_%s_rows" % name] = rows
        out["dv_%s_lnprob" % name] = lnprob
        out["dv_%s_values" % name] = table.values.astype(float)
        specs[name] = {"kind": kind, "N": N, "kwargs": {k: [float(v[0]), float(v[1])] for k, v in kwargs.items()},
                       "init_kwargs": init_kw, "columns": [str(c) for c in table.columns],
                       "param_names": list(mod.param_names), "bands": list(mod.bands)}
        print("%-12s %d columns, %d NaN cells" % (name, table.shape[1], int(np.isnan(table.values.astype(float)).sum())))
    out["dv_specs_json"] = np.array(json.dumps(specs))


def main():
    ref = ref_shim.load()

    os.makedirs(OUT, exist_ok=True)
    only = sys.argv[1:]
    for name, fn in (("interp", golden_interp), ("mags", golden_mags), ("lnpost", golden_lnpost),
                     ("priors", golden_priors), ("eep", golden_eep), ("derived", golden_derived)):
        if only and name not in only:
            continue
        out = {}
        fn(ref, out)
        path = os.path.join(OUT, "golden_%s.npz" % name)
        np.savez_compressed(path, **out)
        print("wrote", path, "%.1f kB" % (os.path.getsize(path) / 1e3))


if __name__ == "__main__":
    main()
