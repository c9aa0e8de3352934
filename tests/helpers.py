"""Shared test helpers: rebuild duck-typed model / prior objects from the specs stored in tests/golden."""
import json
import types

import numpy as np


def prior_from_dict(d):
    """Namespace object carrying the reference's attribute names; class name preserved for the oracle."""
    obj = type(d["cls"], (), {})()
    obj._mro_names = list(d.get("mro", []))     # the oracle converter resolves subclasses through these
    obj._bounds = None if d["_bounds"] is None else tuple(d["_bounds"])
    obj._norm = d["_norm"]
    for k in ("alpha", "mean", "sigma", "norm", "lognorm", "mu", "scale", "log_s", "halo_fraction", "local"):
        if k in d:
            setattr(obj, k, d[k])
    if "components" in d:
        obj.components = [prior_from_dict(c) for c in d["components"]]
        obj.n_components = d["n_components"]
        obj.breakpoints = d["breakpoints"]
        obj.norms = np.array(d["norms"])
        obj.lognorms = np.array(d["lognorms"])
    if "orig_prior" in d:
        obj.orig_prior = prior_from_dict(d["orig_prior"])
        obj.orig_par = d["orig_par"]
        obj.deriv_prop = d["deriv_prop"]
    return obj


class ArrayInterp:
    def __init__(self, grid, axes, columns):
        self.grid = np.ascontiguousarray(grid, dtype=float)
        self.index_columns = tuple(np.asarray(a, dtype=float) for a in axes)
        self.columns = [str(c) for c in columns]
        self.column_index = {c: i for i, c in enumerate(self.columns)}
        self.ndim = len(self.index_columns)
        self.n_columns = len(self.columns)


def golden_grids(gi):
    """(track, iso, bc) dicts as produced by isochrones_b200.synthetic from the golden interp file."""
    out = {}
    for prefix, nd in (("trk", 3), ("iso", 3), ("bc", 4)):
        out[prefix] = {
            "grid": gi[prefix + "_grid"],
            "axes": tuple(gi["%s_ax%d" % (prefix, i)] for i in range(nd)),
            "columns": [str(c) for c in gi[prefix + "_columns"]],
        }
    return out["trk"], out["iso"], out["bc"]


def ns_model_from_spec(spec, model, bc):
    """Duck-typed BasicStarModel stand-in (attributes read by oracle.StarModel) from a golden spec."""
    kind = spec["kind"]
    ic = types.SimpleNamespace(
        param_index_order=[2, 0, 1, 3, 4] if kind == "track" else [1, 2, 0, 3, 4],
        eep_replaces="age" if kind == "track" else "mass",
        model_grid=types.SimpleNamespace(interp=ArrayInterp(model["grid"], model["axes"], model["columns"])),
        bc_grid=types.SimpleNamespace(interp=ArrayInterp(bc["grid"], bc["axes"], bc["columns"]), bands=bc["columns"]),
    )
    kwargs = {k: (np.float64(v[0]), np.float64(v[1])) for k, v in spec["kwargs"].items()}
    mod = types.SimpleNamespace(
        ic=ic, N=spec["N"], kwargs=kwargs, bands=list(spec["bands"]),
        spec_props=[kwargs.get(k, (np.nan, np.nan)) for k in ["Teff", "logg", "feh"]],
        _priors={k: prior_from_dict(v) for k, v in spec["priors"].items()},
        param_names=tuple(spec["param_names"]),
    )
    return mod


def load_specs(gl):
    return json.loads(str(gl["lp_specs_json"]))


def assert_same_special(a, b):
    """NaN / +-inf patterns must be identical."""
    a = np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    assert a.shape == b.shape
    assert np.array_equal(np.isnan(a), np.isnan(b)), "NaN pattern differs"
    assert np.array_equal(np.isposinf(a), np.isposinf(b)), "+inf pattern differs"
    assert np.array_equal(np.isneginf(a), np.isneginf(b)), "-inf pattern differs"


def max_rel_err(a, b):
    a = np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    m = np.isfinite(a) & np.isfinite(b)
    if not m.any():
        return 0.0
    return float(np.max(np.abs(a[m] - b[m]) / np.maximum(np.abs(b[m]), 1e-300)))


def max_abs_err(a, b):
    a = np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    m = np.isfinite(a) & np.isfinite(b)
    if not m.any():
        return 0.0
    return float(np.max(np.abs(a[m] - b[m])))


# ---------------------------------------------------------------------------
# product-side builders (isochrones_b200 host objects from the golden specs)
# ---------------------------------------------------------------------------

def product_ic(kind, model, bc, eep_bounds=(0, 60), ctx=None):
    import isochrones_b200 as ib

    m = dict(model)
    m.setdefault("limits", {"age": (5, 10.13), "feh": (-4, 0.5), "eep": tuple(eep_bounds), "mass": (0.1, 300)})
    return ib.ichrone_from_arrays(kind, m, bc, eep_bounds=eep_bounds, ctx=ctx)


def product_model_from_spec(spec, ic):
    """Build the product's BasicStarModel exactly the way oracle/make_golden.py built the reference's."""
    import isochrones_b200 as ib
    from isochrones_b200 import priors as P

    kwargs = {k: (v[0], v[1]) for k, v in spec["kwargs"].items()}
    init_kw = {k: (tuple(v) if isinstance(v, list) else v) for k, v in spec["init_kwargs"].items()}
    mod = ib.BasicStarModel(ic, N=spec["N"], **init_kw, **kwargs)
    tweaks = spec["tweaks"]
    if "set_bounds" in tweaks:
        mod.set_bounds(**{k: tuple(v) for k, v in tweaks["set_bounds"].items()})
    if "gauss_age" in tweaks:
        mean, sig, b = tweaks["gauss_age"]
        mod.set_prior(age=P.GaussianPrior(mean, sig, bounds=tuple(b)))
    return mod


def product_prior_cases():
    """The prior objects of oracle/make_golden.py::golden_priors, built from the product's classes."""
    from isochrones_b200 import priors as P

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
    return cases


# ---------------------------------------------------------------------------
# host replay of the on-device sampler's random stream (isochrones_b200/csrc/iso_sampler.cu): the same
# Philox4x32-10 counters, so a test can re-run the stretch move on the CPU with the oracle's lnpost
# ---------------------------------------------------------------------------
def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10 (Salmon et al. 2011); all arguments are uint32 arrays / scalars."""
    c = [np.asarray(x, dtype=np.uint64) for x in np.broadcast_arrays(c0, c1, c2, c3)]
    k0, k1 = np.uint64(k0), np.uint64(k1)
    M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = M0 * c[0]
        p1 = M1 * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & mask
        hi1, lo1 = p1 >> np.uint64(32), p1 & mask
        c = [(hi1 ^ c[1] ^ k0) & mask, lo1, (hi0 ^ c[3] ^ k1) & mask, lo0]
        k0 = (k0 + np.uint64(0x9E3779B9)) & mask
        k1 = (k1 + np.uint64(0xBB67AE85)) & mask
    return c


def u01(hi, lo):
    return (((hi << np.uint64(32)) | lo) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def stream(seed, gstep, half, walkers, chain):
    """``(u_z, j_offset_raw, u_accept)`` for the given walkers of one half-step — the device kernel's draws."""
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    ctr = gstep * 2 + half
    c0 = np.uint32(ctr & 0xFFFFFFFF)
    c1 = np.uint32(((ctr >> 32) & 0xFFFFFFFF) ^ ((chain << 8) & 0xFFFFFFFF))
    w = np.asarray(walkers, dtype=np.uint32)
    r = philox4x32_10(c0, c1, w, np.uint32(0), k0, k1)
    r2 = philox4x32_10(c0, c1, w, np.uint32(1), k0, k1)
    return u01(r[0], r[1]), r[2], u01(r2[0], r2[1])


def replay_stretch_move(lnpost_fn, p0, n_steps, seed, a=2.0, chain=0):
    """emcee stretch move driven by the kernel's random stream; ``lnpost_fn(rows[n, ndim]) -> [n]``.
    Returns (chain[n_steps, n_walkers, ndim], lnprob[n_steps, n_walkers], n_accepted)."""
    pos = np.array(p0, dtype=np.float64)
    n_walkers, ndim = pos.shape
    nhalf = n_walkers // 2
    lp = lnpost_fn(pos)
    out = np.empty((n_steps, n_walkers, ndim))
    out_lp = np.empty((n_steps, n_walkers))
    n_acc = 0
    for s in range(n_steps):
        for half in range(2):
            ks = np.arange(half * nhalf, (half + 1) * nhalf)
            other0 = (1 - half) * nhalf
            u, jraw, uacc = stream(seed, s, half, ks, chain)
            js = other0 + (jraw % np.uint64(nhalf)).astype(int)
            zr = (a - 1.0) * u + 1.0
            z = zr * zr / a
            c = pos[js]
            q = c - (c - pos[ks]) * z[:, None]
            lq = lnpost_fn(q)
            with np.errstate(divide="ignore", invalid="ignore"):
                lnpdiff = (ndim - 1) * np.log(z) + lq - lp[ks]
                acc = lnpdiff > np.log(uacc)
            pos[ks[acc]] = q[acc]
            lp[ks[acc]] = lq[acc]
            n_acc += int(acc.sum())
        out[s] = pos
        out_lp[s] = lp
    return out, out_lp, n_acc


# ---------------------------------------------------------------------------
# full-size benchmark grids + the reference's outputs on them (tests/golden/golden_full.npz)
# ---------------------------------------------------------------------------

def full_size_world():
    """(meta, arrays, grids): grids regenerated by isochrones_b200.synthetic and verified bit for bit against the
    SHA-256 digests recorded when oracle/make_golden_full.py ran the reference on them."""
    import hashlib
    import os

    from isochrones_b200 import synthetic as syn

    here = os.path.dirname(os.path.abspath(__file__))
    g = np.load(os.path.join(here, "golden", "golden_full.npz"), allow_pickle=False)
    meta = json.loads(str(g["meta_json"]))
    trk = syn.make_track_grid(columns=tuple(meta["track_columns"]))
    iso = syn.make_iso_grid(columns=tuple(meta["iso_columns"]))
    bc = syn.make_bc_grid(bands=tuple(meta["bands"]))
    for name, grid in (("track", trk), ("iso", iso), ("bc", bc)):
        got = hashlib.sha256(np.ascontiguousarray(grid["grid"]).tobytes()).hexdigest()
        assert got == meta[name + "_sha256"], "synthetic %s grid differs from the one the reference was run on" % name
    return meta, g, {"track": trk, "iso": iso, "bc": bc}


# ---------------------------------------------------------------------------
# a fabricated $ISOCHRONES tree in the reference's on-disk layout (tests of isochrones_b200.mistio / get_ichrone)
# ---------------------------------------------------------------------------

def write_bc_text_tables(datadir, bc, rvs=(2.0, 3.1)):
    """``bc`` (dense dict) -> MIST-layout text tables, one file per [Fe/H] and photometric system (bc.py:66-83: five
    header lines, column names on the sixth, rows ``Teff logg [Fe/H] Av Rv <bands>``); values printed with 17
    significant digits so that the parse is bit-exact.  Rows with ``Rv != 3.1`` carry shifted values."""
    import os

    from isochrones_b200 import bcio

    os.makedirs(datadir, exist_ok=True)
    by_phot = {}
    for j, b in enumerate(bc["columns"]):
        phot, col = bcio.mist_band(b)
        by_phot.setdefault(phot, []).append((j, col))
    teffs, loggs, fehs, avs = bc["axes"]
    for phot, cols in by_phot.items():
        for i_f, feh in enumerate(fehs):
            name = "feh%s%03.0f_%02d.%s" % ("m" if feh < 0 else "p", abs(feh) * 100, i_f, phot)
            with open(os.path.join(datadir, name), "w") as f:
                f.write("# MIST version number  = 1.2\n# MESA revision number =     7503\n# photometric system    = %s\n"
                        "# ABUNDANCES: [Fe/H] = %.2f\n# number of filters = %d\n" % (phot, feh, len(cols)))
                f.write("#  " + "  ".join(["Teff", "logg", "[Fe/H]", "Av", "Rv"] + [c for _, c in cols]) + "\n")
                for i_t, t in enumerate(teffs):
                    for i_g, g in enumerate(loggs):
                        for i_a, av in enumerate(avs):
                            for rv in rvs:
                                vals = [bc["grid"][i_t, i_g, i_f, i_a, j] + (0.0 if rv == 3.1 else 0.5 * rv) for j, _ in cols]
                                f.write(" ".join("%.17g" % v for v in [t, g, feh, av, rv] + vals) + "\n")


def write_isochrones_tree(root, kind, model, bc, sidecar=True):
    """Model grid as the reference's ``full_grid*.npz`` cache (+ axes sidecar) and the BC grid as MIST text tables
    under ``root`` (the layout of ``$ISOCHRONES``)."""
    import os

    from isochrones_b200 import mistio

    mistio.save_model_grid(model, kind, root=root, sidecar=sidecar)
    write_bc_text_tables(os.path.join(root, "BC", "mist"), bc)
