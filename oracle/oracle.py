"""TEST INFRASTRUCTURE — ctypes front-end of the C oracle (``oracle/iso_oracle.c``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this.
The converters are duck-typed on the attribute names the reference's classes use
(``DFInterpolator.grid/index_columns``, ``Prior.bounds/_norm``, ``BasicStarModel.kwargs/_priors``...)
so the same code wraps the reference's own objects (golden-vector generation in the build
container) and the product's host-side mirrors (GPU parity tests).
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD_DIR = os.path.join(HERE, "_build")
LIB_PATH = os.path.join(BUILD_DIR, "libiso_oracle.so")

ORC_MAX_BANDS = 32
ORC_MAX_COMP = 3

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)


def build(force=False):
    src = os.path.join(HERE, "iso_oracle.c")
    hdr = os.path.join(HERE, "iso_oracle.h")
    if not force and os.path.exists(LIB_PATH):
        if os.path.getmtime(LIB_PATH) >= max(os.path.getmtime(src), os.path.getmtime(hdr)):
            return LIB_PATH
    os.makedirs(BUILD_DIR, exist_ok=True)
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC", "-std=c99", "-D_GNU_SOURCE",
           src, "-o", LIB_PATH, "-lm"]
    subprocess.check_call(cmd)
    return LIB_PATH


class OrcGrid(C.Structure):
    _fields_ = [
        ("ndim", C.c_int32), ("ncols", C.c_int32), ("n", C.c_int64 * 4),
        ("grid", c_double_p), ("axes", c_double_p * 4),
    ]


class OrcPrior(C.Structure):
    pass


OrcPrior._fields_ = [
    ("kind", C.c_int32), ("bounded", C.c_int32), ("has_bounds", C.c_int32), ("local", C.c_int32),
    ("lo", C.c_double), ("hi", C.c_double), ("norm", C.c_double), ("a", C.c_double * 4),
    ("n_comp", C.c_int32), ("pad_", C.c_int32),
    ("breakpoints", C.c_double * (ORC_MAX_COMP - 1)), ("norms", C.c_double * ORC_MAX_COMP),
    ("lognorms", C.c_double * ORC_MAX_COMP),
    ("comp", C.POINTER(OrcPrior) * ORC_MAX_COMP), ("orig", C.POINTER(OrcPrior)),
]


class OrcModel(C.Structure):
    _fields_ = [
        ("n_stars", C.c_int32), ("eep_replaces_age", C.c_int32), ("index_order", C.c_int32 * 5),
        ("i_Teff", C.c_int32), ("i_logg", C.c_int32), ("i_feh", C.c_int32), ("i_Mbol", C.c_int32),
        ("i_orig", C.c_int32), ("i_deriv", C.c_int32), ("i_nu_max", C.c_int32), ("i_delta_nu", C.c_int32),
        ("n_bands", C.c_int32), ("has_plax", C.c_int32), ("has_nu_max", C.c_int32), ("has_delta_nu", C.c_int32),
        ("i_mags", C.c_int32 * ORC_MAX_BANDS),
        ("spec_vals", C.c_double * 3), ("spec_uncs", C.c_double * 3),
        ("mag_vals", C.c_double * ORC_MAX_BANDS), ("mag_uncs", C.c_double * ORC_MAX_BANDS),
        ("plax", C.c_double), ("plax_unc", C.c_double),
        ("nu_max", C.c_double), ("nu_max_unc", C.c_double), ("delta_nu", C.c_double), ("delta_nu_unc", C.c_double),
        ("model", C.POINTER(OrcGrid)), ("bc", C.POINTER(OrcGrid)),
        ("prior_eep", C.POINTER(OrcPrior)), ("prior_mass", C.POINTER(OrcPrior)), ("prior_age", C.POINTER(OrcPrior)),
        ("prior_feh", C.POINTER(OrcPrior)), ("prior_distance", C.POINTER(OrcPrior)), ("prior_AV", C.POINTER(OrcPrior)),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.orc_searchsorted.restype = C.c_int64
        L.orc_searchsorted.argtypes = [c_double_p, C.c_int64, C.c_double, c_int32_p]
        L.orc_interp_value.restype = None
        L.orc_interp_value.argtypes = [C.POINTER(OrcGrid), c_double_p, c_int32_p, C.c_int32, c_double_p]
        L.orc_interp_values.restype = None
        L.orc_interp_values.argtypes = [C.POINTER(OrcGrid), C.POINTER(c_double_p), C.c_int64, c_int32_p,
                                        C.c_int32, c_double_p]
        L.orc_interp_mags.restype = None
        L.orc_interp_mags.argtypes = [c_double_p, C.c_int64, c_int32_p, C.POINTER(OrcGrid), C.c_int32, C.c_int32,
                                      C.c_int32, C.c_int32, C.POINTER(OrcGrid), c_int32_p, C.c_int32,
                                      c_double_p, c_double_p, c_double_p, c_double_p]
        L.orc_interp_eeps.restype = None
        L.orc_interp_eeps.argtypes = [c_double_p, c_double_p, c_double_p, C.c_int64, c_double_p, C.c_int64, c_double_p,
                                      C.c_int64, c_double_p, C.c_int64, C.POINTER(C.c_int64), c_double_p]
        for name in ("orc_gauss_lnprob",):
            getattr(L, name).restype = C.c_double
            getattr(L, name).argtypes = [C.c_double] * 3
        L.orc_fast_addmags.restype = C.c_double
        L.orc_fast_addmags.argtypes = [c_double_p, C.c_int32]
        L.orc_prior_call.restype = C.c_double
        L.orc_prior_call.argtypes = [C.POINTER(OrcPrior), C.c_double]
        L.orc_prior_lnpdf.restype = C.c_double
        L.orc_prior_lnpdf.argtypes = [C.POINTER(OrcPrior), C.c_double]
        L.orc_eep_prior_lnpdf.restype = C.c_double
        L.orc_eep_prior_lnpdf.argtypes = [C.POINTER(OrcModel), C.c_double, C.c_double, C.c_double]
        for name in ("orc_lnlike", "orc_lnprior", "orc_lnpost"):
            getattr(L, name).restype = C.c_double
            getattr(L, name).argtypes = [C.POINTER(OrcModel), c_double_p]
        L.orc_lnpost_batch.restype = None
        L.orc_lnpost_batch.argtypes = [C.POINTER(OrcModel), c_double_p, C.c_int64, c_double_p, c_double_p,
                                       c_double_p, C.c_int32]
        L.orc_lnpost_catalog.restype = None
        L.orc_lnpost_catalog.argtypes = [C.POINTER(C.POINTER(OrcModel)), c_int32_p, c_double_p, C.c_int64,
                                         c_double_p, C.c_int32]
        L.orc_stretch_move.restype = None
        L.orc_stretch_move.argtypes = [C.POINTER(C.POINTER(OrcModel)), C.c_int32, C.c_int32, C.c_int32, c_double_p,
                                       c_double_p, C.c_int64, C.c_int32, C.c_uint64, C.c_double, c_double_p, c_double_p,
                                       C.POINTER(C.c_int64), C.c_int32]
        L.orc_mnest_prior.restype = None
        L.orc_mnest_prior.argtypes = [c_double_p, c_double_p, C.c_int32, c_double_p]
        L.orc_max_threads.restype = C.c_int32
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(c_double_p)


def _ip(a):
    return a.ctypes.data_as(c_int32_p)


class Grid:
    """Oracle view of a dense grid; accepts (grid, axes) or anything with .grid / .index_columns."""

    def __init__(self, grid, axes=None):
        if axes is None:
            grid, axes = grid.grid, grid.index_columns
        self.grid = np.ascontiguousarray(grid, dtype=np.float64)
        self.axes = [np.ascontiguousarray(a, dtype=np.float64) for a in axes]
        self.ndim = len(self.axes)
        assert self.grid.ndim == self.ndim + 1 and 2 <= self.ndim <= 4
        s = OrcGrid()
        s.ndim = self.ndim
        s.ncols = self.grid.shape[-1]
        for d in range(self.ndim):
            assert self.grid.shape[d] == len(self.axes[d])
            s.n[d] = len(self.axes[d])
            s.axes[d] = _dp(self.axes[d])
        s.grid = _dp(self.grid)
        self.struct = s

    def interp_value(self, x, icols):
        x = np.ascontiguousarray(x, dtype=np.float64)
        icols = np.ascontiguousarray(icols, dtype=np.int32)
        out = np.empty(len(icols))
        lib().orc_interp_value(C.byref(self.struct), _dp(x), _ip(icols), len(icols), _dp(out))
        return out

    def interp_values(self, xx, icols):
        xx = [np.ascontiguousarray(a, dtype=np.float64) for a in xx]
        n = len(xx[0])
        icols = np.ascontiguousarray(icols, dtype=np.int32)
        out = np.empty((n, len(icols)))
        ptrs = (c_double_p * self.ndim)(*[_dp(a) for a in xx])
        lib().orc_interp_values(C.byref(self.struct), ptrs, n, _ip(icols), len(icols), _dp(out))
        return out


def interp_mags(pars, index_order, model, i_Teff, i_logg, i_feh, i_Mbol, bc, bc_cols):
    """mags.py:64-124 — ``pars`` is ``[5, N]``."""
    pars = np.ascontiguousarray(pars, dtype=np.float64)
    n = pars.shape[1]
    io = np.ascontiguousarray(index_order, dtype=np.int32)
    bc_cols = np.ascontiguousarray(bc_cols, dtype=np.int32)
    teff, logg, feh = np.empty(n), np.empty(n), np.empty(n)
    mags = np.empty((n, len(bc_cols)))
    lib().orc_interp_mags(_dp(pars), n, _ip(io), C.byref(model.struct), i_Teff, i_logg, i_feh, i_Mbol,
                          C.byref(bc.struct), _ip(bc_cols), len(bc_cols), _dp(teff), _dp(logg), _dp(feh), _dp(mags))
    return teff, logg, feh, mags


def interp_eeps(xs, x0s, x1s, ii0, ii1, arrays, lengths):
    """interp.py:488-558 — (age, feh, mass) -> EEP; ``arrays`` is ``[len(ii0) * len(ii1), n_eep]``."""
    xs, x0s, x1s = [np.ascontiguousarray(a, dtype=np.float64) for a in (xs, x0s, x1s)]
    ii0, ii1 = np.ascontiguousarray(ii0, dtype=np.float64), np.ascontiguousarray(ii1, dtype=np.float64)
    arrays = np.ascontiguousarray(arrays, dtype=np.float64)
    lengths = np.ascontiguousarray(lengths, dtype=np.int64)
    assert arrays.shape[0] == len(ii0) * len(ii1) == len(lengths)
    out = np.empty(len(xs))
    lib().orc_interp_eeps(_dp(xs), _dp(x0s), _dp(x1s), len(xs), _dp(ii0), len(ii0), _dp(ii1), len(ii1), _dp(arrays),
                          arrays.shape[1], lengths.ctypes.data_as(C.POINTER(C.c_int64)), _dp(out))
    return out


_KINDS = {
    "FlatPrior": 1, "AVPrior": 1,
    "FlatLogPrior": 2, "AgePrior": 2,
    "PowerLawPrior": 3, "DistancePrior": 3, "QPrior": 3, "SalpeterPrior": 3,
    "GaussianPrior": 4, "LogNormalPrior": 5, "FehPrior": 6,
    "BrokenPrior": 7, "ChabrierPrior": 7, "EEP_prior": 8,
}
_BOUNDED = {1, 2, 3, 4, 8}


class Prior:
    """Oracle image of a prior object (reference class or product mirror — same attribute names)."""

    def __init__(self, obj):
        names = [c.__name__ for c in type(obj).__mro__] + list(getattr(obj, "_mro_names", []))
        kind = None
        for name in names:
            if name in _KINDS:
                kind = _KINDS[name]
                break
        if kind is None:
            raise ValueError("oracle cannot represent prior %r" % (obj,))
        s = OrcPrior()
        s.kind = kind
        s.bounded = 1 if kind in _BOUNDED else 0
        raw_bounds = getattr(obj, "_bounds", None)
        s.has_bounds = 0 if raw_bounds is None else 1
        if raw_bounds is not None:
            s.lo, s.hi = float(raw_bounds[0]), float(raw_bounds[1])
        else:
            s.lo, s.hi = -np.inf, np.inf
        s.norm = float(getattr(obj, "_norm", 1.0))
        self.children = []
        if kind == 3:
            s.a[0] = float(obj.alpha)
        elif kind == 4:
            s.a[0], s.a[1], s.a[2], s.a[3] = float(obj.mean), float(obj.sigma), float(obj.norm), float(obj.lognorm)
        elif kind == 5:
            s.a[0], s.a[1], s.a[2], s.a[3] = float(obj.mu), float(obj.sigma), float(obj.scale), float(obj.log_s)
        elif kind == 6:
            s.a[0] = float(obj.halo_fraction)
            s.local = 1 if obj.local else 0
        elif kind == 7:
            s.n_comp = obj.n_components
            assert s.n_comp <= ORC_MAX_COMP
            for i, b in enumerate(obj.breakpoints):
                s.breakpoints[i] = float(b)
            for i in range(s.n_comp):
                s.norms[i] = float(obj.norms[i])
                s.lognorms[i] = float(obj.lognorms[i])
                child = Prior(obj.components[i])
                self.children.append(child)
                s.comp[i] = C.pointer(child.struct)
        elif kind == 8:
            child = Prior(obj.orig_prior)
            self.children.append(child)
            s.orig = C.pointer(child.struct)
        self.struct = s

    def __call__(self, x):
        return lib().orc_prior_call(C.byref(self.struct), float(x))

    def lnpdf(self, x):
        return lib().orc_prior_lnpdf(C.byref(self.struct), float(x))


class StarModel:
    """Oracle image of a ``BasicStarModel``-like object (reference or product mirror).

    Reads: ``ic.param_index_order``, ``ic.eep_replaces``, ``ic.model_grid.interp``, ``ic.bc_grid.interp``,
    ``N``, ``kwargs``, ``bands``, ``spec_props``, ``_priors`` (starmodel.py:1361-1635).
    """

    def __init__(self, mod, model_grid=None, bc_grid=None):
        ic = mod.ic
        mi = ic.model_grid.interp
        bi = ic.bc_grid.interp
        self.model = model_grid if model_grid is not None else Grid(mi)
        self.bc = bc_grid if bc_grid is not None else Grid(bi)
        s = OrcModel()
        s.n_stars = mod.N
        s.eep_replaces_age = 1 if ic.eep_replaces == "age" else 0
        for i, v in enumerate(ic.param_index_order):
            s.index_order[i] = int(v)
        ci = mi.column_index
        s.i_Teff, s.i_logg, s.i_feh, s.i_Mbol = ci["Teff"], ci["logg"], ci["feh"], ci["Mbol"]
        if s.eep_replaces_age:
            s.i_orig, s.i_deriv = ci["age"], ci["dt_deep"]
        else:
            s.i_orig, s.i_deriv = ci["mass"], ci["dm_deep"]
        s.i_nu_max = ci.get("nu_max", -1)
        s.i_delta_nu = ci.get("delta_nu", -1)
        bands = list(mod.bands)
        s.n_bands = len(bands)
        assert s.n_bands <= ORC_MAX_BANDS
        for i, b in enumerate(bands):
            s.i_mags[i] = bi.column_index[b]
            s.mag_vals[i], s.mag_uncs[i] = [float(v) for v in mod.kwargs[b]]
        for i, (val, unc) in enumerate(mod.spec_props):
            s.spec_vals[i], s.spec_uncs[i] = float(val), float(unc)
        s.has_plax = 1 if "parallax" in mod.kwargs else 0
        if s.has_plax:
            s.plax, s.plax_unc = [float(v) for v in mod.kwargs["parallax"]]
        s.has_nu_max = 1 if "nu_max" in mod.kwargs else 0
        if s.has_nu_max:
            s.nu_max, s.nu_max_unc = [float(v) for v in mod.kwargs["nu_max"]]
        s.has_delta_nu = 1 if "delta_nu" in mod.kwargs else 0
        if s.has_delta_nu:
            s.delta_nu, s.delta_nu_unc = [float(v) for v in mod.kwargs["delta_nu"]]
        s.model = C.pointer(self.model.struct)
        s.bc = C.pointer(self.bc.struct)
        self.priors = {k: Prior(mod._priors[k]) for k in ("eep", "mass", "age", "feh", "distance", "AV")}
        for k, p in self.priors.items():
            setattr(s, "prior_" + k, C.pointer(p.struct))
        self.struct = s
        self.ndim = 4 + mod.N

    def _p(self, pars):
        p = np.ascontiguousarray(pars, dtype=np.float64)
        assert p.shape == (self.ndim,)
        return p

    def lnlike(self, pars):
        return lib().orc_lnlike(C.byref(self.struct), _dp(self._p(pars)))

    def lnprior(self, pars):
        return lib().orc_lnprior(C.byref(self.struct), _dp(self._p(pars)))

    def lnpost(self, pars):
        return lib().orc_lnpost(C.byref(self.struct), _dp(self._p(pars)))

    def lnpost_batch(self, pars, n_threads=1, parts=False):
        pars = np.ascontiguousarray(pars, dtype=np.float64)
        assert pars.ndim == 2 and pars.shape[1] == self.ndim
        n = pars.shape[0]
        lnpost = np.empty(n)
        if parts:
            lnprior, lnlike = np.empty(n), np.empty(n)
            lib().orc_lnpost_batch(C.byref(self.struct), _dp(pars), n, _dp(lnpost), _dp(lnprior), _dp(lnlike),
                                   n_threads)
            return lnpost, lnprior, lnlike
        lib().orc_lnpost_batch(C.byref(self.struct), _dp(pars), n, _dp(lnpost), None, None, n_threads)
        return lnpost


def lnpost_catalog(models, model_of_row, pars, n_threads=1):
    pars = np.ascontiguousarray(pars, dtype=np.float64)
    mor = np.ascontiguousarray(model_of_row, dtype=np.int32)
    n = pars.shape[0]
    arr = (C.POINTER(OrcModel) * len(models))(*[C.pointer(m.struct) for m in models])
    out = np.empty(n)
    lib().orc_lnpost_catalog(arr, _ip(mor), _dp(pars), n, _dp(out), n_threads)
    return out


def stretch_move(models, p0, n_steps, seed, a=2.0, step0=0, n_threads=1, store=True):
    """emcee's stretch move on the CPU with the product sampler's random stream (``orc_stretch_move``): ``models`` is one
    ``StarModel`` (every chain samples it) or one per chain; ``p0`` is ``[n_walkers, ndim]`` or ``[n_chains, n_walkers,
    ndim]``.  Returns ``(chain[n_steps, n_chains, n_walkers, ndim] or None, lnprob[...] or None, pos, lnprob_now,
    n_accepted[n_chains])``."""
    models = list(models) if isinstance(models, (list, tuple)) else [models]
    pos = np.array(p0, dtype=np.float64)
    if pos.ndim == 2:
        pos = pos[None]
    pos = np.ascontiguousarray(pos)
    n_chains, n_walkers, ndim = pos.shape
    assert ndim == models[0].ndim and len(models) in (1, n_chains)
    lp = np.empty((n_chains, n_walkers))
    for c in range(n_chains):
        lp[c] = models[c % len(models)].lnpost_batch(pos[c], n_threads=n_threads)
    arr = (C.POINTER(OrcModel) * len(models))(*[C.pointer(m.struct) for m in models])
    chain = np.empty((n_steps, n_chains, n_walkers, ndim)) if store else None
    lnp = np.empty((n_steps, n_chains, n_walkers)) if store else None
    acc = np.zeros(n_chains, dtype=np.int64)
    lib().orc_stretch_move(arr, len(models), n_chains, n_walkers, _dp(pos), _dp(lp), int(step0), int(n_steps),
                           C.c_uint64(int(seed)), float(a), _dp(chain) if store else None, _dp(lnp) if store else None,
                           acc.ctypes.data_as(C.POINTER(C.c_int64)), int(n_threads))
    return chain, lnp, pos, lp, acc


def max_threads():
    return int(lib().orc_max_threads())
