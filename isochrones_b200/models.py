"""``ModelGridInterpolator`` / ``EvolutionTrackInterpolator`` / ``IsochroneInterpolator`` — host-side mirror of
the part of the reference's ``isochrones/models.py:253-718`` that sits on the lnpost path.

Kept from the reference: ``param_names``, ``eep_replaces``, ``param_index_order`` (models.py:259, 665-669,
692-696), ``bands``, ``model_grid`` / ``bc_grid`` objects exposing ``.interp`` (a ``DFInterpolator``),
``.get_limits(prop)`` and ``.bands``, ``interp_value(pars, props)`` (models.py:390-400), ``interp_mag(pars,
bands)`` (:402-445), ``initialize`` (:349-358) and the property shortcuts (``mass``, ``radius``, ``Teff`` ...).

Out of scope: grid download / parsing (``Grid``, ``StellarModelGrid``, ``MIST*Grid``).  The grids
come in as dense arrays — the reference's own cached ``full_grid*.npz`` dense grids + the MIST BC text tables under
``$ISOCHRONES`` (``get_ichrone`` through :mod:`isochrones_b200.mistio`), or, on explicit request,
synthetic MIST-shaped ones from :mod:`isochrones_b200.synthetic` — and are staged once to HBM.  Also here (SURVEY.md
§8f-3): ``get_eep`` (``interp_eep(s)``, interp.py:488-558, and a batched root bracketing in place of the reference's
per-star scipy minimisation) and ``generate`` on the same kernels.
"""
import ctypes as C

import numpy as np

from . import _lib
from .interp import DFInterpolator

MODEL_PACK_COLUMNS = ("Teff", "logg", "feh", "Mbol", None, None, "nu_max", "delta_nu")   # ISO_MP_* order


class ModelGrid(object):
    """Minimal stand-in for the reference's ``StellarModelGrid`` objects: a dense interpolator + limits."""

    def __init__(self, interp, limits, eep_replaces, name="grid"):
        self.interp = interp
        self._limits = dict(limits)
        self.eep_replaces = eep_replaces
        self.name = name

    def get_limits(self, prop):
        """Declared bounds (mist/models.py:37), else the range of the column over the grid (grid.py:58-61)."""
        if prop not in self._limits:
            col = self.interp.grid[..., self.interp.column_index[prop]]
            self._limits[prop] = (float(np.nanmin(col)), float(np.nanmax(col)))
        return self._limits[prop]

    def get_array_grids(self):
        """``(age_grid[n_feh * n_mass, n_eep], dt_deep_grid, lengths)`` of an evolution-track grid — the irregular
        per-track age arrays ``interp_eep`` searches (host-side data prep, models.py:171-205): the populated leading
        run of every track's age / dt_deep column, NaN beyond it."""
        if self.eep_replaces != "age":
            raise NotImplementedError("Not implemented for isochrone grids yet!")
        if getattr(self, "_array_grids", None) is None:
            it = self.interp
            g = it.grid
            n_eep = g.shape[2]
            age = g[..., it.column_index["age"]].reshape(-1, n_eep)
            dt = g[..., it.column_index["dt_deep"]].reshape(-1, n_eep)
            nan = np.isnan(age)
            lengths = np.where(nan.any(axis=1), nan.argmax(axis=1), n_eep).astype(np.int64)
            mask = np.arange(n_eep)[None, :] < lengths[:, None]
            self._array_grids = (np.where(mask, age, np.nan), np.where(mask, dt, np.nan), lengths)
        return self._array_grids

    age_grid = property(lambda self: self.get_array_grids()[0])
    dt_deep_grid = property(lambda self: self.get_array_grids()[1])
    array_lengths = property(lambda self: self.get_array_grids()[2])

    @property
    def n_masses(self):
        return len(self.masses)

    @property
    def fehs(self):
        i = 0 if self.eep_replaces == "age" else 1
        return self.interp.index_columns[i]

    @property
    def masses(self):
        return self.interp.index_columns[1]

    @property
    def ages(self):
        return self.interp.index_columns[0]


class BCGrid(object):
    """Stand-in for ``BolometricCorrectionGrid`` (bc.py): 4-D interpolator over (Teff, logg, [Fe/H], Av)."""

    def __init__(self, interp, bands=None):
        self.interp = interp
        self.bands = list(bands) if bands is not None else list(interp.columns)


class ModelGridInterpolator(object):
    # transformation from desired param order to that expected by interp functions (models.py:259)
    _param_index_order = (1, 2, 0, 3, 4)
    eep_bounds = (0, 1710)        # mist/isochrone.py:9, 21
    eep_replaces = None
    param_names = None

    def __init__(self, model_grid, bc_grid, bands=None, eep_bounds=None, ctx=None, **kwargs):
        self.bands = list(bands) if bands is not None else list(bc_grid.bands)
        self._model_grid = model_grid
        self._bc_grid = bc_grid
        self.param_index_order = list(self._param_index_order)
        self.kwargs = kwargs
        if eep_bounds is not None:
            self.eep_bounds = tuple(eep_bounds)
        self._ctx = ctx
        self._model_pack = None
        self._bc_packs = {}

    # ---- reference attribute surface -----------------------------------------------------------------------
    @property
    def model_grid(self):
        return self._model_grid

    @property
    def bc_grid(self):
        return self._bc_grid

    @property
    def name(self):
        return self.model_grid.name

    @property
    def ctx(self):
        if self._ctx is None:
            self._ctx = _lib.default_context()
        return self._ctx

    def _limit(self, prop, i):
        return self.model_grid.get_limits(prop)[i]

    minfeh = property(lambda self: self._limit("feh", 0))
    maxfeh = property(lambda self: self._limit("feh", 1))
    mineep = property(lambda self: self._limit("eep", 0))
    maxeep = property(lambda self: self._limit("eep", 1))
    minage = property(lambda self: self._limit("age", 0))
    maxage = property(lambda self: self._limit("age", 1))
    minmass = property(lambda self: self._limit("mass", 0))
    maxmass = property(lambda self: self._limit("mass", 1))

    @property
    def fehs(self):
        return self.model_grid.fehs

    def initialize(self, pars=None):
        """Touch both grids once (stages them in HBM) and check that a typical star is inside them."""
        if pars is None:
            pars = {"age": [1.04, 320.0, -0.35, 10000.0, 0.34], "mass": [320, 9.7, -0.35, 10000.0, 0.34]}[self.eep_replaces]
        Teff, logg, feh, mags = self.interp_mag(pars, self.bands)
        if not (np.all(np.isfinite([Teff, logg, feh])) and np.all(np.isfinite(mags))):
            raise AssertionError("default star %r falls outside the model or BC grid" % (pars,))

    # ---- device-resident packs the fused kernels gather from ---------------------------------------------------
    @property
    def model_pack(self):
        """8-column pack (Teff, logg, feh, Mbol, orig, deriv, nu_max, delta_nu), one 64-byte node per grid point."""
        if self._model_pack is None:
            ci = self.model_grid.interp.column_index
            orig, deriv = ("age", "dt_deep") if self.eep_replaces == "age" else ("mass", "dm_deep")
            names = ["Teff", "logg", "feh", "Mbol", orig, deriv, "nu_max", "delta_nu"]
            for n in names[:6]:
                if n not in ci:
                    raise KeyError("model grid lacks column %r needed by the lnpost path" % n)
            cols = [ci.get(n, -1) for n in names]
            interp = self.model_grid.interp
            if interp._device_grid is not None:
                self._model_pack = interp.device_grid.repack(cols, _lib.ISO_MP_NCOLS)
            else:
                # build the pack on the host so the full (18-column) grid never has to occupy HBM
                g = interp.grid
                pack = np.zeros(g.shape[:-1] + (_lib.ISO_MP_NCOLS,))
                for j, c in enumerate(cols):
                    if c >= 0:
                        pack[..., j] = g[..., c]
                self._model_pack = _lib.DeviceGrid(self.ctx, pack, interp.index_columns)
        return self._model_pack

    def bc_pack(self, bands):
        """(device grid, column of each band) holding ``bands`` padded to a multiple of 4 columns (32-byte sectors)."""
        key = tuple(bands)
        if key not in self._bc_packs:
            interp = self.bc_grid.interp
            cols = [interp.column_index[b] for b in key] or [0]
            if len(cols) > _lib.ISO_MAX_BANDS:
                raise ValueError("at most %d bands per star model" % _lib.ISO_MAX_BANDS)
            ncols_out = 4 * ((len(cols) + 3) // 4)
            g = interp.grid
            pack = np.zeros(g.shape[:-1] + (ncols_out,))
            for j, c in enumerate(cols):
                pack[..., j] = g[..., c]
            self._bc_packs[key] = _lib.DeviceGrid(self.ctx, pack, interp.index_columns)
        return self._bc_packs[key]

    # ---- the hot entry points -----------------------------------------------------------------------------
    def interp_value(self, pars, props, out=None):
        """pars : (mass, eep, feh[, distance, AV]) for tracks / (eep, age, feh[, ..]) for isochrones
        (models.py:390-400); scalars give ``[len(props)]``, arrays ``[N, len(props)]`` (``out``: optional array to
        fill, see ``DFInterpolator.__call__``)."""
        i0, i1, i2 = self.param_index_order[:3]
        try:
            pars = np.atleast_1d(pars[self.param_index_order])
            p = [pars[0], pars[1], pars[2]]
            if pars.ndim == 1:
                p = [float(v) for v in p]
        except (TypeError, IndexError):
            p = [pars[i0], pars[i1], pars[i2]]
        if isinstance(props, str):
            props = [props]
        return self.model_grid.interp(p, props, out=out)

    def interp_mag(self, pars, bands, out=None):
        """pars : five parameters in ``param_names`` order; returns ``(Teff, logg, feh, mags)`` — scalars and a
        ``[n_bands]`` array for a single point, ``[N]`` arrays and ``[N, n_bands]`` otherwise (models.py:402-445).

        ``out`` (optional, batch calls): ``(Teff[N], logg[N], feh[N], mags[N, n_bands])`` float64 arrays to fill —
        page-locked ones (``ctx.pinned_empty``) are written by DMA directly.  The reference stacks the broadcast
        parameters into ``pars[5, N]`` (``np.resize`` per parameter, models.py:416-424); here parameters that already
        are full-size float64 arrays go to the device as they are (``iso_interp_mags_cols``)."""
        bands = list(bands) if bands is not None else []
        scalar = False
        try:
            # a single point: five scalars (checked before any array conversion — the parameters may be large arrays)
            if not (isinstance(pars, np.ndarray) or all(np.ndim(x) == 0 or np.size(x) == 1 for x in pars)):
                raise ValueError
            q = np.atleast_1d(pars).astype(float).squeeze()
            if q.ndim > 1 or q.shape != (5,):
                raise ValueError
            scalar = True
            n = 1
            cols = [q[j:j + 1].copy() for j in range(5)]
        except (TypeError, ValueError):
            b = np.broadcast(*pars)
            n = int(b.size)
            cols = []
            for x in pars:
                if isinstance(x, np.ndarray) and x.dtype == np.float64 and x.size == n and x.flags["C_CONTIGUOUS"]:
                    cols.append(x.reshape(-1))
                elif np.ndim(x) == 0:
                    cols.append(np.full(n, float(x)))
                else:   # the reference's np.resize semantics (cyclic repeat of the flattened values, models.py:419)
                    cols.append(np.ascontiguousarray(np.resize(x, b.shape), dtype=np.float64).ravel())
        bc, = (self.bc_pack(bands),)
        if out is not None and not scalar:
            teff, logg, feh, mags = out
            for a, shp in ((teff, (n,)), (logg, (n,)), (feh, (n,)), (mags, (n, len(bands)))):
                if a.shape != shp or a.dtype != np.float64 or not a.flags["C_CONTIGUOUS"]:
                    raise ValueError("out arrays must be C-contiguous float64 of shapes [N], [N], [N], [N, n_bands]")
        else:
            teff, logg, feh = np.empty(n), np.empty(n), np.empty(n)
            mags = np.empty((n, len(bands)))
        io = np.array(self.param_index_order, dtype=np.int32)
        bc_cols = np.arange(len(bands), dtype=np.int32)
        ptrs = (_lib.c_double_p * 5)(*[_lib.dp(c) for c in cols])
        ctx = self.ctx
        ctx.check(_lib.lib().iso_interp_mags_cols(
            ctx.handle, self.model_pack.handle, bc.handle, _lib.ip(io), 0, 1, 2, 3, _lib.ip(bc_cols), len(bands),
            ptrs, n, _lib.dp(teff), _lib.dp(logg), _lib.dp(feh), _lib.dp(mags)))
        if scalar:
            return float(teff[0]), float(logg[0]), float(feh[0]), mags[0]
        return teff, logg, feh, mags

    def interp_mag_device(self, d_pars, n, bands, d_Teff, d_logg, d_feh, d_mags):
        """``interp_mag`` on device buffers (asynchronous): ``d_pars`` = five device pointers (``param_names`` order,
        each ``[n]``), outputs device ``[n]``, ``[n]``, ``[n]``, ``[n, len(bands)]`` (``iso_interp_mags_device``)."""
        bands = list(bands)
        bc = self.bc_pack(bands)
        io = np.array(self.param_index_order, dtype=np.int32)
        bc_cols = np.arange(len(bands), dtype=np.int32)
        ptrs = (C.c_void_p * 5)(*d_pars)
        self.ctx.check(_lib.lib().iso_interp_mags_device(
            self.ctx.handle, self.model_pack.handle, bc.handle, _lib.ip(io), 0, 1, 2, 3, _lib.ip(bc_cols), len(bands), ptrs,
            int(n), d_Teff, d_logg, d_feh, d_mags))

    def get_eep(self, mass, age, feh, accurate=False, **kwargs):
        """EEP of stars of given (mass, log10 age, feh).  Scalars give a float, anything else is broadcast.

        Evolution-track grids: the fast bracketing interpolation ``interp_eep(s)`` of the reference (interp.py:488-558,
        models.py:501-542) on the GPU (``iso_interp_eeps``).  ``accurate=True`` (the only form isochrone grids have)
        solves ``age(mass, eep, feh) = age`` resp. ``initial_mass(eep, age, feh) = mass`` for the EEP instead: where the
        reference runs one scipy Nelder-Mead minimisation per star on the host (models.py:544-578), this brackets the
        root for ALL stars at once by repeated 64-way subdivision of the EEP range — three interpolation launches of
        ``64 N`` points — and finishes with the secant of the last bracket (:meth:`solve_eep`).  Stars without a root
        give NaN (``return_nan=False``: raise, like the reference; the reference's optimiser keywords are accepted and
        have no meaning here)."""
        return_nan = kwargs.pop("return_nan", True)
        scalar = all(isinstance(v, (float, int)) for v in (mass, age, feh))
        b = np.broadcast(mass, age, feh)
        mass_a, age_a, feh_a = [np.ascontiguousarray(np.atleast_1d(np.resize(x, b.shape)).astype(float).ravel())
                                for x in (mass, age, feh)]
        if accurate or self.eep_replaces != "age":
            out = self.solve_eep(mass_a, age_a, feh_a)
            if not return_nan and np.isnan(out).any():
                raise RuntimeError("no EEP reproduces (mass, age, feh) = %r"
                                   % ((mass_a[np.isnan(out)][0], age_a[np.isnan(out)][0], feh_a[np.isnan(out)][0]),))
        else:
            lengths = np.ascontiguousarray(self.model_grid.array_lengths, dtype=np.int32)
            out = np.empty(len(age_a))
            ctx = self.ctx
            ctx.check(_lib.lib().iso_interp_eeps(ctx.handle, self.model_pack.handle, 4, _lib.ip(lengths), _lib.dp(age_a),
                                                 _lib.dp(feh_a), _lib.dp(mass_a), len(age_a), _lib.dp(out)))
        return float(out[0]) if scalar else out

    def solve_eep(self, mass, age, feh, fan=64, passes=3):
        """Batched root bracketing behind ``get_eep(accurate=True)``: ``[N]`` arrays in, ``[N]`` EEPs (NaN: no root).

        The target column (``age`` along a track, ``initial_mass`` along an isochrone) is non-decreasing in EEP over
        the populated part of the grid, so the root is the first sign change of ``column(eep) - target`` among the
        finite samples; every pass evaluates ``fan + 1`` equally spaced EEPs per star in ONE launch and keeps the
        bracketing interval (resolution after three passes: 1710 / 64^3 = 0.007 EEP)."""
        n = len(mass)
        track = self.eep_replaces == "age"
        target, column = (age, "age") if track else (mass, "initial_mass")
        lo = np.full(n, float(self.eep_bounds[0]) if self.eep_bounds[0] >= 1 else 1.0)
        hi = np.full(n, float(self.model_grid.interp.index_columns[2][-1]))
        t = np.linspace(0.0, 1.0, fan + 1)
        f_lo = f_hi = None
        alive = np.ones(n, dtype=bool)
        for _ in range(passes):
            eeps = lo[:, None] + (hi - lo)[:, None] * t[None, :]                       # [N, fan + 1]
            rep = [np.repeat(x, fan + 1) for x in (mass, age, feh)]
            pars = [rep[0], eeps.ravel(), rep[2]] if track else [eeps.ravel(), rep[1], rep[2]]
            f = self.interp_value(pars, [column])[:, 0].reshape(n, fan + 1) - target[:, None]
            finite = np.isfinite(f)
            # first j with f[j] <= 0 <= f[j + 1] (both finite)
            ok = finite[:, :-1] & finite[:, 1:] & (f[:, :-1] <= 0.0) & (f[:, 1:] >= 0.0)
            alive &= ok.any(axis=1)
            j = np.where(alive, ok.argmax(axis=1), 0)
            rows = np.arange(n)
            lo, hi = eeps[rows, j], eeps[rows, j + 1]
            f_lo, f_hi = f[rows, j], f[rows, j + 1]
        with np.errstate(invalid="ignore", divide="ignore"):
            frac = np.where(f_hi > f_lo, -f_lo / (f_hi - f_lo), 0.0)
        return np.where(alive, lo + frac * (hi - lo), np.nan)

    def generate(self, mass, age, feh, props="all", bands=None, eeps=None, return_df=True, return_dict=False,
                 distance=10, AV=0, all_As=False, **kwargs):
        """Forward-simulate stars of given (mass, log10 age, feh) on an evolution-track grid — the reference's
        ``generate`` (models.py:580-629; same signature and outputs) as three batched launches: ``get_eep``, one
        ``interp_value`` over the requested model-grid columns, one ``interp_mag`` at (distance, AV).

        Output: a DataFrame (default), a dict of columns (``return_dict``) or the bare ``[N, n_props + n_bands]``
        array (``return_df=False``); frames / dicts also carry ``distance``, ``AV``, ``initial_feh``,
        ``requested_age`` and, with ``all_As``, the per-band extinction ``A_<band>`` (magnitude minus its AV = 0
        value)."""
        mass, age, feh, distance, AV = [np.asarray(x, dtype=float) for x in (mass, age, feh, distance, AV)]
        shape = np.broadcast(mass, age, feh, distance, AV).shape
        flat = [np.ascontiguousarray(np.broadcast_to(x, shape)).reshape(-1) for x in (mass, age, feh, distance, AV)]
        m, a, f, d, av = flat
        bands = list(self.bands if bands is None else bands)
        e = self.get_eep(m, a, f, **kwargs) if eeps is None else np.broadcast_to(np.asarray(eeps, dtype=float), shape).reshape(-1)
        names = list(self.model_grid.interp.columns) if isinstance(props, str) and props == "all" else list(props)
        table = {}
        if names:
            table.update(zip(names, self.interp_value([m, e, f], names).T))
        if bands:
            mags = np.atleast_2d(self.interp_mag([m, e, f, d, av], bands)[3])     # a single star comes back 1-D
            table.update(("{}_mag".format(b), mags[:, j]) for j, b in enumerate(bands))
        if not (return_df or return_dict):
            out = np.column_stack(list(table.values())) if table else np.empty((len(m), 0))
            return out[0] if shape == () else out
        table.update(distance=d, AV=av, initial_feh=f, requested_age=a)
        if all_As and bands:
            clear = np.atleast_2d(self.interp_mag([m, e, f, d, np.zeros_like(av)], bands)[3])
            table.update(("A_{}".format(b), table["{}_mag".format(b)] - clear[:, j]) for j, b in enumerate(bands))
        if return_dict:
            return {k: (v[0] if shape == () else v) for k, v in table.items()}
        import pandas as pd

        return pd.DataFrame(table)

    def __call__(self, p1, p2, p3, distance=10.0, AV=0.0):
        """All model-grid columns + magnitudes as a DataFrame (models.py:471-482) — same kernels, wider output."""
        import pandas as pd

        p1, p2, p3, dist, AV = [np.atleast_1d(a).astype(float).ravel()
                                for a in np.broadcast_arrays(p1, p2, p3, distance, AV)]
        pars = [p1, p2, p3, dist, AV]
        prop_cols = list(self.model_grid.interp.columns)
        props = self.interp_value(pars, prop_cols)
        _, _, _, mags = self.interp_mag(pars, self.bands)
        cols = prop_cols + ["{}_mag".format(b) for b in self.bands]
        values = np.concatenate([np.atleast_2d(props), np.atleast_2d(mags)], axis=1)
        return pd.DataFrame(values, columns=cols)


def _column_shortcut(column):
    def shortcut(self, *pars):
        return self.interp_value(pars, [column]).squeeze()

    shortcut.__name__ = column
    shortcut.__doc__ = "``%s`` interpolated at ``pars`` (in ``param_names`` order); scalars or arrays." % column
    return shortcut


# ic.mass(*pars), ic.radius(*pars), ... — the reference's one-column conveniences (models.py:360-388)
for _column in ("mass", "initial_mass", "radius", "Teff", "logg", "feh", "density", "nu_max", "delta_nu"):
    setattr(ModelGridInterpolator, _column, _column_shortcut(_column))
ModelGridInterpolator._prop = lambda self, prop, *pars: self.interp_value(pars, [prop]).squeeze()


class EvolutionTrackInterpolator(ModelGridInterpolator):
    param_names = ("mass", "eep", "feh", "distance", "AV")      # models.py:665
    eep_replaces = "age"
    _param_index_order = (2, 0, 1, 3, 4)                         # models.py:669


class IsochroneInterpolator(ModelGridInterpolator):
    param_names = ("eep", "age", "feh", "distance", "AV")       # models.py:692
    eep_replaces = "mass"
    _param_index_order = (1, 2, 0, 3, 4)                         # models.py:696


def _from_synthetic(cls, model, bc, bands=None, eep_bounds=None, ctx=None):
    mi = DFInterpolator.from_arrays(model["grid"], model["axes"], model["columns"], ctx=ctx)
    bi = DFInterpolator.from_arrays(bc["grid"], bc["axes"], bc["columns"], ctx=ctx)
    replaces = cls.eep_replaces
    mg = ModelGrid(mi, model["limits"], replaces, name="synthetic_" + model.get("kind", "grid"))
    if eep_bounds is None:
        eep_bounds = tuple(model["limits"]["eep"])
    return cls(mg, BCGrid(bi, bc["columns"]), bands=bands, eep_bounds=eep_bounds, ctx=ctx)


def ichrone_from_arrays(kind, model, bc, bands=None, eep_bounds=None, ctx=None):
    """Interpolator over in-memory dense grids: ``model`` / ``bc`` are dicts with ``grid``, ``axes``, ``columns``
    (+ ``limits`` for the model grid), e.g. from :mod:`isochrones_b200.synthetic`; ``kind`` is "track" or "iso"."""
    cls = EvolutionTrackInterpolator if kind == "track" else IsochroneInterpolator
    return _from_synthetic(cls, model, bc, bands=bands, eep_bounds=eep_bounds, ctx=ctx)


def get_ichrone(models="mist", bands=None, tracks=False, synthetic=False, synthetic_shape=None, root=None, ctx=None,
                limits=None, **kwargs):
    """``get_ichrone`` of the reference (isochrone.py:47-78) for the one grid family on the path (MIST).

    By default the grids are the REAL ones: the dense-grid cache the reference leaves under ``$ISOCHRONES``
    (``mist/full_grid_*.npz`` or ``mist/tracks/full_grid_*.npz``) and the MIST bolometric-correction tables under
    ``$ISOCHRONES/BC/mist`` (:mod:`isochrones_b200.mistio`); ``root`` overrides ``$ISOCHRONES``.  Without that data
    this raises :class:`mistio.MistDataNotFound` — nothing is fabricated silently.  ``limits`` overrides entries of the
    MIST parameter limits (mist/models.py:37); other keywords select the grid version (``version``, ``vvcrit``,
    ``iso_kind``).  ``synthetic=True`` (or a
    ``synthetic_shape`` dict of ``make_*_grid`` keyword arguments, which shrinks them for tests) asks for the
    deterministic MIST-*shaped* synthetic grids of the benchmark; their interpolator is named ``synthetic_*``."""
    if not (isinstance(models, str) and models.lower().startswith("mist")):
        raise ValueError("only the MIST grid family is on the accelerated path")
    kind = "track" if tracks else "iso"
    if synthetic or synthetic_shape is not None:
        from . import synthetic as syn

        bands = list(bands) if bands is not None else ["G", "BP", "RP", "J", "H", "K", "W1", "W2", "W3", "TESS", "Kepler"]
        shape = dict(synthetic_shape or {})
        bc = syn.make_bc_grid(bands=tuple(bands), **shape.get("bc", {}))
        model = syn.make_track_grid(**shape.get("track", {})) if tracks else syn.make_iso_grid(**shape.get("iso", {}))
        return ichrone_from_arrays(kind, model, bc, bands=bands, ctx=ctx)
    from . import mistio

    bands = list(bands) if bands is not None else list(mistio.DEFAULT_BANDS)     # mist/bc.py:159
    model = mistio.load_model_grid(kind, root=root, limits=limits, **kwargs)
    bc = mistio.load_bc_grid(bands, root=root)
    cls = EvolutionTrackInterpolator if tracks else IsochroneInterpolator
    mi = DFInterpolator.from_arrays(model["grid"], model["axes"], model["columns"], index_names=model["index_names"], ctx=ctx)
    bi = DFInterpolator.from_arrays(bc["grid"], bc["axes"], bc["columns"], index_names=("Teff", "logg", "[Fe/H]", "Av"), ctx=ctx)
    mg = ModelGrid(mi, model["limits"], cls.eep_replaces, name="mist")
    mg.source = model["source"]
    return cls(mg, BCGrid(bi, bc["columns"]), bands=bands, eep_bounds=tuple(model["limits"]["eep"]), ctx=ctx)
