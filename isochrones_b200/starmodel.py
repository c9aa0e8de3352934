"""``BasicStarModel`` (+ ``SingleStarModel`` / ``BinaryStarModel`` / ``TripleStarModel``) — drop-in for the lnpost
path of the reference's ``isochrones/starmodel.py:1361-2007``.

Same constructor keywords, ``param_names``, ``bands``, ``spec_props``, ``bounds``, ``set_prior``, ``set_bounds``,
``lnprior(p)``, ``lnlike(p)``, ``lnpost(p)``, ``mnest_prior``, ``mnest_loglike`` and ``sample_from_prior``.  The
scalar calls are batches of one; ``lnpost_batch(P[N, ndim])`` is the new batched entry the samplers should use:
one fused CUDA launch per batch (``iso_lnpost_batch``).  The observation tuples and the prior objects are
compiled once into a device struct (``iso_model``) and re-compiled only when ``set_prior`` / ``set_bounds`` change
them.  Priors the device cannot evaluate raise at compile time — there is no CPU fallback.

Out of scope: ini parsing, HDF save/load, corner plots, the obs-tree ``StarModel``.
"""
import ctypes as C
import threading

import numpy as np

from . import _lib
from .priors import AgePrior, AVPrior, ChabrierPrior, DistancePrior, EEP_prior, FehPrior


class CompiledModel(object):
    """Device image of one or more star models that share an interpolator (catalog mode: one per star)."""

    def __init__(self, ic, structs, bands_key):
        self._stage(ic, (_lib.IsoModel * len(structs))(*structs), len(structs), structs[0].n_stars, bands_key)

    @classmethod
    def from_struct_array(cls, ic, arr, n_models, n_stars, bands_key):
        """From a ready ctypes array of ``iso_model`` (the catalog driver packs thousands of stars vectorised)."""
        self = cls.__new__(cls)
        self._stage(ic, arr, n_models, n_stars, bands_key)
        return self

    def _stage(self, ic, arr, n_models, n_stars, bands_key):
        self.ic = ic
        self.ctx = ic.ctx
        self.n_models = int(n_models)
        self.n_stars = int(n_stars)
        self.ndim = 4 + self.n_stars
        self.model_pack = ic.model_pack
        self.bc_pack = ic.bc_pack(tuple(bands_key))
        self.handle = C.c_void_p()
        self.ctx.check(_lib.lib().iso_models_stage(self.ctx.handle, arr, self.n_models, C.byref(self.handle)))
        self._one = threading.local()   # per-thread buffers of the scalar call

    def lnpost_one(self, p, parts=False):
        """``lnpost`` of ONE parameter vector — the reference's scalar interface, as a sampler calls it millions of
        times (``parts``: the ``(lnpost, lnprior, lnlike)`` triple instead).  The row and the results live in small
        page-locked buffers (one set per host thread) whose ctypes pointers are cached, so the call is: five stores,
        one C call (kernel reads / writes the pinned buffers directly: the library's small-call path), one load."""
        one = self._one
        buf = getattr(one, "buf", None)
        if buf is None:
            if self.n_models != 1:
                raise ValueError("model_of_row is required when several models are compiled together")
            h_in, h_out = self.ctx.pinned_empty((1, self.ndim)), self.ctx.pinned_empty((3,))
            buf = one.buf = (h_in, h_out, _lib.dp(h_in), _lib.dp(h_out[0:1]), _lib.dp(h_out[1:2]), _lib.dp(h_out[2:3]),
                             _lib.lib().iso_lnpost_batch, self.ctx.handle, self.model_pack.handle, self.bc_pack.handle)
        h_in, h_out, p_in, p_post, p_prior, p_like, fn, ctxh, mph, bph = buf
        h_in[0, :] = p          # raises on a wrong length
        if parts:
            rc = fn(ctxh, mph, bph, self.handle, None, p_in, 1, p_post, p_prior, p_like)
        else:
            rc = fn(ctxh, mph, bph, self.handle, None, p_in, 1, p_post, None, None)
        if rc:
            self.ctx.check(rc)
        if parts:
            return float(h_out[0]), float(h_out[1]), float(h_out[2])
        return float(h_out[0])

    def cube_one(self, cube, lo, hi):
        """MultiNest's per-live-point call pair in ONE launch: the unit-cube point ``cube`` (any indexable of ``ndim``
        numbers) is mapped through the box ``[lo, hi]`` and evaluated (``iso_mnest_lnpost_batch`` on a page-locked row:
        the library's small-call path).  Returns ``(mapped parameters as a list, lnpost)``."""
        one = self._one
        buf = getattr(one, "cube", None)
        if buf is None:
            if self.n_models != 1:
                raise ValueError("unit-cube rows need a single compiled model")
            h_row, h_out = self.ctx.pinned_empty((1, self.ndim)), self.ctx.pinned_empty((1,))
            buf = one.cube = (h_row, h_out, _lib.dp(h_row), _lib.dp(h_out), _lib.lib().iso_mnest_lnpost_batch, self.ctx.handle,
                              self.model_pack.handle, self.bc_pack.handle)
        h_row, h_out, p_row, p_out, fn, ctxh, mph, bph = buf
        for i in range(self.ndim):
            h_row[0, i] = cube[i]
        rc = fn(ctxh, mph, bph, self.handle, _lib.dp(lo), _lib.dp(hi), p_row, 1, p_out, None, None)
        if rc:
            self.ctx.check(rc)
        return h_row[0].tolist(), float(h_out[0])

    def lnpost(self, pars, parts=False, model_of_row=None, out=None):
        """``pars[N, ndim]`` -> ``lnpost[N]`` (and ``lnprior[N]``, ``lnlike[N]`` when ``parts``)."""
        pars = np.asarray(pars)
        if pars.dtype != np.float64 or not pars.flags["C_CONTIGUOUS"]:
            pars = np.ascontiguousarray(pars, dtype=np.float64)
        if pars.ndim != 2 or pars.shape[1] != self.ndim:
            raise ValueError("expected pars of shape [N, %d], got %r" % (self.ndim, pars.shape))
        n = pars.shape[0]
        lnpost = out if out is not None else np.empty(n)
        mor = None
        if model_of_row is not None:
            mor = np.ascontiguousarray(model_of_row, dtype=np.int32)
            assert mor.shape == (n,)
        elif self.n_models != 1:
            raise ValueError("model_of_row is required when several models are compiled together")
        lnprior = np.empty(n) if parts else None
        lnlike = np.empty(n) if parts else None
        self.ctx.check(_lib.lib().iso_lnpost_batch(
            self.ctx.handle, self.model_pack.handle, self.bc_pack.handle, self.handle,
            _lib.ip(mor) if mor is not None else None, _lib.dp(pars), n, _lib.dp(lnpost),
            _lib.dp(lnprior) if parts else None, _lib.dp(lnlike) if parts else None))
        if parts:
            return lnpost, lnprior, lnlike
        return lnpost

    def lnpost_device(self, d_pars, n, d_lnpost, d_lnprior=None, d_lnlike=None, d_model_of_row=None):
        """Asynchronous launch on device-resident buffers (``iso_lnpost_batch_device``): ``d_*`` are device
        pointers from ``Context.dev_alloc``; the call returns as soon as the kernel is queued."""
        self.ctx.check(_lib.lib().iso_lnpost_batch_device(
            self.ctx.handle, self.model_pack.handle, self.bc_pack.handle, self.handle, d_model_of_row, d_pars, int(n),
            d_lnpost, d_lnprior, d_lnlike))

    def close(self):
        if self.handle:
            _lib.lib().iso_models_destroy(self.ctx.handle, self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class BasicStarModel(object):
    """Bare-bones star model without the observation-tree machinery (starmodel.py:1361-1989)."""

    use_emcee = False
    _not_a_band = ("RA", "dec", "ra", "Dec", "maxAV", "parallax", "AV", "logg", "Teff", "feh", "density",
                   "separation", "PA", "resolution", "relative", "N", "index", "id", "nu_max", "delta_nu")

    def __init__(self, ic, eep_bounds=None, name="", directory=".", N=1, maxAV=None, max_distance=None,
                 halo_fraction=None, ra=None, dec=None, obs=None, use_emcee=False, **kwargs):
        if N not in (1, 2, 3):
            raise ValueError("N must be 1, 2 or 3")
        self._ic, self.N = ic, N
        if N > 1 and self.ic.eep_replaces == "age":
            raise ValueError("binary / triple fits need the isochrone-grid interpolator (per-star EEPs, one shared age)")
        self.name, self._directory = str(name), str(directory)
        self.use_emcee, self.ra, self.dec, self.obs = use_emcee, ra, dec, None
        self.eep_bounds = self.ic.eep_bounds if eep_bounds is None else eep_bounds
        self._samples = self._derived_samples = None
        self._bands = self._spec_props = self._props = self._param_names = self._compiled = None

        self._number_row_slots()
        self.kwargs = self._measurements(kwargs)
        self._install_default_priors(eep_bounds)
        self._clip_box(maxAV, max_distance)
        if halo_fraction is not None:
            # as in the reference (starmodel.py:1478-1479) the two-population [Fe/H] prior keeps ITS OWN default support:
            # it is installed after the box was clipped to the grid
            self._priors["feh"] = FehPrior(halo_fraction=halo_fraction)
            self._priors["feh"]._owner = self.ic

    def _number_row_slots(self):
        """``<par>_index`` attributes (starmodel.py:1403-1421): where each shared parameter sits in a row.  The row starts
        with the per-star slots — N EEPs, or the one mass of a track model — so a shared parameter at position ``j`` of the
        interpolator's own list lands at ``j + N - 1``."""
        for j, par in enumerate(self.ic.param_names):
            if par != "eep":
                setattr(self, par + "_index", j + (self.N - 1 if j else 0))

    @staticmethod
    def _measurements(given):
        """Keyword observations -> ``{name: (value, sigma)}`` in float64.  Anything that does not unpack as a pair is
        dropped, and so is a pair with a NaN in it: "not observed" (starmodel.py:1423-1432, which warns instead)."""
        kept = {}
        for key, item in given.items():
            try:
                value, sigma = item
            except TypeError:
                continue
            pair = (np.float64(value), np.float64(sigma))
            if key != "use_emcee" and not np.isnan(pair).any():
                kept[key] = pair
        return kept

    def _install_default_priors(self, eep_bounds):
        """The reference's default prior set (starmodel.py:1441-1460) and the sampling box that goes with it: mass, [Fe/H]
        and age take the model grid's own limits, distance and A_V their prior's support, EEP the requested bounds."""
        shared = {"mass": ChabrierPrior(), "feh": FehPrior(), "age": AgePrior(), "distance": DistancePrior(), "AV": AVPrior()}
        shared["eep"] = EEP_prior(self.ic, shared[self.ic.eep_replaces], bounds=eep_bounds)
        self._priors, self._bounds = shared, {}
        for par, pr in shared.items():
            pr._owner = self.ic          # stand-alone evaluations (prior.lnpdf(x)) run on the model's own GPU context
            self._bounds[par] = None if par in self._grid_limited else pr.bounds
        for par in self._grid_limited:
            self.bounds(par)

    def _clip_box(self, maxAV, max_distance):
        """Optional tightening of the box (starmodel.py:1462-1476): explicit caps first; without a distance cap a
        measured parallax bounds the distance at twice the parallax distance (of the value, or of |sigma| if negative)."""
        if maxAV is not None:
            self.set_bounds(AV=(0, maxAV))
        plx = self.kwargs.get("parallax")
        if max_distance is not None:
            self.set_bounds(distance=(0, max_distance))
        elif plx is not None and plx[0] != 0:
            self.set_bounds(distance=(0, 1.0 / (plx[0] if plx[0] > 0 else np.abs(plx[1])) * 2000))

    _grid_limited = ("mass", "feh", "age")

    # ---- reference attribute surface ---------------------------------------------------------------------
    @property
    def ic(self):
        if type(self._ic) == type:
            self._ic = self._ic()
        return self._ic

    labelstring = property(lambda self: ("single", "binary", "triple")[self.N - 1])

    @property
    def param_names(self):
        """Row layout: the interpolator's names for one star; ``eep_0 .. eep_{N-1}`` then the shared ones for N > 1."""
        if self._param_names is None:
            base = tuple(self.ic.param_names)
            self._param_names = base if self.N == 1 else tuple("eep_%d" % k for k in range(self.N)) + base[1:]
        return self._param_names

    def _cached(self, attr, build):
        if getattr(self, attr) is None:
            setattr(self, attr, build())
        return getattr(self, attr)

    bands = property(lambda self: self._cached("_bands", lambda: [k for k in self.kwargs if k in self.ic.bc_grid.bands]))
    props = property(lambda self: self._cached("_props", lambda: [k for k in self.kwargs if k in self._not_a_band]))
    spec_props = property(lambda self: self._cached(
        "_spec_props", lambda: [self.kwargs.get(k, (np.nan, np.nan)) for k in ("Teff", "logg", "feh")]))

    @property
    def n_params(self):
        return len(self.param_names)

    def bounds(self, prop):
        """(lo, hi) of one row parameter (starmodel.py:1538-1552); ``eep_k`` share the ``eep`` entry.  The grid-limited
        parameters resolve lazily to the model grid's limits, which also become the support of their prior."""
        key = "eep" if prop in ("eep_0", "eep_1", "eep_2") else prop
        box = self._bounds.get(key)
        if box is None:
            if key not in self._grid_limited:
                raise ValueError("Unknown property {}".format(prop))
            box = tuple(self.ic.model_grid.get_limits(key))
            self._bounds[key] = self._priors[key].bounds = box
            self._compiled = None
        return box

    def set_bounds(self, **kwargs):
        self._mnest_last = None
        for k, v in kwargs.items():
            if len(v) != 2:
                raise ValueError("Must provide (min, max)")
            self._bounds[k] = v
            self._priors[k].bounds = v
        self._compiled = None

    def set_prior(self, **kwargs):
        self._mnest_last = None
        for prop, prior in kwargs.items():
            self._priors[prop] = prior
            self._bounds[prop] = prior.bounds
            prior._owner = self.ic
        self._compiled = None

    def prior(self, prop, val, **kwargs):
        return self._priors[prop](val, **kwargs)

    # ---- compilation to the device struct ------------------------------------------------------------------
    def to_struct(self, band_columns=None):
        """``iso_model`` image: observations (starmodel.py:1574-1580, 1599-1612) and the prior objects
        (starmodel.py:1441-1448).  ``band_columns`` maps band name -> BC-pack column (catalog mode)."""
        ic = self.ic
        s = _lib.IsoModel()
        s.n_stars = self.N
        s.eep_replaces_age = 1 if ic.eep_replaces == "age" else 0
        for i, v in enumerate(ic.param_index_order):
            s.index_order[i] = int(v)
        bands = list(self.bands)
        if len(bands) > _lib.ISO_MAX_BANDS:
            raise ValueError("at most %d bands per star model" % _lib.ISO_MAX_BANDS)
        s.n_bands = len(bands)
        for i, b in enumerate(bands):
            s.band_col[i] = i if band_columns is None else band_columns[b]
            s.mag_val[i], s.mag_unc[i] = [float(v) for v in self.kwargs[b]]
        for i, (val, unc) in enumerate(self.spec_props):
            s.spec_val[i], s.spec_unc[i] = float(val), float(unc)
        for key, flag in (("parallax", "has_plax"), ("nu_max", "has_nu_max"), ("delta_nu", "has_delta_nu")):
            setattr(s, flag, 1 if key in self.kwargs else 0)
        s.plax, s.plax_unc = [float(v) for v in self.kwargs.get("parallax", (np.nan, np.nan))]
        s.nu_max, s.nu_max_unc = [float(v) for v in self.kwargs.get("nu_max", (np.nan, np.nan))]
        s.delta_nu, s.delta_nu_unc = [float(v) for v in self.kwargs.get("delta_nu", (np.nan, np.nan))]
        if s.has_nu_max:
            ci = ic.model_grid.interp.column_index
            if "nu_max" not in ci or "delta_nu" not in ci:
                raise KeyError("model grid has no nu_max / delta_nu columns")
        eep = self._priors["eep"]
        if not isinstance(eep, EEP_prior):
            raise TypeError("the 'eep' prior must be an EEP_prior (it is evaluated inside the fused kernel)")
        s.eep_has_bounds = 0 if eep._bounds is None else 1
        if eep._bounds is not None:
            s.eep_lo, s.eep_hi = float(eep._bounds[0]), float(eep._bounds[1])
        s.eep_norm = float(eep._norm)
        s.eep_orig = eep.orig_prior.to_struct()
        for k in ("mass", "age", "feh", "distance", "AV"):
            setattr(s, k, self._priors[k].to_struct())
        return s

    @property
    def compiled(self):
        if self._compiled is None:
            self._compiled = CompiledModel(self.ic, [self.to_struct()], tuple(self.bands))
        return self._compiled

    # ---- the hot path ------------------------------------------------------------------------------------------
    def lnpost_batch(self, pars, parts=False, out=None):
        """Batched ``lnpost`` over rows of ``pars[N, n_params]`` — one fused kernel launch per chunk."""
        return self.compiled.lnpost(pars, parts=parts, out=out)

    def lnprior_batch(self, pars):
        return self.compiled.lnpost(pars, parts=True)[1]

    def lnlike_batch(self, pars):
        return self.compiled.lnpost(pars, parts=True)[2]

    def lnlike(self, pars):
        return self.compiled.lnpost_one(pars, parts=True)[2]

    def lnprior(self, pars):
        return self.compiled.lnpost_one(pars, parts=True)[1]

    def lnpost(self, p, **kwargs):
        """``lnprior + lnlike`` with ``-inf`` when the prior is not finite (starmodel.py:538-542)."""
        return self.compiled.lnpost_one(p)

    def mnest_prior(self, cube, ndim=None, nparams=None):
        """Unit cube -> parameter box, in place (starmodel.py:1637-1640); ``cube`` may be ``[ndim]`` — a list, an array
        or the ctypes ``double *`` pymultinest hands over — or a float64 array ``[N, ndim]``.

        MultiNest calls ``mnest_prior(cube)`` and then ``mnest_loglike(cube)`` for every live point.  The single-point
        form therefore runs the fused cube kernel (mapping + lnpost in one launch) and remembers the lnpost of the
        mapped point; the ``mnest_loglike`` that follows finds its cube unchanged and returns it without a second
        launch."""
        lo, hi = self._box()
        if isinstance(cube, np.ndarray) and cube.ndim == 2:
            if cube.dtype != np.float64 or not cube.flags["C_CONTIGUOUS"]:
                raise ValueError("a batch of cube points must be a C-contiguous float64 array (it is mapped in place)")
            ctx = self.ic.ctx
            ctx.check(_lib.lib().iso_mnest_prior(ctx.handle, _lib.dp(lo), _lib.dp(hi), len(lo), _lib.dp(cube), cube.shape[0]))
            return
        mapped, lnpost = self.compiled.cube_one(cube, lo, hi)
        for i, v in enumerate(mapped):
            cube[i] = v
        self._mnest_last = (mapped, lnpost)

    def mnest_loglike(self, cube, ndim=None, nparams=None):
        vals = [cube[i] for i in range(self.n_params)]
        last = getattr(self, "_mnest_last", None)
        if last is not None and last[0] == vals:      # the point mnest_prior has just mapped (and evaluated)
            return last[1]
        return self.lnpost(vals)

    def _box(self):
        """``(lo, hi)`` of every parameter; cached on the compiled model, which is dropped whenever a bound changes."""
        c = self.compiled
        box = getattr(c, "box", None)
        if box is None:
            lo = np.array([self.bounds(par)[0] for par in self.param_names], dtype=np.float64)
            hi = np.array([self.bounds(par)[1] for par in self.param_names], dtype=np.float64)
            if self._compiled is not c:       # bounds() of an unset parameter re-compiles: take the current model
                c = self.compiled
            box = c.box = (lo, hi)
        return box

    def mnest_lnpost_batch(self, cube, parts=False):
        """``mnest_prior`` + ``mnest_loglike`` (starmodel.py:1637-1645) of a whole set of live points in ONE launch:
        ``cube[N, ndim]`` (float64, C-contiguous) holds unit-cube points on entry and the mapped parameters on return —
        exactly what ``mnest_prior`` leaves in MultiNest's cube — and the returned ``lnpost[N]`` is ``mnest_loglike`` of
        every row (``parts``: also lnprior and lnlike).  ``iso_mnest_lnpost_batch``."""
        if not (isinstance(cube, np.ndarray) and cube.dtype == np.float64 and cube.flags["C_CONTIGUOUS"]
                and cube.ndim == 2 and cube.shape[1] == self.n_params):
            raise ValueError("cube must be a C-contiguous float64 array of shape [N, %d] (it is mapped in place)" % self.n_params)
        lo, hi = self._box()
        c = self.compiled
        n = cube.shape[0]
        lnpost = np.empty(n)
        lnprior = np.empty(n) if parts else None
        lnlike = np.empty(n) if parts else None
        c.ctx.check(_lib.lib().iso_mnest_lnpost_batch(
            c.ctx.handle, c.model_pack.handle, c.bc_pack.handle, c.handle, _lib.dp(lo), _lib.dp(hi), _lib.dp(cube), n,
            _lib.dp(lnpost), _lib.dp(lnprior) if parts else None, _lib.dp(lnlike) if parts else None))
        return (lnpost, lnprior, lnlike) if parts else lnpost

    def prior_box_draws(self, n, seed=0, row0=0, return_pars=True, out=None):
        """``n`` uniform draws from the parameter box (``bounds`` of every parameter) evaluated on the device where they
        are made (``iso_lnpost_prior_draws``): nothing is shipped per row.  Row ``i`` is the Philox4x32-10 point of
        counter ``row0 + i`` under ``seed`` — the same rows whatever the batch split or the number of GPUs.  Returns
        ``(pars[n, ndim], lnpost[n])`` or just ``lnpost`` (``return_pars=False``); the live-point / walker
        initialisation of ``sample_from_prior`` (starmodel.py:1716-1748) keeps the finite ones."""
        lo, hi = self._box()
        c = self.compiled
        lnpost = out if out is not None else np.empty(n)
        pars = np.empty((n, self.n_params)) if return_pars else None
        c.ctx.check(_lib.lib().iso_lnpost_prior_draws(
            c.ctx.handle, c.model_pack.handle, c.bc_pack.handle, c.handle, _lib.dp(lo), _lib.dp(hi), C.c_uint64(int(seed)),
            int(row0), int(n), _lib.dp(pars) if return_pars else None, _lib.dp(lnpost)))
        return (pars, lnpost) if return_pars else lnpost

    def fit_mcmc(self, nwalkers=200, nburn=100, niter=200, p0=None, seed=0, thin=1, a=2.0):
        """emcee-style fit on the device (the reference's ``fit_mcmc_old``, starmodel.py:889-972: burn-in, reset,
        production run) — two kernel launches in total.  Initial walkers come from ``sample_from_prior`` as the
        reference's emcee3 driver does (``fit.py:86``).  Returns the ``DeviceEnsembleSampler`` (``.chain``,
        ``.lnprobability``, ``.acceptance_fraction`` as in emcee 2.x)."""
        from .sampler import DeviceEnsembleSampler

        if p0 is None:
            p0 = self.sample_from_prior(nwalkers, values=True, require_valid=True)
        sampler = DeviceEnsembleSampler(self.compiled, nwalkers, p0, seed=seed, a=a)
        if nburn > 0:
            sampler.run_mcmc(nburn, store=False)
            sampler.reset()
        sampler.run_mcmc(niter, thin=thin)
        self._sampler = sampler
        self._samples = self._derived_samples = None
        self._sample_source = "mcmc"
        return sampler

    def fit_nested(self, n_live_points=1000, evidence_tolerance=0.5, seed=0, **kwargs):
        """Nested sampling of this model — the reference's default ``fit()`` path (``fit_multinest``,
        starmodel.py:717-802: ``pymultinest.run(self.mnest_loglike, self.mnest_prior, ...)`` with
        ``n_live_points=1000``, ``evidence_tolerance=0.5``) with every batch of likelihood evaluations as one GPU launch
        (:mod:`isochrones_b200.nested`).  Returns a ``NestedResult`` (``logZ``, weighted ``samples`` ...)."""
        from .nested import nested_sample

        self._nested = nested_sample(self, n_live=n_live_points, dlogz=evidence_tolerance, seed=seed, **kwargs)
        self._samples = self._derived_samples = None
        self._sample_source = "nested"
        return self._nested

    # ---- posterior samples and what follows from them ------------------------------------------
    @property
    def samples(self):
        """Equal-weight posterior samples of the last fit as a DataFrame: the parameter columns + ``lnprob`` — what the
        reference reads back from MultiNest's ``post_equal_weights.dat`` (starmodel.py:1652-1660).  After ``fit_nested``
        the weighted points are resampled to equal weights; after ``fit_mcmc`` it is the flattened production chain."""
        import pandas as pd

        if self._samples is None:
            source = getattr(self, "_sample_source", None)
            if source == "nested":
                res = self._nested
                n = max(1, int(round(1.0 / np.sum(res.weights ** 2))))
                pos = (np.random.default_rng(0).random() + np.arange(n)) / n
                pick = np.minimum(np.searchsorted(np.cumsum(res.weights), pos), len(res.weights) - 1)
                values, lnprob = res.samples[pick], res.lnpost[pick]
            elif source == "mcmc":
                values = self._sampler.chain.reshape(-1, self.n_params)
                lnprob = self._sampler.lnprobability.reshape(-1)
            else:
                raise AttributeError("no samples yet: run fit_nested() or fit_mcmc() first")
            df = pd.DataFrame(values, columns=list(self.param_names))
            df["lnprob"] = lnprob
            self._samples = df
        return self._samples

    @property
    def derived_samples(self):
        """The table ``derive(samples)`` of the last fit (starmodel.py:1646-1650)."""
        if self._derived_samples is None:
            self._derived_samples = self.derive(self.samples)
        return self._derived_samples

    def derive(self, samples):
        """Posterior samples -> every model-grid column and every band's magnitude of every star, the reference's
        derived-sample table (``_make_samples``, starmodel.py:1662-1714) — two batched launches per star (all-column
        ``interp_values`` + ``interp_mags``) instead of a DataFrame round trip per call.

        ``samples``: DataFrame with the ``param_names`` columns (others, e.g. ``lnprob``, are carried along for
        multi-star models, as the reference does), or an ``[n, n_params]`` array.  Columns follow the reference: single
        star = the interpolator's table (grid columns, ``<band>_mag``); N stars = the sample columns, then per star k
        ``<column>_k`` / ``<band>_mag_k`` (without eep, age, distance, AV), then the combined ``<band>_mag``; always
        ``parallax = 1000 / distance``, ``distance``, ``AV`` last."""
        import pandas as pd

        names = list(self.param_names)
        if not isinstance(samples, pd.DataFrame):
            samples = pd.DataFrame(np.atleast_2d(np.asarray(samples, dtype=np.float64)), columns=names)
        col = {c: samples[c].to_numpy(dtype=np.float64) for c in names}
        if self.N == 1:
            table = self.ic(*[col[c] for c in names])
        else:
            shared = names[self.N:]                                 # age, feh, distance, AV
            pieces = [samples.reset_index(drop=True)]
            flux = {b: 0.0 for b in self.bands}
            for k in range(self.N):
                star = self.ic(col["eep_%d" % k], *[col[c] for c in shared])
                for b in self.bands:                                # utils.py:43-58: fluxes add
                    flux[b] = flux[b] + 10.0 ** (-0.4 * star[b + "_mag"].to_numpy())
                star = star.drop(columns=["eep", "age"])
                star.columns = ["%s_%d" % (c, k) for c in star.columns]
                pieces.append(star)
            table = pd.concat(pieces, axis=1)
            for b in self.bands:
                table[b + "_mag"] = -2.5 * np.log10(flux[b])
        table["parallax"] = 1000.0 / col["distance"]
        table["distance"] = col["distance"]
        table["AV"] = col["AV"]
        return table

    def sample_from_prior(self, n, values=False, require_valid=True):
        """Prior draws, re-drawn until ``lnpost`` is finite (starmodel.py:1716-1748); host RNG, batched validity check."""
        import pandas as pd

        if n == 0:
            return pd.DataFrame(columns=self.param_names)
        names = list(self.param_names)
        cols = {}
        for p in names:
            if not p.startswith("eep"):
                cols[p] = self._priors[p].sample(n)
        for p in names:
            if p.startswith("eep"):
                if self.ic.eep_replaces == "age":
                    cols[p] = self._priors["eep"].sample(n, mass=cols["mass"], feh=cols["feh"])
                else:
                    cols[p] = self._priors["eep"].sample(n, age=cols["age"], feh=cols["feh"])
        if self.N > 1:      # keep the ordering the prior demands (starmodel.py:1618-1623)
            e = np.sort(np.array([cols[p] for p in names[:self.N]]), axis=0)[::-1]
            for k in range(self.N):
                cols[names[k]] = e[k]
        df = pd.DataFrame({p: np.asarray(cols[p], dtype=float) for p in names})
        if require_valid:
            bad = ~np.isfinite(self.lnpost_batch(df[names].values))
            if bad.any():
                new = self.sample_from_prior(int(bad.sum()), require_valid=True)
                df.loc[bad, names] = new[names].values
        return df[names].values if values else df


def _fixed_multiplicity(name, n_stars):
    """``BasicStarModel`` with the number of stars pinned (the reference's Single / Binary / TripleStarModel)."""
    def __init__(self, *args, **kwargs):
        BasicStarModel.__init__(self, *args, **dict(kwargs, N=n_stars))

    return type(name, (BasicStarModel,), {"__init__": __init__, "__doc__": "%d-star ``BasicStarModel``." % n_stars})


SingleStarModel = _fixed_multiplicity("SingleStarModel", 1)
BinaryStarModel = _fixed_multiplicity("BinaryStarModel", 2)
TripleStarModel = _fixed_multiplicity("TripleStarModel", 3)


def compile_catalog(models):
    """Catalog mode: compile many star models that share one interpolator into a single device
    array; returns ``(CompiledModel, bands)`` where row ``i`` of a batch selects its model through ``model_of_row``."""
    ic = models[0].ic
    bands = []
    for m in models:
        if m.ic is not ic:
            raise ValueError("catalog models must share one interpolator")
        for b in m.bands:
            if b not in bands:
                bands.append(b)
    col = {b: i for i, b in enumerate(bands)}
    structs = [m.to_struct(band_columns=col) for m in models]
    return CompiledModel(ic, structs, tuple(bands)), bands
