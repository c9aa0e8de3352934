"""Prior objects of the lnpost path — the product's counterpart of the reference's ``isochrones/priors.py``.

The public surface is the reference's (class names, constructor arguments, ``bounds`` / ``_bounds`` / ``_norm`` and the
per-class attributes ``alpha``, ``mean``, ``sigma``, ``norm``, ``lognorm``, ``mu``, ``scale``, ``log_s``,
``halo_fraction``, ``local``, ``components``, ``breakpoints``, ``norms``, ``lognorms``, ``orig_prior`` ...), because
``BasicStarModel`` users hand these objects to ``set_prior``.  The implementation is organised differently:

* **Evaluation is a device job.**  ``prior(x)``, ``prior.pdf(x)`` and ``prior.lnpdf(x)`` pack the object into the
  ``iso_prior`` struct of the C ABI and run ``iso_prior_eval`` on the GPU (scalars or arrays).  The evaluation rules —
  which classes test their bounds where, ``log(pdf) if pdf else -inf`` (priors.py:54-66, 112-140, 205-211 of the
  reference) — live in ``csrc/iso_prior.cuh``, once.  There is no CPU evaluation path behind the public methods, and
  a class without a device kind cannot be compiled into a model (``TypeError``).
* **Construction-time constants are a host job**, as in the reference: normalisation integrals use
  ``scipy.integrate.quad`` over ``density(x)``, the plain host formula of each class (``_norm`` of unbounded-support
  priors, the piece norms of a ``BrokenPrior``, priors.py:42-49, 176-203).
* **Sampling** (``sample(n)``, needed for initial walkers) is one generic numerical inverse-CDF over the host density
  instead of per-class formulas.
"""
import ctypes as C

import numpy as np
from scipy.integrate import quad
from scipy.special import ndtr

from . import _lib

_SQRT_2PI = float(np.sqrt(2.0 * np.pi))
_INF = float("inf")


def _scalar_like(x):
    return np.ndim(x) == 0


def _gauss(x, mu, sig):
    """Normal density N(mu, sig) on the host."""
    z = (x - mu) / sig
    return np.exp(-0.5 * z * z) / (_SQRT_2PI * sig)


# ---------------------------------------------------------------------------------------------------------------------
# base machinery
# ---------------------------------------------------------------------------------------------------------------------
class Prior(object):
    """A one-dimensional prior.  Subclasses give ``KIND`` (the device kind), ``_device_params()`` and the host density
    ``_pdf``; ``CLIPS`` says whether the class is a ``BoundedPrior`` (hard −inf outside ``bounds``) in the reference's
    hierarchy, which changes how ``lnpdf`` treats the bounds of classes that own a closed-form ``_lnpdf``."""

    KIND = None
    CLIPS = False
    SUPPORT = (-_INF, _INF)       # where sample() may look when no bounds are set

    def __init__(self, *args, **kwargs):
        self._norm = 1.0

    # -- bounds ------------------------------------------------------------------------------------------------
    def _get_bounds(self):
        b = getattr(self, "_bounds", None)
        return (-_INF, _INF) if b is None else b

    def _set_bounds(self, new):
        # an unbounded-support density is re-normalised over the new interval (reference: Prior.bounds setter)
        self._norm = quad(self._pdf, new[0], new[1])[0]
        self._bounds = new
        self._verify_unit_integral(new)

    bounds = property(lambda self: self._get_bounds(), lambda self, new: self._set_bounds(new))

    def _verify_unit_integral(self, new):
        try:
            self.test_integral()
        except AssertionError:
            raise ValueError(f"Problem setting bounds to {new}; integral test failed.")

    # -- device side -------------------------------------------------------------------------------------------
    def _device_params(self):
        return ()

    def _device_flags(self):
        return 0

    def _leaf_struct(self):
        leaf = _lib.IsoPriorLeaf()
        leaf.kind = self.KIND
        raw = getattr(self, "_bounds", None)
        leaf.flags = self._device_flags() | (_lib.ISO_PF_BOUNDED if self.CLIPS else 0) | \
            (_lib.ISO_PF_HAS_BOUNDS if raw is not None else 0)
        leaf.lo, leaf.hi = (-np.inf, np.inf) if raw is None else (float(raw[0]), float(raw[1]))
        leaf.norm = float(getattr(self, "_norm", 1.0))
        for i, v in enumerate(self._device_params()):
            leaf.a[i] = float(v)
        return leaf

    def to_struct(self):
        """The ``iso_prior`` image of this object; ``TypeError`` when the class has no device implementation."""
        if self.KIND is None:
            raise TypeError("prior class %s has no device implementation (isochrones_b200 has no CPU fallback)"
                            % type(self).__name__)
        out = _lib.IsoPrior()
        out.self = self._leaf_struct()
        return out

    _owner = None   # the interpolator of the model that holds this prior (BasicStarModel sets it): its context evaluates

    def _on_device(self, x, which):
        ctx = self._owner.ctx if self._owner is not None else _lib.default_context()
        flat = _lib.f64(np.atleast_1d(x)).ravel()
        res = np.empty_like(flat)
        image = self.to_struct()
        ctx.check(_lib.lib().iso_prior_eval(ctx.handle, C.byref(image), which, _lib.dp(flat), flat.size, _lib.dp(res)))
        return float(res[0]) if _scalar_like(x) else res.reshape(np.shape(x))

    def __call__(self, x, **kwargs):
        return self._on_device(x, 1)

    def pdf(self, x, **kwargs):
        return self._on_device(x, 1)

    def lnpdf(self, x, **kwargs):
        return self._on_device(x, 0)

    # -- host side: densities for quadrature, sampling, self-checks ------------------------------------------------
    def _pdf(self, x, **kwargs):
        raise NotImplementedError

    def density(self, x):
        """Normalised host density with the class's bounds rule (used by quadrature and sampling only)."""
        lo, hi = self._get_bounds()
        if x < lo or x > hi:
            return 0.0
        return self._pdf(x) / self._norm

    def test_integral(self):
        lo, hi = self._get_bounds()
        assert np.isclose(1, quad(self.density, lo, hi)[0])

    def _sampling_window(self):
        lo, hi = self._get_bounds()
        slo, shi = self.SUPPORT
        return (slo if not np.isfinite(lo) else lo), (shi if not np.isfinite(hi) else hi)

    def sample(self, n):
        """``n`` draws by numerical inverse-CDF of the host density (tabulated on 8193 points, linear inversion)."""
        lo, hi = self._sampling_window()
        if not (np.isfinite(lo) and np.isfinite(hi)):
            raise NotImplementedError("no finite sampling window for %s" % type(self).__name__)
        # a window spanning decades (masses 0.1 .. 100) is tabulated on a logarithmic grid
        xs = np.geomspace(lo, hi, 8193) if (lo > 0 and hi / lo > 50) else np.linspace(lo, hi, 8193)
        with np.errstate(all="ignore"):
            dens = np.nan_to_num(np.array([self.density(float(v)) for v in xs]), nan=0.0, posinf=0.0)
        cdf = np.concatenate([[0.0], np.cumsum(0.5 * (dens[1:] + dens[:-1]) * np.diff(xs))])
        if cdf[-1] <= 0:
            raise ValueError("density of %s vanishes on its sampling window" % type(self).__name__)
        cdf /= cdf[-1]
        keep = np.concatenate([[True], np.diff(cdf) > 0])
        return np.interp(np.random.random(n), cdf[keep], xs[keep])

    def test_sampling(self, n=100000, plot=False):
        """Ten-bin histogram of ``sample(n)`` against the bin-averaged density, within 6 sigma per bin."""
        draws = self.sample(n)
        window = None if tuple(self._get_bounds()) == (-_INF, _INF) else self._get_bounds()
        counts, edges = np.histogram(draws, range=window)
        expected = np.array([quad(self.density, a, b)[0] for a, b in zip(edges[:-1], edges[1:])]) * n
        busy = counts > 50
        pull = np.abs(counts[busy] - expected[busy]) / np.sqrt(np.maximum(expected[busy], 1.0))
        assert pull.max() < 6


class BoundedPrior(Prior):
    """A prior that is exactly zero (lnpdf −inf) outside ``bounds``; re-bounding does NOT re-normalise."""

    CLIPS = True

    def __init__(self, bounds=None):
        super().__init__()
        self._bounds = bounds

    def _get_bounds(self):
        return self._bounds

    def _set_bounds(self, new):
        self._bounds = new
        self._verify_unit_integral(new)

    bounds = property(lambda self: self._get_bounds(), lambda self, new: self._set_bounds(new))

    def density(self, x):
        b = self._bounds
        if b is not None and (x < b[0] or x > b[1]):
            return 0.0
        return self._pdf(x) / self._norm

    def test_integral(self):
        lo, hi = self._bounds if self._bounds is not None else (-_INF, _INF)
        assert np.isclose(1, quad(self.density, lo, hi)[0])

    def _sampling_window(self):
        return self._bounds if self._bounds is not None else self.SUPPORT


# ---------------------------------------------------------------------------------------------------------------------
# elementary densities
# ---------------------------------------------------------------------------------------------------------------------
class FlatPrior(BoundedPrior):
    KIND = _lib.ISO_PRIOR_FLAT

    def __init__(self, bounds):
        BoundedPrior.__init__(self, bounds=bounds)

    def _pdf(self, x):
        return 1.0 / (self._bounds[1] - self._bounds[0])


class FlatLogPrior(BoundedPrior):
    """Flat in 10**x: the density of x = log10(t) for t uniform."""

    KIND = _lib.ISO_PRIOR_FLATLOG

    def __init__(self, bounds):
        BoundedPrior.__init__(self, bounds=bounds)

    def _pdf(self, x):
        lo, hi = self._bounds
        return np.log(10) * 10 ** x / (10 ** hi - 10 ** lo)


class PowerLawPrior(BoundedPrior):
    KIND = _lib.ISO_PRIOR_POWERLAW

    def __init__(self, alpha, bounds=None):
        BoundedPrior.__init__(self, bounds=bounds)
        self.alpha = alpha

    def _device_params(self):
        return (self.alpha,)

    def _pdf(self, x):
        lo, hi = (np.float64(v) for v in self._bounds)
        k = 1.0 + self.alpha
        with np.errstate(divide="ignore"):
            return k / (hi ** k - lo ** k) * np.float64(x) ** self.alpha


class GaussianPrior(BoundedPrior):
    """Normal density, truncated (and re-normalised through ``norm``) when ``bounds`` are given."""

    KIND = _lib.ISO_PRIOR_GAUSSIAN
    SUPPORT = None

    def __init__(self, mean, sigma, bounds=None):
        BoundedPrior.__init__(self, bounds=bounds)
        self.mean, self.sigma = mean, sigma
        if bounds:
            self.norm = float(ndtr((bounds[1] - mean) / sigma) - ndtr((bounds[0] - mean) / sigma))
        else:
            self.norm = 1.0
        self.lognorm = float(np.log(self.norm))

    def _device_params(self):
        return (self.mean, self.sigma, self.norm, self.lognorm)

    def _pdf(self, x):
        return _gauss(x, self.mean, self.sigma) / self.norm

    def _sampling_window(self):
        return self._bounds if self._bounds is not None else (self.mean - 9 * self.sigma, self.mean + 9 * self.sigma)


class LogNormalPrior(Prior):
    """ln(x) ~ N(mu, sigma); support (0, inf)."""

    KIND = _lib.ISO_PRIOR_LOGNORMAL

    def __init__(self, mu, sigma, bounds=None):
        Prior.__init__(self)
        self.mu, self.sigma = mu, sigma
        self.scale = np.exp(mu)
        self.log_s = np.log(sigma)
        self._bounds = (0, np.inf)

    def _device_params(self):
        return (self.mu, self.sigma, self.scale, self.log_s)

    def _pdf(self, x):
        y = x / self.scale
        return _gauss(np.log(y), 0.0, self.sigma) / (y * self.scale)

    def _sampling_window(self):
        return (float(np.exp(self.mu - 9 * self.sigma)), float(np.exp(self.mu + 9 * self.sigma)))

    def sample(self, n):
        return np.exp(self.mu + self.sigma * np.random.standard_normal(n))


class FehPrior(Prior):
    """Metallicity distribution of the solar neighbourhood: a two-Gaussian disk (Casagrande et al. 2011 fit, as used by
    the reference, priors.py:345-381) or a single-Gaussian disk (``local=False``), plus a halo component."""

    KIND = _lib.ISO_PRIOR_FEH
    SUPPORT = (-4.5, 1.5)
    _DISK_LOCAL = ((0.8, 0.016, 0.15), (0.2, -0.15, 0.22))     # (weight, mean, sigma)
    _DISK_WIDE = ((1.0, -0.3, 0.3),)
    _HALO = (-1.5, 0.4)

    def __init__(self, halo_fraction=0.001, local=True, **kwargs):
        Prior.__init__(self, **kwargs)
        self.halo_fraction = halo_fraction
        self.local = local

    def _device_params(self):
        return (self.halo_fraction,)

    def _device_flags(self):
        return _lib.ISO_PF_LOCAL if self.local else 0

    def _pdf(self, x):
        disk = sum(w * _gauss(x, mu, sig) for w, mu, sig in (self._DISK_LOCAL if self.local else self._DISK_WIDE))
        if self.local:
            disk *= _SQRT_2PI / 2.5066282746310007      # the reference's rounded normalisation of the two-Gaussian fit
        return self.halo_fraction * _gauss(x, *self._HALO) + (1.0 - self.halo_fraction) * disk


# ---------------------------------------------------------------------------------------------------------------------
# piecewise prior
# ---------------------------------------------------------------------------------------------------------------------
class BrokenPrior(Prior):
    """Pieces ``components[i]`` on the intervals cut by ``breakpoints``, scaled to be continuous at the breaks and to
    integrate to one over ``bounds`` (``norms`` / ``lognorms``; the reference's priors.py:143-232)."""

    KIND = _lib.ISO_PRIOR_BROKEN

    def __init__(self, components, breakpoints, bounds=None):
        Prior.__init__(self)
        self.components = components
        self.n_components = len(components)
        self.breakpoints = breakpoints
        self.quad_args = dict(limit=200)
        self._bounds = (-np.inf, np.inf) if bounds is None else bounds
        self._initialize()

    def _set_bounds(self, new):
        self._bounds = new
        self._initialize()

    bounds = property(lambda self: self._get_bounds(), lambda self, new: self._set_bounds(new))

    def _initialize(self):
        edges = [self._get_bounds()[0], *self.breakpoints, self._get_bounds()[1]]
        self.domains = list(zip(edges[:-1], edges[1:]))
        # continuity: piece i is divided by the ratio of the two pieces at the break between them
        ratio = np.ones(self.n_components)
        for i, x in enumerate(self.breakpoints, start=1):
            ratio[i] = self.components[i].density(x) / self.components[i - 1].density(x)
        mass = np.array([quad(lambda x, c=c, r=r: c.density(x) / r, a, b, **self.quad_args)[0]
                         for c, r, (a, b) in zip(self.components, ratio, self.domains)])
        self.norms = ratio * mass.sum()
        self.lognorms = np.log(self.norms)
        self.cumnorm = mass / mass.sum()

    def _piece(self, x):
        return int(np.digitize(x, self.breakpoints))

    def _pdf(self, x):
        i = self._piece(x)
        return self.components[i].density(x) / self.norms[i]

    def to_struct(self):
        if self.n_components > _lib.ISO_MAX_COMP:
            raise TypeError("BrokenPrior with more than %d components is not supported on the device" % _lib.ISO_MAX_COMP)
        out = _lib.IsoPrior()
        out.self = self._leaf_struct()
        out.n_comp = self.n_components
        for i, b in enumerate(self.breakpoints):
            out.breakpoints[i] = float(b)
        for i, comp in enumerate(self.components):
            if comp.KIND is None or comp.KIND == _lib.ISO_PRIOR_BROKEN:
                raise TypeError("BrokenPrior component %s has no device implementation" % type(comp).__name__)
            out.norms[i], out.lognorms[i] = float(self.norms[i]), float(self.lognorms[i])
            out.comp[i] = comp._leaf_struct()
        return out

    def _sampling_window(self):
        lo, hi = self._get_bounds()
        if np.isfinite(lo) and np.isfinite(hi):
            return lo, hi
        wins = [c._sampling_window() for c in self.components]
        return (lo if np.isfinite(lo) else min(w[0] for w in wins)), (hi if np.isfinite(hi) else max(w[1] for w in wins))


# ---------------------------------------------------------------------------------------------------------------------
# EEP prior: evaluated inside the fused kernel; the stand-alone object is for users and for sampling
# ---------------------------------------------------------------------------------------------------------------------
class EEP_prior(BoundedPrior):
    """Prior on EEP induced by the prior on the parameter EEP replaces: ``pdf(eep) = orig_prior(value) * d value / d EEP``
    with ``value`` (age on track grids, mass on isochrone grids) and its EEP-derivative interpolated from the model grid
    (the reference's priors.py:409-429)."""

    _DERIV = {"age": "dt_deep", "mass": "dm_deep"}

    def __init__(self, ic, orig_prior, bounds=None):
        BoundedPrior.__init__(self, bounds=ic.eep_bounds if bounds is None else bounds)
        self.ic = ic
        self.orig_prior = orig_prior
        self.orig_par = ic.eep_replaces
        if self.orig_par not in self._DERIV:
            raise ValueError("wtf.")
        self.deriv_prop = self._DERIV[self.orig_par]

    def to_struct(self):
        raise TypeError("EEP_prior is evaluated inside the fused lnpost kernel, not as a stand-alone prior struct")

    def _grid_coordinates(self, eep, others):
        return [others["mass"], eep, others["feh"]] if self.orig_par == "age" else [eep, others["age"], others["feh"]]

    def pdf(self, x, **kwargs):
        """Both factors come from the GPU (grid interpolation, original prior); scalars or equal-length arrays."""
        scalar = all(_scalar_like(v) for v in (x, *kwargs.values()))
        coords = [np.atleast_1d(np.asarray(v, dtype=float)) for v in self._grid_coordinates(x, kwargs)]
        vals = np.atleast_2d(self.ic.interp_value(coords, [self.orig_par, self.deriv_prop]))
        with np.errstate(invalid="ignore"):
            dens = np.atleast_1d(self.orig_prior(vals[:, 0])) * vals[:, 1] / self._norm
            if self._bounds is not None:
                e = np.broadcast_to(np.atleast_1d(np.asarray(x, dtype=float)), dens.shape)
                dens = np.where((e < self._bounds[0]) | (e > self._bounds[1]), 0.0, dens)
        return float(dens[0]) if scalar else dens

    __call__ = pdf

    def lnpdf(self, x, **kwargs):
        dens = np.asarray(self.pdf(x, **kwargs), dtype=float)
        with np.errstate(divide="ignore", invalid="ignore"):
            out = np.where(dens == 0, -np.inf, np.log(np.where(dens == 0, 1.0, dens)))
        return float(out) if out.ndim == 0 else out

    def sample(self, n, **kwargs):
        """Integer EEPs drawn with weights pdf(eep | other parameters)."""
        candidates = np.random.choice(np.arange(self._bounds[0], self._bounds[1]), size=n).astype(float)
        others = {k: np.resize(np.asarray(v, dtype=float), n) for k, v in kwargs.items()}
        w = np.clip(np.nan_to_num(np.asarray(self.pdf(candidates, **others)), nan=0.0, posinf=0.0, neginf=0.0), 0, None)
        if w.sum() <= 0:
            return self.sample(n, **kwargs)
        return np.random.choice(candidates, size=n, p=w / w.sum())

    def test_integral(self):
        pass


# ---------------------------------------------------------------------------------------------------------------------
# the named priors of the star models
# ---------------------------------------------------------------------------------------------------------------------
class AgePrior(FlatLogPrior):
    """Uniform in linear age; the parameter is log10(age / yr)."""

    def __init__(self, **kwargs):
        FlatLogPrior.__init__(self, bounds=(5, 10.15), **kwargs)


class DistancePrior(PowerLawPrior):
    """Constant space density: p(d) ∝ d²."""

    def __init__(self, max_distance=10000, **kwargs):
        PowerLawPrior.__init__(self, alpha=2.0, bounds=(0, max_distance), **kwargs)


class AVPrior(FlatPrior):
    def __init__(self, **kwargs):
        FlatPrior.__init__(self, bounds=kwargs.pop("bounds", (0, 1.0)))


class QPrior(PowerLawPrior):
    def __init__(self, **kwargs):
        PowerLawPrior.__init__(self, alpha=0.3, bounds=kwargs.pop("bounds", (0.1, 1)), **kwargs)


class SalpeterPrior(PowerLawPrior):
    def __init__(self, **kwargs):
        PowerLawPrior.__init__(self, alpha=-2.35, bounds=kwargs.pop("bounds", (0.1, 10)), **kwargs)


class ChabrierPrior(BrokenPrior):
    """Chabrier (2003, eq. 17) system IMF: log-normal below 1 Msun, Salpeter slope above."""

    def __init__(self, **kwargs):
        pieces = [LogNormalPrior(np.log(0.079), 0.69 * np.log(10)), PowerLawPrior(-2.35, (1.0, 100.0))]
        BrokenPrior.__init__(self, pieces, [1.0], bounds=kwargs.pop("bounds", (0.1, 100.0)), **kwargs)
