"""Prior classes — drop-in for the evaluation path of the reference's ``isochrones/priors.py``.

Same class names, constructor signatures and attributes (``bounds``, ``_bounds``, ``_norm``, ``alpha``,
``mean``, ``sigma``, ``norm``, ``lognorm``, ``mu``, ``scale``, ``log_s``, ``halo_fraction``, ``local``,
``components``, ``breakpoints``, ``norms``, ``lognorms``, ``orig_prior``...).

* ``prior(x)``, ``prior.pdf(x)`` and ``prior.lnpdf(x)`` are evaluated on the GPU (``iso_prior_eval``; scalars or
  arrays) with the reference's exact rules (priors.py:35-66, 112-140, 205-211) — there is no CPU evaluation
  path behind them.
* ``_pdf`` is the host-side formula the reference also keeps on the host: it is used only at construction
  time by ``scipy.integrate.quad`` for the normalisation constants (``_norm``, ``norms``, ``lognorms``,
  priors.py:42-49, 176-203) that are then shipped to the device inside the prior struct.
* A prior that is not one of the classes below cannot be compiled for the device and raises ``TypeError``
  when a model is built (no CPU fallback).
"""
import ctypes as C

import numpy as np
from scipy.integrate import quad
from scipy.special import ndtr as _norm_cdf

from . import _lib

_norm_pdf_C = np.sqrt(2 * np.pi)
ONE_OVER_ROOT_2PI = 1.0 / _norm_pdf_C
_norm_pdf_logC = np.log(_norm_pdf_C)
LOG_ONE_OVER_ROOT_2PI = np.log(ONE_OVER_ROOT_2PI)


def _is_scalar(x):
    return np.ndim(x) == 0


class Prior(object):
    """Base class (priors.py:31-104): unbounded unless ``_bounds`` is set; ``lnpdf = log(pdf)`` or ``_lnpdf``."""

    _kind = None          # ISO_PRIOR_* code of the concrete class
    _bounded = False      # subclass of BoundedPrior
    _has_lnpdf = False

    def __init__(self, *args, **kwargs):
        self._norm = 1.0

    # ---- device evaluation -------------------------------------------------------------------------------
    def _leaf_struct(self):
        s = _lib.IsoPriorLeaf()
        s.kind = self._kind
        raw = getattr(self, "_bounds", None)
        s.flags = (_lib.ISO_PF_BOUNDED if self._bounded else 0) | (_lib.ISO_PF_HAS_BOUNDS if raw is not None else 0)
        if raw is not None:
            s.lo, s.hi = float(raw[0]), float(raw[1])
        else:
            s.lo, s.hi = -np.inf, np.inf
        s.norm = float(getattr(self, "_norm", 1.0))
        self._fill_params(s)
        return s

    def _fill_params(self, s):
        pass

    def to_struct(self):
        """``iso_prior`` image of this object (raises ``TypeError`` for classes the device cannot evaluate)."""
        if self._kind is None:
            raise TypeError("prior class %s has no device implementation (isochrones_b200 has no CPU fallback)"
                            % type(self).__name__)
        s = _lib.IsoPrior()
        s.self = self._leaf_struct()
        return s

    def _eval(self, x, which, ctx=None):
        ctx = ctx or _lib.default_context()
        xs = _lib.f64(np.atleast_1d(x)).ravel()
        out = np.empty_like(xs)
        s = self.to_struct()
        ctx.check(_lib.lib().iso_prior_eval(ctx.handle, C.byref(s), which, _lib.dp(xs), len(xs), _lib.dp(out)))
        if _is_scalar(x):
            return float(out[0])
        return out.reshape(np.shape(x))

    def __call__(self, x, **kwargs):
        return self._eval(x, 1)

    def pdf(self, x, **kwargs):
        """``Prior.pdf`` (priors.py:54-59).  For the classes here it differs from ``__call__`` only through
        BoundedPrior's extra (identical) bounds test, so both share the device entry point."""
        return self._eval(x, 1)

    def lnpdf(self, x, **kwargs):
        return self._eval(x, 0)

    # ---- host-side construction helpers --------------------------------------------------------------------
    @property
    def bounds(self):
        return (-np.inf, np.inf) if getattr(self, "_bounds", None) is None else self._bounds

    @bounds.setter
    def bounds(self, new):
        self._norm = quad(self._pdf, *new)[0]          # priors.py:42-44
        self._bounds = new
        try:
            self.test_integral()
        except AssertionError:
            raise ValueError(f"Problem setting bounds to {new}; integral test failed.")

    def _pdf(self, x, **kwargs):
        raise NotImplementedError

    def _host_call(self, x):
        """Host mirror of ``__call__`` used only by construction-time quadrature."""
        lo, hi = self.bounds
        if x < lo or x > hi:
            return 0
        return self._pdf(x) / self._norm

    def test_integral(self):
        assert np.isclose(1, quad(self._host_call, *self.bounds)[0])

    def sample(self, n):
        if hasattr(self, "distribution"):
            return self.distribution.rvs(n)
        raise NotImplementedError

    def test_sampling(self, n=100000, plot=False):
        """Histogram of ``sample(n)`` against bin-averaged pdf, within 6 sigma (priors.py:77-104; host-side check)."""
        x = self.sample(n)
        rng = None if tuple(self.bounds) == (-np.inf, np.inf) else self.bounds
        hn, _ = np.histogram(x, range=rng)
        h, b = np.histogram(x, density=True, range=rng)
        pdf = np.array([quad(self._host_call, lo, hi)[0] / (hi - lo) for lo, hi in zip(b[:-1], b[1:])])
        with np.errstate(divide="ignore", invalid="ignore"):
            sigma = 1.0 / np.sqrt(hn)
            resid = np.absolute(pdf - h) / pdf
            assert max((resid / sigma)[hn > 50]) < 6


class BoundedPrior(Prior):
    """priors.py:107-140: ``-inf`` / 0 outside ``bounds``."""

    _bounded = True

    def __init__(self, bounds=None):
        self._bounds = bounds
        super(BoundedPrior, self).__init__()

    @property
    def bounds(self):
        return self._bounds

    @bounds.setter
    def bounds(self, new):
        self._bounds = new
        try:
            self.test_integral()
        except AssertionError:
            raise ValueError(f"Problem setting bounds to {new}; integral test failed.")

    def _host_call(self, x):
        if self.bounds is not None:
            lo, hi = self.bounds
            if x < lo or x > hi:
                return 0
        return self._pdf(x) / self._norm


class BrokenPrior(Prior):
    """Composition of stitched-together priors with breakpoints (priors.py:143-232)."""

    _kind = _lib.ISO_PRIOR_BROKEN
    _has_lnpdf = True

    def __init__(self, components, breakpoints, bounds=None):
        self.components = components
        self.n_components = len(components)
        self.breakpoints = breakpoints
        if bounds is None:
            bounds = (-np.inf, np.inf)
        self._bounds = bounds
        self._norm = 1.0
        self.quad_args = dict(limit=200)
        self._initialize()

    @property
    def bounds(self):
        return (-np.inf, np.inf) if getattr(self, "_bounds", None) is None else self._bounds

    @bounds.setter
    def bounds(self, new):
        self._bounds = new
        self._initialize()

    def _initialize(self):
        # priors.py:176-203 (host-side normalisation of the pieces)
        lo, hi = self.bounds
        full_domain = [lo] + list(self.breakpoints) + [hi]
        self.domains = [(a, b) for a, b in zip(full_domain[:-1], full_domain[1:])]
        norms = np.ones(self.n_components)
        for i in range(1, self.n_components):
            x = self.breakpoints[i - 1]
            norms[i] = self.components[i]._host_call(x) / self.components[i - 1]._host_call(x)
        tot = 0
        for comp, (a, b), norm in zip(self.components, self.domains, norms):
            tot += quad(lambda x: comp._host_call(x) / norm, a, b, **self.quad_args)[0]
        self.norms = norms * tot
        self.lognorms = np.log(self.norms)
        cumnorm = np.zeros(self.n_components)
        for i, (comp, (a, b), norm) in enumerate(zip(self.components, self.domains, self.norms)):
            cumnorm[i] = quad(lambda x: comp._host_call(x) / norm, a, b, **self.quad_args)[0]
        self.cumnorm = cumnorm

    def _pdf(self, x):
        i = np.digitize(x, self.breakpoints)
        return self.components[i]._host_call(x) / self.norms[i]

    def to_struct(self):
        if self.n_components > _lib.ISO_MAX_COMP:
            raise TypeError("BrokenPrior with more than %d components is not supported on the device" % _lib.ISO_MAX_COMP)
        s = _lib.IsoPrior()
        s.self = self._leaf_struct()
        s.n_comp = self.n_components
        for i, b in enumerate(self.breakpoints):
            s.breakpoints[i] = float(b)
        for i, comp in enumerate(self.components):
            if comp._kind is None or comp._kind == _lib.ISO_PRIOR_BROKEN:
                raise TypeError("BrokenPrior component %s has no device implementation" % type(comp).__name__)
            s.norms[i] = float(self.norms[i])
            s.lognorms[i] = float(self.lognorms[i])
            s.comp[i] = comp._leaf_struct()
        return s

    def sample(self, n):
        u = np.random.random(n)
        x = np.zeros(n)
        u_cumthresh = 0
        for comp, u_thresh, (a, b) in zip(self.components, self.cumnorm, self.domains):
            u_cumthresh += u_thresh
            mask = (u < u_cumthresh) & (x == 0.0)
            n_comp = mask.sum()
            samples = comp.sample(n_comp)
            oob = (samples < a) | (samples > b)
            while oob.sum():
                samples[oob] = comp.sample(oob.sum())
                oob = (samples < a) | (samples > b)
            x[mask] = samples
        return x


class GaussianPrior(BoundedPrior):
    _kind = _lib.ISO_PRIOR_GAUSSIAN
    _has_lnpdf = True

    def __init__(self, mean, sigma, bounds=None):
        import scipy.stats

        self.mean = mean
        self.sigma = sigma
        self._bounds = bounds
        self._norm = 1.0
        if bounds:
            lo, hi = bounds
            a, b = (lo - mean) / sigma, (hi - mean) / sigma
            self.distribution = scipy.stats.truncnorm(a, b, loc=mean, scale=sigma)
            self.norm = _norm_cdf(b) - _norm_cdf(a)
            self.lognorm = np.log(self.norm)
        else:
            self.distribution = scipy.stats.norm(mean, sigma)
            self.norm = 1.0
            self.lognorm = 0.0

    def _fill_params(self, s):
        s.a[0], s.a[1], s.a[2], s.a[3] = float(self.mean), float(self.sigma), float(self.norm), float(self.lognorm)

    def _pdf(self, x):
        z = (x - self.mean) / self.sigma
        return np.exp(-(z ** 2) / 2.0) / _norm_pdf_C / self.sigma / self.norm


class LogNormalPrior(Prior):
    _kind = _lib.ISO_PRIOR_LOGNORMAL
    _has_lnpdf = True

    def __init__(self, mu, sigma, bounds=None):
        from scipy.stats import lognorm

        self.mu = mu
        self.sigma = sigma
        self.scale = np.exp(mu)
        self.log_s = np.log(sigma)
        self.distribution = lognorm(sigma, scale=np.exp(mu))
        self._bounds = (0, np.inf)
        super().__init__(self)

    def _fill_params(self, s):
        s.a[0], s.a[1], s.a[2], s.a[3] = float(self.mu), float(self.sigma), float(self.scale), float(self.log_s)

    def _pdf(self, x):
        s = self.sigma
        y = x / self.scale
        return ONE_OVER_ROOT_2PI / (s * y) * np.exp(-0.5 * (np.log(y) / s) ** 2) / self.scale


class FlatPrior(BoundedPrior):
    _kind = _lib.ISO_PRIOR_FLAT

    def __init__(self, bounds):
        super().__init__(bounds=bounds)

    def _pdf(self, x):
        lo, hi = self.bounds
        return 1.0 / (hi - lo)

    def sample(self, n):
        lo, hi = self.bounds
        return np.random.random(n) * (hi - lo) + lo


class FlatLogPrior(BoundedPrior):
    _kind = _lib.ISO_PRIOR_FLATLOG

    def __init__(self, bounds):
        super(FlatLogPrior, self).__init__(bounds=bounds)

    def _pdf(self, x):
        lo, hi = self.bounds
        return np.log(10) * 10 ** x / (10 ** hi - 10 ** lo)

    def sample(self, n):
        lo, hi = self.bounds
        return np.log10(np.random.random(n) * (10 ** hi - 10 ** lo) + 10 ** lo)


class PowerLawPrior(BoundedPrior):
    _kind = _lib.ISO_PRIOR_POWERLAW
    _has_lnpdf = True

    def __init__(self, alpha, bounds=None):
        self.alpha = alpha
        super(PowerLawPrior, self).__init__(bounds=bounds)

    def _fill_params(self, s):
        s.a[0] = float(self.alpha)

    def _pdf(self, x):
        lo, hi = [np.float64(b) for b in self.bounds]
        with np.errstate(divide="ignore"):
            C_ = (1 + self.alpha) / (hi ** (1 + self.alpha) - lo ** (1 + self.alpha))
        return C_ * np.float64(x) ** self.alpha

    def sample(self, n):
        lo, hi = self.bounds
        C_ = (1 + self.alpha) / (hi ** (1 + self.alpha) - lo ** (1 + self.alpha))
        u = np.random.random(n)
        a = self.alpha
        return ((a + 1) * (u / C_ + (lo ** (a + 1) / (a + 1)))) ** (1 / (a + 1))


class FehPrior(Prior):
    """feh PDF based on the local SDSS distribution (priors.py:345-406)."""

    _kind = _lib.ISO_PRIOR_FEH

    def __init__(self, halo_fraction=0.001, local=True, **kwargs):
        self.halo_fraction = halo_fraction
        self.local = local
        super().__init__(**kwargs)

    def _fill_params(self, s):
        s.a[0] = float(self.halo_fraction)
        if self.local:
            s.flags |= _lib.ISO_PF_LOCAL

    def _pdf(self, x):
        feh = x
        if self.local:
            disk_norm = 2.5066282746310007
            disk_fehdist = (1.0 / disk_norm * (0.8 / 0.15 * np.exp(-0.5 * (feh - 0.016) ** 2.0 / 0.15 ** 2.0)
                                               + 0.2 / 0.22 * np.exp(-0.5 * (feh + 0.15) ** 2.0 / 0.22 ** 2.0)))
        else:
            mu, sig = -0.3, 0.3
            disk_fehdist = 1.0 / np.sqrt(2 * np.pi) / sig * np.exp(-0.5 * (feh - mu) ** 2 / sig ** 2)
        halo_mu, halo_sig = -1.5, 0.4
        halo_fehdist = 1.0 / np.sqrt(2 * np.pi * halo_sig ** 2) * np.exp(-0.5 * (feh - halo_mu) ** 2 / halo_sig ** 2)
        return self.halo_fraction * halo_fehdist + (1 - self.halo_fraction) * disk_fehdist

    def sample(self, n):
        if self.local:
            w2, mu1, sig1, mu2, sig2 = 0.2, 0.016, 0.15, -0.15, 0.22
        else:
            w2, mu1, sig1, mu2, sig2 = 0.0, -0.3, 0.3, 0, 1
        x = np.random.randn(n) * sig1 + mu1
        x2 = np.random.randn(n) * sig2 + mu2
        xhalo = np.random.randn(n) * 0.4 - 1.5
        m1 = np.random.random(n) < w2
        x[m1] = x2[m1]
        m2 = np.random.random(n) < self.halo_fraction
        x[m2] = xhalo[m2]
        return x


class EEP_prior(BoundedPrior):
    """Prior on EEP induced by the prior of the parameter it replaces (priors.py:409-465):
    ``pdf(eep) = orig_prior(orig_val) * d(orig)/d(EEP)`` with both factors interpolated from the model grid."""

    def __init__(self, ic, orig_prior, bounds=None):
        self.ic = ic
        self.orig_prior = orig_prior
        self._bounds = bounds if bounds is not None else ic.eep_bounds
        self._norm = 1.0
        self.orig_par = ic.eep_replaces
        if self.orig_par == "age":
            self.deriv_prop = "dt_deep"
        elif self.orig_par == "mass":
            self.deriv_prop = "dm_deep"
        else:
            raise ValueError("wtf.")

    def to_struct(self):
        raise TypeError("EEP_prior is evaluated inside the fused lnpost kernel, not as a stand-alone prior struct")

    def _pars(self, eep, kwargs):
        if self.orig_par == "age":
            return [kwargs["mass"], eep, kwargs["feh"]]
        return [eep, kwargs["age"], kwargs["feh"]]

    def pdf(self, x, **kwargs):
        """Prior.pdf (priors.py:54-59) over EEP_prior._pdf (:423-429); grid interpolation and the original
        prior both run on the GPU.  Scalars or equal-length arrays."""
        eep = x
        scalar = all(_is_scalar(v) for v in [eep] + list(kwargs.values()))
        pars = [np.atleast_1d(np.asarray(v, dtype=float)) for v in self._pars(eep, kwargs)]
        vals = np.atleast_2d(self.ic.interp_value(pars, [self.orig_par, self.deriv_prop]))
        orig = np.atleast_1d(self.orig_prior(vals[:, 0]))
        with np.errstate(invalid="ignore"):
            pdf = orig * vals[:, 1] / self._norm
            e = np.broadcast_to(np.atleast_1d(np.asarray(eep, dtype=float)), pdf.shape)
            if self._bounds is not None:
                lo, hi = self._bounds
                pdf = np.where((e < lo) | (e > hi), 0.0, pdf)
        return float(pdf[0]) if scalar else pdf

    def __call__(self, x, **kwargs):
        return self.pdf(x, **kwargs)

    def lnpdf(self, x, **kwargs):
        pdf = self.pdf(x, **kwargs)
        with np.errstate(divide="ignore", invalid="ignore"):
            out = np.where(np.asarray(pdf) == 0, -np.inf, np.log(np.where(np.asarray(pdf) == 0, 1.0, pdf)))
        return float(out) if _is_scalar(pdf) else out

    def sample(self, n, **kwargs):
        rng = np.random
        eeps = rng.choice(np.arange(self.bounds[0], self.bounds[1]), size=n, replace=True).astype(float)
        w = np.nan_to_num(np.asarray(self.pdf(eeps, **{k: np.resize(np.asarray(v, dtype=float), n) for k, v in kwargs.items()})),
                          nan=0.0, posinf=0.0, neginf=0.0)
        w = np.clip(w, 0, None)
        if w.sum() <= 0:
            return self.sample(n, **kwargs)
        return rng.choice(eeps, size=n, replace=True, p=w / w.sum())

    def test_integral(self):
        pass


class AgePrior(FlatLogPrior):
    """Uniform true age prior, where 'age' is actually log10(age) (priors.py:483-488)."""

    def __init__(self, **kwargs):
        super().__init__(bounds=(5, 10.15), **kwargs)


class DistancePrior(PowerLawPrior):
    def __init__(self, max_distance=10000, **kwargs):
        super().__init__(alpha=2.0, bounds=(0, max_distance), **kwargs)


class AVPrior(FlatPrior):
    def __init__(self, **kwargs):
        bounds = kwargs.pop("bounds", (0, 1.0))
        super().__init__(bounds=bounds)


class QPrior(PowerLawPrior):
    def __init__(self, **kwargs):
        bounds = kwargs.pop("bounds", (0.1, 1))
        super().__init__(alpha=0.3, bounds=bounds, **kwargs)


class SalpeterPrior(PowerLawPrior):
    def __init__(self, **kwargs):
        bounds = kwargs.pop("bounds", (0.1, 10))
        super().__init__(alpha=-2.35, bounds=bounds, **kwargs)


class ChabrierPrior(BrokenPrior):
    def __init__(self, **kwargs):
        bounds = kwargs.pop("bounds", (0.1, 100.0))
        super().__init__(
            [LogNormalPrior(np.log(0.079), 0.69 * np.log(10)), PowerLawPrior(-2.35, (1.0, 100.0))], [1.0],
            bounds=bounds, **kwargs
        )  # Chabrier 2003, eqn 17
