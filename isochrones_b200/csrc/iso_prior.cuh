// isochrones_b200 — device evaluation of the prior classes of priors.py.
//
// Class-by-class restatement of the reference's evaluation rules (NOT of its class machinery):
//   Prior.__call__ / pdf / lnpdf        priors.py:35-36, 54-66
//   BoundedPrior.__call__ / lnpdf       priors.py:112-117, 131-140
//   BrokenPrior._pdf / _lnpdf           priors.py:205-211   (ChabrierPrior :514-519)
//   GaussianPrior :235-257, LogNormalPrior :260-280, FlatPrior :283-293, FlatLogPrior :296-306,
//   PowerLawPrior :309-342 (powerlaw_pdf / powerlaw_lnpdf :469-480), FehPrior :345-381.
// Quirks that decide finite / -inf / NaN outcomes are kept: Prior.lnpdf of a class with _lnpdf does no bounds
// test (Chabrier above its upper bound is only cut by its power-law component at 100), `log(pdf) if pdf else
// -inf` maps 0 to -inf but negative / NaN pdf to NaN, np.digitize sends NaN to the last component.
#pragma once

#include <math.h>

#include "iso_common.cuh"

// ---- constants the library derives from the public struct at staging time (iso_prior_leaf.k) -----------------
static inline void iso_prior_leaf_fill(iso_prior_leaf *p)
{
    p->k[0] = p->k[1] = 0.0;
    switch (p->kind) {
    case ISO_PRIOR_FLAT: {  // priors.py:287-289; lnpdf is the constant log(pdf / _norm) (or -inf when that is 0)
        p->k[0] = 1.0 / (p->hi - p->lo);
        double pdf = p->k[0] / p->norm;
        p->k[1] = pdf == 0.0 ? -INFINITY : log(pdf);
        break;
    }
    case ISO_PRIOR_FLATLOG:  // priors.py:300-302
        p->k[0] = log(10.0);
        p->k[1] = pow(10.0, p->hi) - pow(10.0, p->lo);
        break;
    case ISO_PRIOR_POWERLAW: {  // priors.py:469-480
        double alpha = p->a[0];
        p->k[0] = (1 + alpha) / (pow(p->hi, 1 + alpha) - pow(p->lo, 1 + alpha));
        p->k[1] = log(p->k[0]);
        break;
    }
    case ISO_PRIOR_GAUSSIAN:  // priors.py:22-27
        p->k[0] = log(sqrt(2 * M_PI));
        p->k[1] = log(p->a[1]);
        break;
    case ISO_PRIOR_LOGNORMAL:  // priors.py:272-280
        p->k[0] = log(1.0 / sqrt(2 * M_PI));
        p->k[1] = 1.0 / sqrt(2 * M_PI);
        break;
    default:
        break;
    }
}

static inline void iso_prior_fill(iso_prior *p)
{
    iso_prior_leaf_fill(&p->self);
    for (int i = 0; i < ISO_MAX_COMP; i++) iso_prior_leaf_fill(&p->comp[i]);
}

static inline bool iso_prior_leaf_valid(const iso_prior_leaf &p)
{
    return p.kind >= ISO_PRIOR_FLAT && p.kind <= ISO_PRIOR_FEH;
}

static inline bool iso_prior_valid(const iso_prior &p)
{
    if (p.self.kind == ISO_PRIOR_BROKEN) {
        if (p.n_comp < 1 || p.n_comp > ISO_MAX_COMP) return false;
        for (int i = 0; i < p.n_comp; i++)
            if (!iso_prior_leaf_valid(p.comp[i])) return false;
        return true;
    }
    return iso_prior_leaf_valid(p.self);
}

#ifdef __CUDACC__

__device__ __forceinline__ bool iso_outside(double x, double lo, double hi) { return (x < lo) || (x > hi); }

// `np.log(pdf) if pdf else -np.inf`  (priors.py:66, 140)
__device__ __forceinline__ double iso_log_or_neginf(double pdf) { return pdf == 0.0 ? iso_neg_inf() : log(pdf); }

// FehPrior._pdf  priors.py:359-381
__device__ __forceinline__ double iso_feh_pdf(const iso_prior_leaf &p, double feh)
{
    const double halo_fraction = p.a[0];
    double disk;
    if (p.flags & ISO_PF_LOCAL) {
        const double disk_norm = 2.5066282746310007;
        double d1 = feh - 0.016, d2 = feh + 0.15;
        disk = 1.0 / disk_norm * (0.8 / 0.15 * exp(-0.5 * (d1 * d1) / (0.15 * 0.15)) +
                                  0.2 / 0.22 * exp(-0.5 * (d2 * d2) / (0.22 * 0.22)));
    } else {
        double d = feh - (-0.3);
        disk = 1.0 / 2.5066282746310002 / 0.3 * exp(-0.5 * (d * d) / (0.3 * 0.3));
    }
    double dh = feh - (-1.5);
    double halo = 1.0 / 1.0026513098524001 * exp(-0.5 * (dh * dh) / (0.4 * 0.4));   // 1 / sqrt(2 pi 0.4^2)
    return halo_fraction * halo + (1 - halo_fraction) * disk;
}

// self._pdf(x) of a non-broken class
__device__ __forceinline__ double iso_leaf_pdf_raw(const iso_prior_leaf &p, double x)
{
    switch (p.kind) {
    case ISO_PRIOR_FLAT:
        return p.k[0];
    case ISO_PRIOR_FLATLOG:
        return p.k[0] * exp10(x) / p.k[1];
    case ISO_PRIOR_POWERLAW:
        return p.k[0] * pow(x, p.a[0]);
    case ISO_PRIOR_GAUSSIAN: {
        double z = (x - p.a[0]) / p.a[1];
        return exp(-(z * z) / 2.0) / 2.5066282746310002 / p.a[1] / p.a[2];
    }
    case ISO_PRIOR_LOGNORMAL: {
        double s = p.a[1], yv = x / p.a[2], t = log(yv) / s;
        return p.k[1] / (s * yv) * exp(-0.5 * (t * t)) / p.a[2];
    }
    case ISO_PRIOR_FEH:
        return iso_feh_pdf(p, x);
    default:
        return iso_nan();
    }
}

// Prior.pdf  priors.py:54-59
__device__ __forceinline__ double iso_leaf_pdf(const iso_prior_leaf &p, double x)
{
    if ((p.flags & ISO_PF_HAS_BOUNDS) && iso_outside(x, p.lo, p.hi)) return 0.0;
    return iso_leaf_pdf_raw(p, x) / p.norm;
}

// Prior.__call__ priors.py:35-36 / BoundedPrior.__call__ :112-117
__device__ __forceinline__ double iso_leaf_call(const iso_prior_leaf &p, double x)
{
    if ((p.flags & ISO_PF_BOUNDED) && (p.flags & ISO_PF_HAS_BOUNDS) && iso_outside(x, p.lo, p.hi)) return 0.0;
    return iso_leaf_pdf(p, x);
}

__device__ __forceinline__ bool iso_kind_has_lnpdf(int kind)
{
    return kind == ISO_PRIOR_POWERLAW || kind == ISO_PRIOR_GAUSSIAN || kind == ISO_PRIOR_LOGNORMAL ||
           kind == ISO_PRIOR_BROKEN;
}

// self._lnpdf(x) of a non-broken class that has one
__device__ __forceinline__ double iso_leaf_lnpdf_raw(const iso_prior_leaf &p, double x)
{
    switch (p.kind) {
    case ISO_PRIOR_POWERLAW:   // priors.py:476-480
        return p.k[1] + p.a[0] * log(x);
    case ISO_PRIOR_GAUSSIAN: { // priors.py:256-257, 26-27
        double z = (x - p.a[0]) / p.a[1];
        return (-(z * z) / 2.0 - p.k[0]) - p.k[1] - p.a[3];
    }
    case ISO_PRIOR_LOGNORMAL: { // priors.py:277-280
        double s = p.a[1], ly = log(x / p.a[2]), t = ly / s;
        return p.k[0] - (p.a[3] + ly) - 0.5 * (t * t) - p.a[0];
    }
    default:
        return iso_nan();
    }
}

// Prior.lnpdf priors.py:61-66 / BoundedPrior.lnpdf :131-140 of a non-broken class
__device__ __forceinline__ double iso_leaf_lnpdf(const iso_prior_leaf &p, double x)
{
    if (p.flags & ISO_PF_BOUNDED) {
        if ((p.flags & ISO_PF_HAS_BOUNDS) && iso_outside(x, p.lo, p.hi)) return iso_neg_inf();
        if (iso_kind_has_lnpdf(p.kind)) return iso_leaf_lnpdf_raw(p, x);
        if (p.kind == ISO_PRIOR_FLAT) return p.k[1];   // x-independent: log evaluated once at staging time
        return iso_log_or_neginf(iso_leaf_pdf(p, x));
    }
    if (iso_kind_has_lnpdf(p.kind)) return iso_leaf_lnpdf_raw(p, x);
    return iso_log_or_neginf(iso_leaf_call(p, x));
}

// np.digitize(x, breakpoints): number of breakpoints <= x; NaN sorts after everything
__device__ __forceinline__ int iso_digitize(const iso_prior &p, double x)
{
    int n = p.n_comp - 1;
    if (x != x) return n;
    int i = 0;
    while (i < n && p.breakpoints[i] <= x) i++;
    return i;
}

// any prior object: __call__(x)
__device__ __forceinline__ double iso_prior_call(const iso_prior &p, double x)
{
    if (p.self.kind != ISO_PRIOR_BROKEN) return iso_leaf_call(p.self, x);
    // BrokenPrior is a plain Prior: __call__ = pdf (bounds test when set) of _pdf / _norm   priors.py:35-36, 54-59
    if ((p.self.flags & ISO_PF_HAS_BOUNDS) && iso_outside(x, p.self.lo, p.self.hi)) return 0.0;
    int i = iso_digitize(p, x);
    double raw = (i == 0 ? iso_leaf_call(p.comp[0], x) : i == 1 ? iso_leaf_call(p.comp[1], x) : iso_leaf_call(p.comp[2], x)) /
                 p.norms[i];   // priors.py:205-207
    return raw / p.self.norm;
}

// any prior object: lnpdf(x)
__device__ __forceinline__ double iso_prior_lnpdf(const iso_prior &p, double x)
{
    if (p.self.kind != ISO_PRIOR_BROKEN) return iso_leaf_lnpdf(p.self, x);
    int i = iso_digitize(p, x);   // priors.py:209-211 — no bounds test on this path
    double l = i == 0 ? iso_leaf_lnpdf(p.comp[0], x) : i == 1 ? iso_leaf_lnpdf(p.comp[1], x) : iso_leaf_lnpdf(p.comp[2], x);
    return l - p.lognorms[i];
}

#endif  // __CUDACC__
