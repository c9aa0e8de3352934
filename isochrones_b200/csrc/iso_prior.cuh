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
//
// Every formula exists once, in the KIND-templated leaf functions.  The fused kernel instantiates them with the
// compile-time kinds of the reference's default BasicStarModel priors (small, switch-free code); any other
// combination goes through the *_dyn dispatchers, which are deliberately NOT inlined so the kernel's code stays
// inside the instruction cache.  Divisions by per-prior constants are multiplications by reciprocals computed at
// staging time (<= 1 ulp from the reference's quotient).
#pragma once

#include <math.h>

#include "iso_common.cuh"

// ---- constants the library derives from the public struct at staging time (iso_prior_leaf.k) -----------------
static inline void iso_prior_leaf_fill(iso_prior_leaf *p)
{
    for (int i = 0; i < 4; i++) p->k[i] = 0.0;
    const double inv_norm = 1.0 / p->norm;
    switch (p->kind) {
    case ISO_PRIOR_FLAT: {  // priors.py:287-289; lnpdf is the constant log(pdf / _norm) (or -inf when that is 0)
        p->k[0] = 1.0 / (p->hi - p->lo);
        double pdf = p->k[0] / p->norm;
        p->k[1] = pdf == 0.0 ? -INFINITY : log(pdf);
        p->k[2] = pdf;
        break;
    }
    case ISO_PRIOR_FLATLOG:  // priors.py:300-302: pdf / _norm = ln10 10^x / (10^hi - 10^lo) / _norm = k2 10^x
        p->k[0] = log(10.0);
        p->k[1] = pow(10.0, p->hi) - pow(10.0, p->lo);
        p->k[2] = p->k[0] / p->k[1] / p->norm;
        p->k[3] = log(p->k[2]);
        break;
    case ISO_PRIOR_POWERLAW: {  // priors.py:469-480
        double alpha = p->a[0];
        p->k[0] = (1 + alpha) / (pow(p->hi, 1 + alpha) - pow(p->lo, 1 + alpha));
        p->k[1] = log(p->k[0]);
        p->k[2] = p->k[0] * inv_norm;
        break;
    }
    case ISO_PRIOR_GAUSSIAN:  // priors.py:22-27, 253-257
        p->k[0] = log(sqrt(2 * M_PI));
        p->k[1] = log(p->a[1]);
        p->k[2] = 1.0 / p->a[1];                                               // 1 / sigma
        p->k[3] = 1.0 / sqrt(2 * M_PI) / p->a[1] / p->a[2] * inv_norm;         // pdf prefactor incl. 1 / _norm
        break;
    case ISO_PRIOR_LOGNORMAL:  // priors.py:272-280
        p->k[0] = log(1.0 / sqrt(2 * M_PI));
        p->k[1] = 1.0 / sqrt(2 * M_PI) / p->a[1] * inv_norm;                   // pdf prefactor: .. / (s y) .. / scale
        p->k[2] = 1.0 / p->a[2];                                               // 1 / scale
        p->k[3] = 1.0 / p->a[1];                                               // 1 / s
        break;
    case ISO_PRIOR_FEH:
        p->k[0] = inv_norm;
        break;
    default:
        break;
    }
}

static inline void iso_prior_fill(iso_prior *p)
{
    iso_prior_leaf_fill(&p->self);
    for (int i = 0; i < ISO_MAX_COMP; i++) {
        iso_prior_leaf_fill(&p->comp[i]);
        p->inv_norms[i] = 1.0 / p->norms[i];
    }
    p->inv_norm = 1.0 / p->self.norm;
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

// ChabrierPrior as the reference builds it: BrokenPrior([LogNormalPrior, PowerLawPrior], [bp])  (priors.py:514-519)
static inline bool iso_prior_is_chabrier_like(const iso_prior &p)
{
    return p.self.kind == ISO_PRIOR_BROKEN && p.n_comp == 2 && p.comp[0].kind == ISO_PRIOR_LOGNORMAL &&
           p.comp[1].kind == ISO_PRIOR_POWERLAW;
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
        disk = (1.0 / disk_norm) * ((0.8 / 0.15) * exp(-0.5 * (d1 * d1) * (1.0 / (0.15 * 0.15))) +
                                    (0.2 / 0.22) * exp(-0.5 * (d2 * d2) * (1.0 / (0.22 * 0.22))));
    } else {
        double d = feh - (-0.3);
        disk = (1.0 / 2.5066282746310002 / 0.3) * exp(-0.5 * (d * d) * (1.0 / (0.3 * 0.3)));
    }
    double dh = feh - (-1.5);
    double halo = (1.0 / 1.0026513098524001) * exp(-0.5 * (dh * dh) * (1.0 / (0.4 * 0.4)));   // 1 / sqrt(2 pi 0.4^2)
    return halo_fraction * halo + (1 - halo_fraction) * disk;
}

// self._pdf(x) / self._norm of a non-broken class (KIND is a compile-time constant)
template <int KIND>
__device__ __forceinline__ double iso_leaf_pdf_normed(const iso_prior_leaf &p, double x)
{
    if (KIND == ISO_PRIOR_FLAT) return p.k[2];
    if (KIND == ISO_PRIOR_FLATLOG) return p.k[2] * exp10(x);
    if (KIND == ISO_PRIOR_POWERLAW) return p.k[2] * pow(x, p.a[0]);
    if (KIND == ISO_PRIOR_GAUSSIAN) {
        double z = (x - p.a[0]) * p.k[2];
        return exp(-0.5 * (z * z)) * p.k[3];
    }
    if (KIND == ISO_PRIOR_LOGNORMAL) {
        double yv = x * p.k[2], t = log(yv) * p.k[3];
        return p.k[1] / yv * exp(-0.5 * (t * t)) * p.k[2];
    }
    if (KIND == ISO_PRIOR_FEH) return iso_feh_pdf(p, x) * p.k[0];
    return iso_nan();
}

// Prior.pdf  priors.py:54-59
template <int KIND>
__device__ __forceinline__ double iso_leaf_pdf(const iso_prior_leaf &p, double x)
{
    if ((p.flags & ISO_PF_HAS_BOUNDS) && iso_outside(x, p.lo, p.hi)) return 0.0;
    return iso_leaf_pdf_normed<KIND>(p, x);
}

// Prior.__call__ priors.py:35-36 / BoundedPrior.__call__ :112-117 (the second bounds test is the same test)
template <int KIND>
__device__ __forceinline__ double iso_leaf_call(const iso_prior_leaf &p, double x)
{
    return iso_leaf_pdf<KIND>(p, x);
}

// Prior.lnpdf priors.py:61-66 / BoundedPrior.lnpdf :131-140 of a non-broken class.
// `lnx` lets a caller that already holds log(x) share it (only read by the POWERLAW / LOGNORMAL kinds).
template <int KIND, bool HAVE_LNX = false>
__device__ __forceinline__ double iso_leaf_lnpdf(const iso_prior_leaf &p, double x, double lnx = 0.0)
{
    constexpr bool has_lnpdf = KIND == ISO_PRIOR_POWERLAW || KIND == ISO_PRIOR_GAUSSIAN || KIND == ISO_PRIOR_LOGNORMAL;
    if ((p.flags & ISO_PF_BOUNDED) && (p.flags & ISO_PF_HAS_BOUNDS) && iso_outside(x, p.lo, p.hi)) return iso_neg_inf();
    if (has_lnpdf) {   // self._lnpdf(x): no bounds test for plain Prior subclasses
        if (KIND == ISO_PRIOR_POWERLAW)   // priors.py:476-480
            return p.k[1] + p.a[0] * (HAVE_LNX ? lnx : log(x));
        if (KIND == ISO_PRIOR_GAUSSIAN) { // priors.py:256-257, 26-27
            double z = (x - p.a[0]) * p.k[2];
            return (-0.5 * (z * z) - p.k[0]) - p.k[1] - p.a[3];
        }
        // LOGNORMAL priors.py:277-280: log(y) with y = x / scale
        double ly = HAVE_LNX ? lnx - p.a[0] : log(x * p.k[2]), t = ly * p.k[3];
        return p.k[0] - (p.a[3] + ly) - 0.5 * (t * t) - p.a[0];
    }
    if (KIND == ISO_PRIOR_FLAT) {
        // log(pdf / _norm) is x-independent: evaluated once at staging time.  Plain-Prior semantics (bounds only
        // through pdf) cannot occur: FlatPrior is a BoundedPrior.
        if (!(p.flags & ISO_PF_BOUNDED) && (p.flags & ISO_PF_HAS_BOUNDS) && iso_outside(x, p.lo, p.hi)) return iso_neg_inf();
        return p.k[1];
    }
    return iso_log_or_neginf(iso_leaf_pdf<KIND>(p, x));
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

// BrokenPrior of two components with compile-time kinds (Chabrier: LOGNORMAL below the break, POWERLAW above)
template <int K0, int K1, bool HAVE_LNX = false>
__device__ __forceinline__ double iso_broken2_lnpdf(const iso_prior &p, double x, double lnx = 0.0)
{
    // priors.py:209-211 — no bounds test on this path; NaN goes to the last component
    if (x != x || p.breakpoints[0] <= x) return iso_leaf_lnpdf<K1, HAVE_LNX>(p.comp[1], x, lnx) - p.lognorms[1];
    return iso_leaf_lnpdf<K0, HAVE_LNX>(p.comp[0], x, lnx) - p.lognorms[0];
}

template <int K0, int K1>
__device__ __forceinline__ double iso_broken2_call(const iso_prior &p, double x)
{
    // BrokenPrior is a plain Prior: __call__ = pdf (bounds test when set) of _pdf / _norm   priors.py:35-36, 54-59, 205-207
    if ((p.self.flags & ISO_PF_HAS_BOUNDS) && iso_outside(x, p.self.lo, p.self.hi)) return 0.0;
    double raw = (x != x || p.breakpoints[0] <= x) ? iso_leaf_call<K1>(p.comp[1], x) * p.inv_norms[1]
                                                   : iso_leaf_call<K0>(p.comp[0], x) * p.inv_norms[0];
    return raw * p.inv_norm;
}

// ---- run-time dispatch (any supported prior object); one out-of-line copy each --------------------------------
static __device__ __noinline__ double iso_leaf_call_dyn(const iso_prior_leaf *p, double x)
{
    switch (p->kind) {
    case ISO_PRIOR_FLAT: return iso_leaf_call<ISO_PRIOR_FLAT>(*p, x);
    case ISO_PRIOR_FLATLOG: return iso_leaf_call<ISO_PRIOR_FLATLOG>(*p, x);
    case ISO_PRIOR_POWERLAW: return iso_leaf_call<ISO_PRIOR_POWERLAW>(*p, x);
    case ISO_PRIOR_GAUSSIAN: return iso_leaf_call<ISO_PRIOR_GAUSSIAN>(*p, x);
    case ISO_PRIOR_LOGNORMAL: return iso_leaf_call<ISO_PRIOR_LOGNORMAL>(*p, x);
    case ISO_PRIOR_FEH: return iso_leaf_call<ISO_PRIOR_FEH>(*p, x);
    default: return iso_nan();
    }
}

static __device__ __noinline__ double iso_leaf_lnpdf_dyn(const iso_prior_leaf *p, double x)
{
    switch (p->kind) {
    case ISO_PRIOR_FLAT: return iso_leaf_lnpdf<ISO_PRIOR_FLAT>(*p, x);
    case ISO_PRIOR_FLATLOG: return iso_leaf_lnpdf<ISO_PRIOR_FLATLOG>(*p, x);
    case ISO_PRIOR_POWERLAW: return iso_leaf_lnpdf<ISO_PRIOR_POWERLAW>(*p, x);
    case ISO_PRIOR_GAUSSIAN: return iso_leaf_lnpdf<ISO_PRIOR_GAUSSIAN>(*p, x);
    case ISO_PRIOR_LOGNORMAL: return iso_leaf_lnpdf<ISO_PRIOR_LOGNORMAL>(*p, x);
    case ISO_PRIOR_FEH: return iso_leaf_lnpdf<ISO_PRIOR_FEH>(*p, x);
    default: return iso_nan();
    }
}

// any prior object: __call__(x)
__device__ __forceinline__ double iso_prior_call_dyn(const iso_prior *p, double x)
{
    if (p->self.kind != ISO_PRIOR_BROKEN) return iso_leaf_call_dyn(&p->self, x);
    if ((p->self.flags & ISO_PF_HAS_BOUNDS) && iso_outside(x, p->self.lo, p->self.hi)) return 0.0;
    int i = iso_digitize(*p, x);
    return iso_leaf_call_dyn(&p->comp[i], x) * p->inv_norms[i] * p->inv_norm;   // priors.py:205-207
}

// any prior object: lnpdf(x)
__device__ __forceinline__ double iso_prior_lnpdf_dyn(const iso_prior *p, double x)
{
    if (p->self.kind != ISO_PRIOR_BROKEN) return iso_leaf_lnpdf_dyn(&p->self, x);
    int i = iso_digitize(*p, x);   // priors.py:209-211 — no bounds test on this path
    return iso_leaf_lnpdf_dyn(&p->comp[i], x) - p->lognorms[i];
}

#endif  // __CUDACC__
