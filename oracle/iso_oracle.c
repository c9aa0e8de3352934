/*
 * TEST INFRASTRUCTURE — NOT PRODUCT CODE.  See iso_oracle.h for scope and parity status.
 *
 * Plain-C restatement of the reference's lnpost hot path.  Compile with
 *   gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC iso_oracle.c -o _build/libiso_oracle.so -lm
 * (-ffp-contract=off: the reference's numba code performs separate multiply and add).
 */
#include "iso_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* likelihood.py:7 / priors.py:16-19 */
static double log_one_over_root_2pi(void) { return log(1.0 / sqrt(2 * M_PI)); }

/* interp.py:10-35 — binary search; returns L, *eq set when arr[m] == x was hit (then L = m). */
int64_t orc_searchsorted(const double *arr, int64_t n, double x, int32_t *eq)
{
    int64_t L = 0, R = n - 1, m;
    int done = 0;
    *eq = 0;
    m = (L + R) / 2;
    while (!done) {
        double xm = arr[m];
        if (xm < x) {
            L = m + 1;
        } else if (xm > x) {
            R = m - 1;
        } else if (xm == x) {
            L = m;
            *eq = 1;
            done = 1;
        }
        /* Python floor division: (L + R) // 2 with L + R possibly -1 */
        m = (L + R) >= 0 ? (L + R) / 2 : -1;
        if (L > R)
            done = 1;
    }
    return L;
}

/* interp.py:96-143 (3d), 146-205 (4d), 63-93 (2d): bounds test is inclusive on both ends; an exact
 * node hit gives index = node, distance = 0, otherwise index = L-1 and the normalised distance. */
static int find_indices(const orc_grid *g, const double *x, int64_t *indices, double *norm_distances)
{
    int d;
    for (d = 0; d < g->ndim; d++) {
        const double *ii = g->axes[d];
        if (x[d] < ii[0] || x[d] > ii[g->n[d] - 1])
            return 1; /* out of bounds */
    }
    for (d = 0; d < g->ndim; d++) {
        const double *ii = g->axes[d];
        int32_t eq;
        int64_t ix = orc_searchsorted(ii, g->n[d], x[d], &eq);
        if (eq) {
            indices[d] = ix;
            norm_distances[d] = 0;
        } else {
            double c0 = ii[ix - 1];
            indices[d] = ix - 1;
            norm_distances[d] = (x[d] - c0) / (ii[ix] - c0);
        }
    }
    return 0;
}

/* interp.py:208-249 (2d), 252-293 (3d), 296-338 (4d).
 * NaN in -> NaN out; out of bounds -> NaN; 2^ndim corners, dim 0 = most significant bit; weight is the
 * product over dims of (1 - y) or y; values[c] += grid[corner, icol] * weight in corner order.
 * Zero-weight corners ARE read and multiplied (NaN * 0 = NaN propagates from NaN-padded neighbours).
 * The reference indexes without bounds checks, so an "index + 1" equal to the axis length on a trailing
 * axis lands on the next row of the flat array; that flat-offset arithmetic is reproduced here.  Only a
 * corner beyond the end of the whole array (undefined behaviour in the reference, weight always 0 there)
 * is replaced by the value 0.0. */
void orc_interp_value(const orc_grid *g, const double *x, const int32_t *icols, int32_t ncols, double *out)
{
    int ndim = g->ndim, d, c;
    int64_t indices[4];
    double norm_distances[4];
    int64_t n_nodes = 1;
    int n_edges = 1 << ndim, j;

    for (d = 0; d < ndim; d++) {
        if (x[d] != x[d]) {
            for (c = 0; c < ncols; c++) out[c] = NAN;
            return;
        }
        n_nodes *= g->n[d];
    }
    if (find_indices(g, x, indices, norm_distances)) {
        for (c = 0; c < ncols; c++) out[c] = NAN;
        return;
    }
    for (c = 0; c < ncols; c++) out[c] = 0.0;

    for (j = 0; j < n_edges; j++) {
        double weight = 1.0;
        int64_t node = 0;
        for (d = 0; d < ndim; d++) {
            int64_t ei = indices[d] + ((j >> (ndim - 1 - d)) & 1);
            if (ei == indices[d])
                weight *= 1 - norm_distances[d];
            else
                weight *= norm_distances[d];
            node = node * g->n[d] + ei;
        }
        for (c = 0; c < ncols; c++) {
            double v = node < n_nodes ? g->grid[node * g->ncols + icols[c]] : 0.0;
            out[c] += v * weight;
        }
    }
}

/* interp.py:341-392 — serial loop over points; out is [N, ncols] row-major. */
void orc_interp_values(const orc_grid *g, const double *const *xx, int64_t N, const int32_t *icols,
                       int32_t ncols, double *out)
{
    int64_t i;
    for (i = 0; i < N; i++) {
        double x[4];
        int d;
        for (d = 0; d < g->ndim; d++) x[d] = xx[d][i];
        orc_interp_value(g, x, icols, ncols, out + i * ncols);
    }
}

/* mags.py:8-61 */
void orc_interp_mag(const double *pars, const int32_t *index_order, const orc_grid *model, int32_t i_Teff,
                    int32_t i_logg, int32_t i_feh, int32_t i_Mbol, const orc_grid *bc, const int32_t *bc_cols,
                    int32_t n_bands, double *Teff, double *logg, double *feh, double *mags)
{
    double x3[3], x4[4], star_props[4], bcv[ORC_MAX_BANDS], dist_mod;
    int32_t cols[4];
    int i;
    cols[0] = i_Teff; cols[1] = i_logg; cols[2] = i_feh; cols[3] = i_Mbol;
    x3[0] = pars[index_order[0]];
    x3[1] = pars[index_order[1]];
    x3[2] = pars[index_order[2]];
    orc_interp_value(model, x3, cols, 4, star_props);
    *Teff = star_props[0];
    *logg = star_props[1];
    *feh = star_props[2];
    x4[0] = *Teff; x4[1] = *logg; x4[2] = *feh;
    x4[3] = pars[index_order[4]];                       /* AV */
    orc_interp_value(bc, x4, bc_cols, n_bands, bcv);
    dist_mod = 5 * log10(pars[index_order[3]] / 10.0);
    for (i = 0; i < n_bands; i++)
        mags[i] = star_props[3] + dist_mod - bcv[i];
}

/* mags.py:64-124 — pars is [5, N] parameter-major; mags is [N, n_bands]. */
void orc_interp_mags(const double *pars, int64_t N, const int32_t *index_order, const orc_grid *model,
                     int32_t i_Teff, int32_t i_logg, int32_t i_feh, int32_t i_Mbol, const orc_grid *bc,
                     const int32_t *bc_cols, int32_t n_bands, double *Teffs, double *loggs, double *fehs,
                     double *mags)
{
    int64_t i;
    for (i = 0; i < N; i++) {
        double p[5];
        int j;
        for (j = 0; j < 5; j++) p[j] = pars[j * N + i];
        orc_interp_mag(p, index_order, model, i_Teff, i_logg, i_feh, i_Mbol, bc, bc_cols, n_bands, &Teffs[i],
                       &loggs[i], &fehs[i], mags + i * n_bands);
    }
}

/* interp.py:502-558 interp_eep (+ :488-499 interp_eeps): (age x, feh x0, mass x1) -> EEP on an evolution-track grid.
 * find_indices_2d (interp.py:63-93) brackets (feh, mass); searchsorted (interp.py:10-35) over the first lengths[t]
 * ages of each of the four bracketing tracks gives the EEP index (EEP = index + 1); a track that ends before x
 * borrows its mass-neighbour's EEP; the four EEPs are blended bilinearly.  `arrays` is [n0 * n1, n_eep].
 * The reference indexes tracks i0 * n1 + (i1 + 1) and (i0 + 1) * n1 + .. without bounds checks; a track index beyond
 * the array (weight always 0 there) is treated as an empty track. */
double orc_interp_eep(double x, double x0, double x1, const double *ii0, int64_t n0, const double *ii1, int64_t n1,
                      const double *arrays, int64_t n_eep, const int64_t *lengths)
{
    int64_t i0, i1, ind[4], i_eep[4], len[4], n_tracks = n0 * n1;
    double d0, d1, eep[4], eep_0, eep_1;
    int32_t eq;
    int k;
    if (x != x || x0 != x0 || x1 != x1) return NAN;
    if (x0 < ii0[0] || x0 > ii0[n0 - 1] || x1 < ii1[0] || x1 > ii1[n1 - 1]) return NAN;
    i0 = orc_searchsorted(ii0, n0, x0, &eq);
    if (eq) d0 = 0; else { i0 -= 1; d0 = (x0 - ii0[i0]) / (ii0[i0 + 1] - ii0[i0]); }
    i1 = orc_searchsorted(ii1, n1, x1, &eq);
    if (eq) d1 = 0; else { i1 -= 1; d1 = (x1 - ii1[i1]) / (ii1[i1 + 1] - ii1[i1]); }
    ind[0] = i0 * n1 + i1;            /* 00 */
    ind[1] = i0 * n1 + (i1 + 1);      /* 01 */
    ind[2] = (i0 + 1) * n1 + i1;      /* 10 */
    ind[3] = (i0 + 1) * n1 + (i1 + 1);/* 11 */
    for (k = 0; k < 4; k++) {
        if (ind[k] < n_tracks && lengths[ind[k]] > 0) {
            len[k] = lengths[ind[k]];
            i_eep[k] = orc_searchsorted(arrays + ind[k] * n_eep, len[k], x, &eq);
        } else {
            len[k] = 0;
            i_eep[k] = 0;
        }
        if (i_eep[k] > n_eep - 1) return NAN;   /* max_i_eep = weight_arrays.shape[1] - 1 */
        eep[k] = (double)(i_eep[k] + 1);
    }
    if (i_eep[0] >= len[0]) eep[0] = eep[1];   /* sequential, as written in the reference */
    if (i_eep[1] >= len[1]) eep[1] = eep[0];
    if (i_eep[2] >= len[2]) eep[2] = eep[3];
    if (i_eep[3] >= len[3]) eep[3] = eep[2];
    eep_0 = (1 - d1) * eep[0] + d1 * eep[1];
    eep_1 = (1 - d1) * eep[2] + d1 * eep[3];
    return (1 - d0) * eep_0 + d0 * eep_1;
}

void orc_interp_eeps(const double *xs, const double *x0s, const double *x1s, int64_t N, const double *ii0, int64_t n0,
                     const double *ii1, int64_t n1, const double *arrays, int64_t n_eep, const int64_t *lengths,
                     double *out)
{
    int64_t i;
    for (i = 0; i < N; i++) out[i] = orc_interp_eep(xs[i], x0s[i], x1s[i], ii0, n0, ii1, n1, arrays, n_eep, lengths);
}

/* likelihood.py:10-13 — note the + log(unc) */
double orc_gauss_lnprob(double val, double unc, double model_val)
{
    double resid = val - model_val;
    return log_one_over_root_2pi() + log(unc) - 0.5 * resid * resid / (unc * unc);
}

/* utils.py:67-75 */
double orc_fast_addmags(const double *mags, int32_t n)
{
    double tot = 0;
    int i;
    for (i = 0; i < n; i++) tot += pow(10.0, -0.4 * mags[i]);
    return -2.5 * log10(tot);
}

/* ------------------------------------------------------------------ priors.py */

static double prior_pdf_raw(const orc_prior *p, double x);   /* self._pdf(x) */
static double prior_pdf(const orc_prior *p, double x);       /* self.pdf(x)  (Prior.pdf, priors.py:54-59) */

/* np.digitize(x, breakpoints) for increasing bins: number of bins b with b <= x; NaN sorts last */
static int digitize(double x, const double *bp, int n)
{
    int i = 0;
    if (x != x) return n;
    while (i < n && bp[i] <= x) i++;
    return i;
}

/* FehPrior._pdf priors.py:359-381 */
static double feh_pdf(const orc_prior *p, double feh)
{
    double halo_fraction = p->a[0], disk_fehdist, halo_fehdist;
    double halo_mu = -1.5, halo_sig = 0.4;
    if (p->local) {
        double disk_norm = 2.5066282746310007;
        disk_fehdist = 1.0 / disk_norm *
                       (0.8 / 0.15 * exp(-0.5 * pow(feh - 0.016, 2.0) / pow(0.15, 2.0)) +
                        0.2 / 0.22 * exp(-0.5 * pow(feh + 0.15, 2.0) / pow(0.22, 2.0)));
    } else {
        double mu = -0.3, sig = 0.3;
        disk_fehdist = 1.0 / sqrt(2 * M_PI) / sig * exp(-0.5 * pow(feh - mu, 2) / pow(sig, 2));
    }
    halo_fehdist = 1.0 / sqrt(2 * M_PI * pow(halo_sig, 2)) * exp(-0.5 * pow(feh - halo_mu, 2) / pow(halo_sig, 2));
    return halo_fraction * halo_fehdist + (1 - halo_fraction) * disk_fehdist;
}

static double powerlaw_C(const orc_prior *p)
{
    double alpha = p->a[0]; /* priors.py:316, 322 */
    return (1 + alpha) / (pow(p->hi, 1 + alpha) - pow(p->lo, 1 + alpha));
}

static double prior_pdf_raw(const orc_prior *p, double x)
{
    switch (p->kind) {
    case ORC_PRIOR_FLAT: /* priors.py:287-289 */
        return 1.0 / (p->hi - p->lo);
    case ORC_PRIOR_FLATLOG: /* priors.py:300-302 */
        return log(10) * pow(10, x) / (pow(10, p->hi) - pow(10, p->lo));
    case ORC_PRIOR_POWERLAW: /* priors.py:314-318 */
        return powerlaw_C(p) * pow(x, p->a[0]);
    case ORC_PRIOR_GAUSSIAN: { /* priors.py:253-254, 22-23 */
        double z = (x - p->a[0]) / p->a[1];
        return exp(-pow(z, 2) / 2.0) / sqrt(2 * M_PI) / p->a[1] / p->a[2];
    }
    case ORC_PRIOR_LOGNORMAL: { /* priors.py:272-275 */
        double s = p->a[1], y = x / p->a[2];
        return (1.0 / sqrt(2 * M_PI)) / (s * y) * exp(-0.5 * pow(log(y) / s, 2)) / p->a[2];
    }
    case ORC_PRIOR_FEH:
        return feh_pdf(p, x);
    case ORC_PRIOR_BROKEN: { /* priors.py:205-207 */
        int i = digitize(x, p->breakpoints, p->n_comp - 1);
        return orc_prior_call(p->comp[i], x) / p->norms[i];
    }
    default:
        return NAN;
    }
}

/* Prior.pdf priors.py:54-59 (bounds property: (-inf, inf) when _bounds is None, priors.py:38-40) */
static double prior_pdf(const orc_prior *p, double x)
{
    if (p->has_bounds && (x < p->lo || x > p->hi))
        return 0;
    return prior_pdf_raw(p, x) / p->norm;
}

/* Prior.__call__ priors.py:35-36 / BoundedPrior.__call__ priors.py:112-117 */
double orc_prior_call(const orc_prior *p, double x)
{
    if (p->bounded && p->has_bounds && (x < p->lo || x > p->hi))
        return 0;
    return prior_pdf(p, x);
}

static int has_lnpdf(const orc_prior *p)
{
    return p->kind == ORC_PRIOR_POWERLAW || p->kind == ORC_PRIOR_GAUSSIAN || p->kind == ORC_PRIOR_LOGNORMAL ||
           p->kind == ORC_PRIOR_BROKEN;
}

static double prior_lnpdf_raw(const orc_prior *p, double x) /* self._lnpdf(x) */
{
    switch (p->kind) {
    case ORC_PRIOR_POWERLAW: /* priors.py:320-323 */
        return log(powerlaw_C(p)) + p->a[0] * log(x);
    case ORC_PRIOR_GAUSSIAN: { /* priors.py:256-257, 26-27 */
        double z = (x - p->a[0]) / p->a[1];
        return (-pow(z, 2) / 2.0 - log(sqrt(2 * M_PI))) - log(p->a[1]) - p->a[3];
    }
    case ORC_PRIOR_LOGNORMAL: { /* priors.py:277-280 */
        double s = p->a[1], y = x / p->a[2];
        return log(1.0 / sqrt(2 * M_PI)) - (p->a[3] + log(y)) - 0.5 * pow(log(y) / s, 2) - p->a[0];
    }
    case ORC_PRIOR_BROKEN: { /* priors.py:209-211 */
        int i = digitize(x, p->breakpoints, p->n_comp - 1);
        return orc_prior_lnpdf(p->comp[i], x) - p->lognorms[i];
    }
    default:
        return NAN;
    }
}

static double log_or_neginf(double pdf) /* `np.log(pdf) if pdf else -np.inf` priors.py:66, 140 */
{
    if (pdf == 0) return -INFINITY;
    return log(pdf);
}

/* Prior.lnpdf priors.py:61-66 / BoundedPrior.lnpdf priors.py:131-140 (not valid for EEP priors, which
 * need keyword arguments — see orc_eep_prior_lnpdf) */
double orc_prior_lnpdf(const orc_prior *p, double x)
{
    if (p->bounded) {
        if (p->has_bounds && (x < p->lo || x > p->hi))
            return -INFINITY;
        if (has_lnpdf(p))
            return prior_lnpdf_raw(p, x);
        return log_or_neginf(prior_pdf(p, x));
    }
    if (has_lnpdf(p))
        return prior_lnpdf_raw(p, x);
    return log_or_neginf(orc_prior_call(p, x));
}

/* EEP_prior.lnpdf(eep, mass=|age=, feh=): BoundedPrior.lnpdf priors.py:131-140 -> Prior.pdf :54-59 ->
 * EEP_prior._pdf :423-429 -> ic.interp_value (models.py:390-400) on [orig_par, deriv_prop]. */
double orc_eep_prior_lnpdf(const orc_model *m, double eep, double other, double feh)
{
    const orc_prior *p = m->prior_eep;
    double pars[3], x[3], vals[2], pdf;
    int32_t cols[2];
    if (p->has_bounds && (eep < p->lo || eep > p->hi))
        return -INFINITY;
    /* Prior.pdf repeats the bounds test (same outcome), then _pdf / _norm */
    if (m->eep_replaces_age) { /* pars = [mass, eep, feh] */
        pars[0] = other; pars[1] = eep; pars[2] = feh;
    } else { /* pars = [eep, age, feh] */
        pars[0] = eep; pars[1] = other; pars[2] = feh;
    }
    x[0] = pars[m->index_order[0]];
    x[1] = pars[m->index_order[1]];
    x[2] = pars[m->index_order[2]];
    cols[0] = m->i_orig;
    cols[1] = m->i_deriv;
    orc_interp_value(m->model, x, cols, 2, vals);
    pdf = orc_prior_call(p->orig, vals[0]) * vals[1];
    pdf = pdf / p->norm;
    return log_or_neginf(pdf);
}

/* ------------------------------------------------------------------ likelihood.py / starmodel.py */

/* star_lnlike likelihood.py:16-147 wrapped by BasicStarModel.lnlike starmodel.py:1563-1614 */
double orc_lnlike(const orc_model *m, const double *pars)
{
    int n_pars = 4 + m->n_stars, k, i;
    double star_pars[3][5], Teff = NAN, logg = NAN, feh = NAN;
    double mags[3][ORC_MAX_BANDS], tot[ORC_MAX_BANDS], lnlike = 0;
    double primary_pars[5];

    for (k = 0; k < m->n_stars; k++) { /* likelihood.py:43-54 */
        star_pars[k][0] = pars[k];
        for (i = 1; i < 5; i++) star_pars[k][i] = pars[m->n_stars - 1 + i];
    }
    memcpy(primary_pars, star_pars[0], sizeof(primary_pars));

    for (k = 0; k < m->n_stars; k++) {
        double t, g, f;
        orc_interp_mag(star_pars[k], m->index_order, m->model, m->i_Teff, m->i_logg, m->i_feh, m->i_Mbol,
                       m->bc, m->i_mags, m->n_bands, &t, &g, &f, mags[k]);
        if (k == 0) { Teff = t; logg = g; feh = f; } /* companions' Teff/logg/feh discarded :76, :96 */
    }
    for (i = 0; i < m->n_bands; i++) { /* likelihood.py:115-120 */
        if (m->n_stars == 1) {
            tot[i] = mags[0][i];
        } else {
            double mm[3];
            for (k = 0; k < m->n_stars; k++) mm[k] = mags[k][i];
            tot[i] = orc_fast_addmags(mm, m->n_stars);
        }
    }
    { /* likelihood.py:122-140 — skipped when the observed value is NaN */
        double model_vals[3];
        model_vals[0] = Teff; model_vals[1] = logg; model_vals[2] = feh;
        for (i = 0; i < 3; i++) {
            double val = m->spec_vals[i];
            if (val == val)
                lnlike += orc_gauss_lnprob(val, m->spec_uncs[i], model_vals[i]);
        }
    }
    for (i = 0; i < m->n_bands; i++) /* likelihood.py:142-145 */
        lnlike += orc_gauss_lnprob(m->mag_vals[i], m->mag_uncs[i], tot[i]);

    if (m->has_plax) /* starmodel.py:1599-1601; distance_index = n_stars + 2 (starmodel.py:1398-1419) */
        lnlike += orc_gauss_lnprob(m->plax, m->plax_unc, 1000.0 / pars[n_pars - 2]);

    if (m->has_nu_max) { /* starmodel.py:1604-1612 */
        double x[3], vals[2];
        int32_t cols[2];
        x[0] = primary_pars[m->index_order[0]];
        x[1] = primary_pars[m->index_order[1]];
        x[2] = primary_pars[m->index_order[2]];
        cols[0] = m->i_nu_max;
        cols[1] = m->i_delta_nu;
        orc_interp_value(m->model, x, cols, 2, vals);
        lnlike += orc_gauss_lnprob(m->nu_max, m->nu_max_unc, vals[0]);
        if (m->has_delta_nu) /* the reference passes delta_nu as its own uncertainty (:1612) */
            lnlike += orc_gauss_lnprob(m->delta_nu, m->delta_nu, vals[1]);
    }
    return lnlike;
}

/* BasicStarModel.lnprior starmodel.py:1616-1635 */
double orc_lnprior(const orc_model *m, const double *pars)
{
    int N = m->n_stars, i;
    double lnp = 0;
    if (N == 2) {
        if (pars[1] > pars[0]) return -INFINITY;
    } else if (N == 3) {
        if (!(pars[0] > pars[1]) && (pars[1] > pars[2])) return -INFINITY; /* precedence as written */
    }
    if (m->eep_replaces_age) { /* param_names = (mass, eep, feh, distance, AV), models.py:665 */
        lnp += orc_prior_lnpdf(m->prior_mass, pars[0]);
        lnp += orc_eep_prior_lnpdf(m, pars[1], pars[0], pars[2]); /* mass_index 0, feh_index 2 */
        lnp += orc_prior_lnpdf(m->prior_feh, pars[2]);
        lnp += orc_prior_lnpdf(m->prior_distance, pars[3]);
        lnp += orc_prior_lnpdf(m->prior_AV, pars[4]);
    } else { /* (eep[_0.._N-1], age, feh, distance, AV), models.py:692, starmodel.py:1510-1518 */
        double age = pars[N], feh = pars[N + 1];
        for (i = 0; i < N; i++)
            lnp += orc_eep_prior_lnpdf(m, pars[i], age, feh);
        lnp += orc_prior_lnpdf(m->prior_age, age);
        lnp += orc_prior_lnpdf(m->prior_feh, feh);
        lnp += orc_prior_lnpdf(m->prior_distance, pars[N + 2]);
        lnp += orc_prior_lnpdf(m->prior_AV, pars[N + 3]);
    }
    return lnp;
}

/* StarModel.lnpost starmodel.py:538-542 */
double orc_lnpost(const orc_model *m, const double *pars)
{
    double lnpr = orc_lnprior(m, pars);
    if (!isfinite(lnpr)) return -INFINITY;
    return lnpr + orc_lnlike(m, pars);
}

void orc_lnpost_batch(const orc_model *m, const double *pars, int64_t N, double *lnpost, double *lnprior,
                      double *lnlike, int32_t n_threads)
{
    int ndim = 4 + m->n_stars;
    int64_t i;
    (void)n_threads;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(n_threads > 0 ? n_threads : 1)
#endif
    for (i = 0; i < N; i++) {
        const double *p = pars + i * ndim;
        double lnpr = orc_lnprior(m, p);
        double ll = NAN;
        if (lnprior) lnprior[i] = lnpr;
        if (lnlike || isfinite(lnpr)) ll = orc_lnlike(m, p);
        if (lnlike) lnlike[i] = ll;
        if (lnpost) lnpost[i] = isfinite(lnpr) ? lnpr + ll : -INFINITY;
    }
}

void orc_lnpost_catalog(const orc_model *const *models, const int32_t *model_of_row, const double *pars,
                        int64_t N, double *lnpost, int32_t n_threads)
{
    int64_t i;
    (void)n_threads;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(n_threads > 0 ? n_threads : 1)
#endif
    for (i = 0; i < N; i++) {
        const orc_model *m = models[model_of_row[i]];
        lnpost[i] = orc_lnpost(m, pars + i * (4 + m->n_stars));
    }
}

/* BasicStarModel.mnest_prior starmodel.py:1637-1640 */
void orc_mnest_prior(const double *bounds_lo, const double *bounds_hi, int32_t ndim, double *cube)
{
    int i;
    for (i = 0; i < ndim; i++) cube[i] = (bounds_hi[i] - bounds_lo[i]) * cube[i] + bounds_lo[i];
}

/* ---------------------------------------------------------------------------------------------------------------
 * Ensemble sampler driver (test infrastructure / CPU baseline of BASELINE configs 3 and 4).
 *
 * The reference fits with emcee.EnsembleSampler(nwalkers, npars, self.lnpost) (starmodel.py:966, fit_mcmc_old
 * :889-972); emcee itself is third-party and un-vendored (setup.py:60 "emcee>=2.0", unpinned).  What is restated here is
 * its published algorithm, the stretch move of Goodman & Weare (2010) as emcee 2.x runs it: the ensemble is split in two
 * halves updated in turn; walker k of the active half draws a partner c_j uniformly from the other half and
 *     z = ((a - 1) u + 1)^2 / a,  q = c_j - z (c_j - x_k),  accepted iff  ln u' < (ndim - 1) ln z + lnpost(q) - lnpost(x_k).
 * The uniforms come from Philox4x32-10 (Salmon et al. 2011) with the counters the product's device sampler uses
 * (isochrones_b200/csrc/iso_stretch.cuh), so that a product chain can be replayed decision for decision.
 * --------------------------------------------------------------------------------------------------------------- */
static void orc_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4])
{
    int r;
    for (r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0;
    out[1] = c1;
    out[2] = c2;
    out[3] = c3;
}

static double orc_u01(uint32_t hi, uint32_t lo)
{
    return (double)((((uint64_t)hi << 32) | lo) >> 11) * (1.0 / 9007199254740992.0);
}

/* one proposal + accept test of walker k; returns 1 when accepted (pos / lnprob updated in place) */
static int orc_stretch_one(const orc_model *m, double *pos, double *lnprob, int ndim, int nhalf, int half, int k, int chain,
                           uint64_t gstep, uint64_t seed, double a)
{
    uint32_t r[4], r2[4];
    uint64_t ctr = gstep * 2 + (uint64_t)half;
    uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32) ^ ((uint32_t)chain << 8);
    uint32_t k0 = (uint32_t)(seed & 0xffffffffu), k1 = (uint32_t)(seed >> 32);
    double q[8], u, u_acc, zr, z, lq, lnpdiff;
    int j, d;
    orc_philox4x32_10(c0, c1, (uint32_t)k, 0u, k0, k1, r);
    orc_philox4x32_10(c0, c1, (uint32_t)k, 1u, k0, k1, r2);
    u = orc_u01(r[0], r[1]);
    j = (1 - half) * nhalf + (int)(r[2] % (uint32_t)nhalf);
    u_acc = orc_u01(r2[0], r2[1]);
    zr = (a - 1.0) * u + 1.0;
    z = zr * zr / a;
    for (d = 0; d < ndim; d++) {
        double c = pos[j * ndim + d], x = pos[k * ndim + d];
        q[d] = c - (c - x) * z;
    }
    lq = orc_lnpost(m, q);
    lnpdiff = (ndim - 1) * log(z) + lq - lnprob[k];
    if (lnpdiff > log(u_acc)) { /* NaN compares false: a NaN lnpost is a rejection */
        for (d = 0; d < ndim; d++) pos[k * ndim + d] = q[d];
        lnprob[k] = lq;
        return 1;
    }
    return 0;
}

/* n_chains independent ensembles of n_walkers walkers advanced by n_steps; chain c samples models[c % n_models].
 * pos [n_chains, n_walkers, ndim] and lnprob [n_chains, n_walkers] are updated in place (lnprob must hold lnpost of pos
 * on entry); chain_out [n_steps, n_chains, n_walkers, ndim] / lnprob_out [n_steps, n_chains, n_walkers] may be NULL.
 * Threads: over chains when there are several, else over the walkers of a half-step (they are independent). */
void orc_stretch_move(const orc_model *const *models, int32_t n_models, int32_t n_chains, int32_t n_walkers, double *pos,
                      double *lnprob, int64_t step0, int32_t n_steps, uint64_t seed, double a, double *chain_out,
                      double *lnprob_out, int64_t *n_accepted, int32_t n_threads)
{
    int ndim = 4 + models[0]->n_stars, nhalf = n_walkers / 2;
    int c;
    (void)n_threads;
    if (n_chains > 1) {
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads > 0 ? n_threads : 1)
#endif
        for (c = 0; c < n_chains; c++) {
            const orc_model *m = models[c % n_models];
            double *p = pos + (size_t)c * n_walkers * ndim, *lp = lnprob + (size_t)c * n_walkers;
            int64_t acc = 0;
            int s, half, t;
            for (s = 0; s < n_steps; s++) {
                for (half = 0; half < 2; half++)
                    for (t = 0; t < nhalf; t++)
                        acc += orc_stretch_one(m, p, lp, ndim, nhalf, half, half * nhalf + t, c, (uint64_t)(step0 + s), seed, a);
                if (chain_out)
                    memcpy(chain_out + ((size_t)s * n_chains + c) * n_walkers * ndim, p, sizeof(double) * n_walkers * ndim);
                if (lnprob_out) memcpy(lnprob_out + ((size_t)s * n_chains + c) * n_walkers, lp, sizeof(double) * n_walkers);
            }
            if (n_accepted) n_accepted[c] = acc;
        }
    } else {
        const orc_model *m = models[0];
        int64_t acc = 0;
        int s, half, t;
        for (s = 0; s < n_steps; s++) {
            for (half = 0; half < 2; half++) {
#ifdef _OPENMP
#pragma omp parallel for schedule(static) reduction(+ : acc) num_threads(n_threads > 0 ? n_threads : 1)
#endif
                for (t = 0; t < nhalf; t++)
                    acc += orc_stretch_one(m, pos, lnprob, ndim, nhalf, half, half * nhalf + t, 0, (uint64_t)(step0 + s), seed, a);
            }
            if (chain_out) memcpy(chain_out + (size_t)s * n_walkers * ndim, pos, sizeof(double) * n_walkers * ndim);
            if (lnprob_out) memcpy(lnprob_out + (size_t)s * n_walkers, lnprob, sizeof(double) * n_walkers);
        }
        if (n_accepted) n_accepted[0] = acc;
    }
}

int32_t orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
