/*
 * TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, IEEE double, no fused multiply-add) of the reference's
 * lnpost hot path, used only as the parity checker by tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs.  Nothing under isochrones_b200/
 * links, loads or calls this.
 *
 * Parity status: PINNED — tests/test_oracle_vs_golden.py checks this restatement against
 * outputs of the unmodified reference (timothydmorton/isochrones @ ac230d8a, run in the
 * build container through oracle/ref_shim.py; vectors committed under tests/golden/ by
 * oracle/make_golden.py), against the reference's own tests/test_interp.py case and the
 * data-free known answers in docs/interpolate.ipynb.
 *
 * Every function cites the reference file:line it follows (paths relative to the
 * reference root, package dir isochrones/).
 */
#ifndef ISO_ORACLE_H
#define ISO_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_BANDS 32
#define ORC_MAX_COMP 3

/* dense grid[n0, n1, (n2, (n3,)) ncols] float64, C order, columns innermost (interp.py:607-609) */
typedef struct {
    int32_t ndim;            /* 2, 3 or 4 */
    int32_t ncols;
    int64_t n[4];            /* axis lengths */
    const double *grid;
    const double *axes[4];   /* DFInterpolator.index_columns (interp.py:583) */
} orc_grid;

/* prior classes of priors.py */
enum {
    ORC_PRIOR_FLAT = 1,      /* FlatPrior       priors.py:283-293 */
    ORC_PRIOR_FLATLOG = 2,   /* FlatLogPrior    priors.py:296-306 */
    ORC_PRIOR_POWERLAW = 3,  /* PowerLawPrior   priors.py:309-342 */
    ORC_PRIOR_GAUSSIAN = 4,  /* GaussianPrior   priors.py:235-257 */
    ORC_PRIOR_LOGNORMAL = 5, /* LogNormalPrior  priors.py:260-280 */
    ORC_PRIOR_FEH = 6,       /* FehPrior        priors.py:345-381 */
    ORC_PRIOR_BROKEN = 7,    /* BrokenPrior     priors.py:143-232 */
    ORC_PRIOR_EEP = 8        /* EEP_prior       priors.py:409-429 */
};

typedef struct orc_prior {
    int32_t kind;
    int32_t bounded;         /* 1: subclass of BoundedPrior (priors.py:107-140) */
    int32_t has_bounds;      /* 0: self._bounds is None */
    int32_t local;           /* FehPrior.local */
    double lo, hi;           /* self.bounds */
    double norm;             /* self._norm (Prior.pdf divides by it, priors.py:59) */
    /* class parameters: GAUSSIAN mean,sigma,norm,lognorm | LOGNORMAL mu,sigma,scale,log_s |
       POWERLAW alpha | FEH halo_fraction */
    double a[4];
    int32_t n_comp;          /* BROKEN */
    int32_t pad_;
    double breakpoints[ORC_MAX_COMP - 1];
    double norms[ORC_MAX_COMP];
    double lognorms[ORC_MAX_COMP];
    const struct orc_prior *comp[ORC_MAX_COMP];
    const struct orc_prior *orig;   /* EEP: orig_prior */
} orc_prior;

/* the pieces of a BasicStarModel that lnprior/lnlike/lnpost read (starmodel.py:1361-1635) */
typedef struct {
    int32_t n_stars;             /* N = 1, 2, 3 */
    int32_t eep_replaces_age;    /* 1: evolution-track grid, 0: isochrone grid (models.py:666, 693) */
    int32_t index_order[5];      /* ic.param_index_order (models.py:669, 696) */
    int32_t i_Teff, i_logg, i_feh, i_Mbol;       /* model-grid column indices */
    int32_t i_orig, i_deriv;     /* EEP_prior columns: (age, dt_deep) | (mass, dm_deep) priors.py:416-419 */
    int32_t i_nu_max, i_delta_nu;
    int32_t n_bands;
    int32_t has_plax, has_nu_max, has_delta_nu;
    int32_t i_mags[ORC_MAX_BANDS];
    double spec_vals[3], spec_uncs[3];           /* NaN = absent (likelihood.py:127) */
    double mag_vals[ORC_MAX_BANDS], mag_uncs[ORC_MAX_BANDS];
    double plax, plax_unc;
    double nu_max, nu_max_unc, delta_nu, delta_nu_unc;
    const orc_grid *model;
    const orc_grid *bc;
    const orc_prior *prior_eep, *prior_mass, *prior_age, *prior_feh, *prior_distance, *prior_AV;
} orc_model;

int64_t orc_searchsorted(const double *arr, int64_t n, double x, int32_t *eq);
void orc_interp_value(const orc_grid *g, const double *x, const int32_t *icols, int32_t ncols, double *out);
void orc_interp_values(const orc_grid *g, const double *const *xx, int64_t N, const int32_t *icols,
                       int32_t ncols, double *out);
void orc_interp_mag(const double *pars, const int32_t *index_order, const orc_grid *model, int32_t i_Teff,
                    int32_t i_logg, int32_t i_feh, int32_t i_Mbol, const orc_grid *bc, const int32_t *bc_cols,
                    int32_t n_bands, double *Teff, double *logg, double *feh, double *mags);
void orc_interp_mags(const double *pars, int64_t N, const int32_t *index_order, const orc_grid *model,
                     int32_t i_Teff, int32_t i_logg, int32_t i_feh, int32_t i_Mbol, const orc_grid *bc,
                     const int32_t *bc_cols, int32_t n_bands, double *Teffs, double *loggs, double *fehs,
                     double *mags);
double orc_interp_eep(double x, double x0, double x1, const double *ii0, int64_t n0, const double *ii1, int64_t n1,
                      const double *arrays, int64_t n_eep, const int64_t *lengths);
void orc_interp_eeps(const double *xs, const double *x0s, const double *x1s, int64_t N, const double *ii0, int64_t n0,
                     const double *ii1, int64_t n1, const double *arrays, int64_t n_eep, const int64_t *lengths,
                     double *out);
double orc_gauss_lnprob(double val, double unc, double model_val);
double orc_fast_addmags(const double *mags, int32_t n);
double orc_prior_call(const orc_prior *p, double x);
double orc_prior_lnpdf(const orc_prior *p, double x);
double orc_eep_prior_lnpdf(const orc_model *m, double eep, double other, double feh);
double orc_lnlike(const orc_model *m, const double *pars);
double orc_lnprior(const orc_model *m, const double *pars);
double orc_lnpost(const orc_model *m, const double *pars);
/* serial loop of the scalar functions over rows of pars[N, ndim]; n_threads > 1 fans rows out with OpenMP
   (the reference has no threaded path; its batch recipe is a process pool over rows/stars) */
void orc_lnpost_batch(const orc_model *m, const double *pars, int64_t N, double *lnpost, double *lnprior,
                      double *lnlike, int32_t n_threads);
/* catalog mode: row i uses models[model_of_row[i]] */
void orc_lnpost_catalog(const orc_model *const *models, const int32_t *model_of_row, const double *pars,
                        int64_t N, double *lnpost, int32_t n_threads);
/* emcee's stretch move (the sampler the reference drives at starmodel.py:966) with the product's Philox counters:
   n_chains ensembles advanced n_steps; see iso_oracle.c */
void orc_stretch_move(const orc_model *const *models, int32_t n_models, int32_t n_chains, int32_t n_walkers, double *pos,
                      double *lnprob, int64_t step0, int32_t n_steps, uint64_t seed, double a, double *chain_out,
                      double *lnprob_out, int64_t *n_accepted, int32_t n_threads);
void orc_mnest_prior(const double *bounds_lo, const double *bounds_hi, int32_t ndim, double *cube);
int32_t orc_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
