/*
 * isochrones_b200 — C ABI of the B200-native lnpost hot path.
 *
 * The reference (timothydmorton/isochrones @ ac230d8a, pure Python + numba) has no FFI: its seam is a
 * set of Python callables.  Each entry point below replaces the numba function(s) cited beside it
 * (paths relative to the reference root, package dir isochrones/); the Python host layer in
 * isochrones_b200/ keeps the reference's call signatures and binds these symbols with ctypes
 * (INTEGRATION.md shows the stub a maintainer of the reference would add).
 *
 * Conventions
 *   - plain C types only; every function returns 0 on success or a negative ISO_E_* code, and
 *     iso_last_error(ctx) returns a human-readable description of the last failure on that context
 *     (ctx == NULL: the last failure of a call that had no context);
 *   - one context per GPU; calls on one context are serialised on its CUDA streams, different
 *     contexts are independent;
 *   - `const double *` / `double *` parameters named h_* (or documented as host) are caller-owned
 *     HOST buffers; parameters named d_* are DEVICE pointers obtained from iso_dev_alloc;
 *   - all floating point data is IEEE float64 (the reference dtype, interp.py:609);
 *   - NaN means "outside the grid", -inf means "zero prior probability" (SURVEY.md §8b);
 *   - there is no CPU fallback: without a CUDA device every compute call fails with ISO_E_CUDA.
 */
#ifndef ISOCHRONES_B200_H
#define ISOCHRONES_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ISO_ABI_VERSION 3

#define ISO_OK 0
#define ISO_E_INVALID (-1) /* bad argument */
#define ISO_E_CUDA (-2)    /* CUDA runtime / driver failure (including "no device") */
#define ISO_E_NOMEM (-3)
#define ISO_E_NCCL (-4)
#define ISO_E_UNSUPPORTED (-5)
#define ISO_E_TIMEOUT (-6) /* a peer rank did not publish an exchange step in time (iso_peer_*) */

#define ISO_MAX_BANDS 16 /* photometric bands per star model */
#define ISO_MAX_COMP 3   /* components of a BrokenPrior */
#define ISO_MAX_DIM 4
#define ISO_MAX_STARS 3

typedef struct iso_ctx iso_ctx;
typedef struct iso_grid iso_grid;
typedef struct iso_models iso_models;
typedef struct iso_sampler iso_sampler;

/* ------------------------------------------------------------------------------------------------
 * context
 * ---------------------------------------------------------------------------------------------- */
int iso_abi_version(void);
/* sizeof of the public structs as this library was compiled: which = 0 iso_prior_leaf, 1 iso_prior, 2 iso_model
 * (lets a binding check its struct layout before passing one in); -1 for an unknown code */
int64_t iso_struct_size(int which);
int iso_device_count(int *count);
int iso_ctx_create(int device, iso_ctx **out);
int iso_ctx_destroy(iso_ctx *ctx);
int iso_ctx_sync(iso_ctx *ctx);
const char *iso_last_error(iso_ctx *ctx);
/* name[256], sm count, L2 bytes, total HBM bytes, compute capability major*10+minor */
int iso_ctx_info(iso_ctx *ctx, char *name, int *sm_count, int64_t *l2_bytes, int64_t *hbm_bytes, int *cc);

/* device / pinned-host memory and CUDA-event timing on the context's compute stream (bench plumbing) */
int iso_dev_alloc(iso_ctx *ctx, int64_t bytes, void **d_ptr);
int iso_dev_free(iso_ctx *ctx, void *d_ptr);
int iso_host_alloc(iso_ctx *ctx, int64_t bytes, void **h_ptr); /* page-locked */
int iso_host_free(iso_ctx *ctx, void *h_ptr);
int iso_memcpy_h2d(iso_ctx *ctx, void *d_dst, const void *h_src, int64_t bytes);
int iso_memcpy_d2h(iso_ctx *ctx, void *h_dst, const void *d_src, int64_t bytes);
int iso_memset(iso_ctx *ctx, void *d_dst, int value, int64_t bytes);
int iso_timer_start(iso_ctx *ctx);            /* cudaEventRecord on the compute stream */
int iso_timer_stop(iso_ctx *ctx, float *ms);  /* record + synchronize + elapsed */
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
int iso_launch_count(iso_ctx *ctx, int64_t *count);

/* ------------------------------------------------------------------------------------------------
 * grids — replace DFInterpolator.grid / .index_columns (interp.py:571-614)
 * ---------------------------------------------------------------------------------------------- */
/* Copy a dense grid[n0, .., n_{ndim-1}, ncols] (C order, columns innermost, interp.py:607-609) and its
 * ndim axis arrays (interp.py:583) to HBM once.  shape has ndim + 1 entries (last = ncols); ndim in 2..4. */
int iso_grid_stage(iso_ctx *ctx, const double *h_grid, int ndim, const int64_t *shape,
                   const double *const *h_axes, iso_grid **out);
/* Device-side column selection into a compact grid with `ncols_out` columns per node (>= ncols; the
 * extra columns are zero).  cols[i] < 0 also yields a zero column.  Used to build the 8-column model
 * pack and the per-model band pack that the fused lnpost kernel gathers. */
int iso_grid_repack(iso_ctx *ctx, const iso_grid *src, const int32_t *cols, int ncols, int ncols_out,
                    iso_grid **out);
int iso_grid_destroy(iso_ctx *ctx, iso_grid *grid);
int iso_grid_shape(const iso_grid *grid, int *ndim, int64_t *shape /* ISO_MAX_DIM + 1 */);

/* interp_values_2d/3d/4d (interp.py:341-392) over interp_value_* (:208-338), searchsorted (:10-35) and
 * find_indices_* (:63-205).  h_x: ndim pointers to [N] coordinate arrays; h_out: [N, ncols] row-major. */
int iso_interp_values(iso_ctx *ctx, const iso_grid *grid, const double *const *h_x, int64_t N,
                      const int32_t *icols, int ncols, double *h_out);

/* interp_mags (mags.py:64-124) over interp_mag (mags.py:8-61).  h_pars is [5, N] parameter-major exactly as
 * the reference passes it (mags.py:86-87); h_mags is [N, n_bands]. */
int iso_interp_mags(iso_ctx *ctx, const iso_grid *model, const iso_grid *bc, const int32_t index_order[5],
                    int i_Teff, int i_logg, int i_feh, int i_Mbol, const int32_t *bc_cols, int n_bands,
                    const double *h_pars, int64_t N, double *h_Teff, double *h_logg, double *h_feh,
                    double *h_mags);
/* The same with the five parameters as separate [N] arrays (what ModelGridInterpolator.interp_mag holds before the
 * reference stacks them into pars[5, N] at models.py:416-424): saves the caller that host-side copy. */
int iso_interp_mags_cols(iso_ctx *ctx, const iso_grid *model, const iso_grid *bc, const int32_t index_order[5],
                         int i_Teff, int i_logg, int i_feh, int i_Mbol, const int32_t *bc_cols, int n_bands,
                         const double *const *h_par, int64_t N, double *h_Teff, double *h_logg, double *h_feh,
                         double *h_mags);

/* Device-buffer forms of the two calls above (asynchronous on the context's compute stream; every d_* is a device
 * pointer, d_x / d_par are HOST arrays of device pointers; the column lists are host arrays): for samples that already
 * live in HBM, e.g. the chains of the on-device sampler. */
int iso_interp_values_device(iso_ctx *ctx, const iso_grid *grid, const double *const *d_x, int64_t N,
                             const int32_t *icols, int ncols, double *d_out);
int iso_interp_mags_device(iso_ctx *ctx, const iso_grid *model, const iso_grid *bc, const int32_t index_order[5],
                           int i_Teff, int i_logg, int i_feh, int i_Mbol, const int32_t *bc_cols, int n_bands,
                           const double *const *d_par, int64_t N, double *d_Teff, double *d_logg, double *d_feh,
                           double *d_mags);

/* interp_eeps (interp.py:488-499) over interp_eep (:502-558): (age, feh, mass) -> EEP on an evolution-track grid
 * staged as a 3-D (feh, mass, eep) iso_grid whose column i_age holds log10 age (the per-track age arrays of
 * StellarModelGrid.get_array_grids, models.py:171-205); h_lengths[n_feh * n_mass] is the number of populated EEPs
 * of every track.  NaN in / out of bounds / beyond the last tabulated EEP -> NaN. */
int iso_interp_eeps(iso_ctx *ctx, const iso_grid *track_grid, int i_age, const int32_t *h_lengths, const double *h_age,
                    const double *h_feh, const double *h_mass, int64_t N, double *h_eep);

/* ------------------------------------------------------------------------------------------------
 * priors — replace the lnpdf / __call__ evaluation of priors.py
 * ---------------------------------------------------------------------------------------------- */
enum {
    ISO_PRIOR_FLAT = 1,      /* FlatPrior      priors.py:283-293 (AVPrior :496-499) */
    ISO_PRIOR_FLATLOG = 2,   /* FlatLogPrior   priors.py:296-306 (AgePrior :483-488) */
    ISO_PRIOR_POWERLAW = 3,  /* PowerLawPrior  priors.py:309-342 (DistancePrior, QPrior, SalpeterPrior) */
    ISO_PRIOR_GAUSSIAN = 4,  /* GaussianPrior  priors.py:235-257 */
    ISO_PRIOR_LOGNORMAL = 5, /* LogNormalPrior priors.py:260-280 */
    ISO_PRIOR_FEH = 6,       /* FehPrior       priors.py:345-381 */
    ISO_PRIOR_BROKEN = 7     /* BrokenPrior    priors.py:143-232 (ChabrierPrior :514-519) */
};
#define ISO_PF_BOUNDED 1    /* subclass of BoundedPrior (priors.py:107-140) */
#define ISO_PF_HAS_BOUNDS 2 /* self._bounds is not None */
#define ISO_PF_LOCAL 4      /* FehPrior.local */

typedef struct {
    int32_t kind;
    int32_t flags;
    double lo, hi; /* self.bounds */
    double norm;   /* self._norm (Prior.pdf divides by it, priors.py:59) */
    /* GAUSSIAN mean, sigma, norm, lognorm | LOGNORMAL mu, sigma, scale, log_s | POWERLAW alpha |
       FEH halo_fraction */
    double a[4];
    /* derived constants (reciprocals, logs, prefactors), filled in by the library when the struct is staged;
       callers leave them 0 */
    double k[4];
} iso_prior_leaf;

typedef struct {
    iso_prior_leaf self; /* for BROKEN: kind, flags, bounds, norm of the BrokenPrior itself */
    int32_t n_comp;
    int32_t pad_;
    double breakpoints[ISO_MAX_COMP - 1];
    double norms[ISO_MAX_COMP];
    double lognorms[ISO_MAX_COMP];
    double inv_norms[ISO_MAX_COMP]; /* derived, filled in by the library */
    double inv_norm;                /* derived, filled in by the library */
    iso_prior_leaf comp[ISO_MAX_COMP];
} iso_prior;

/* lnpdf(x) / __call__(x) of one prior over a host vector (Prior.lnpdf priors.py:61-66, BoundedPrior.lnpdf
 * :131-140, Prior.__call__ :35-36, BoundedPrior.__call__ :112-117).  which: 0 = lnpdf, 1 = __call__ (pdf). */
int iso_prior_eval(iso_ctx *ctx, const iso_prior *prior, int which, const double *h_x, int64_t N, double *h_out);

/* ------------------------------------------------------------------------------------------------
 * star models — replace BasicStarModel.lnlike / lnprior / lnpost (starmodel.py:1563-1635, 538-542),
 * star_lnlike (likelihood.py:16-147), gauss_lnprob (:10-13), fast_addmags (utils.py:67-75) and
 * EEP_prior (priors.py:409-429)
 * ---------------------------------------------------------------------------------------------- */
/* column order of the 8-column "model pack" built with iso_grid_repack */
enum {
    ISO_MP_TEFF = 0, ISO_MP_LOGG = 1, ISO_MP_FEH = 2, ISO_MP_MBOL = 3,
    ISO_MP_ORIG = 4,    /* age (track grids) | mass (isochrone grids): EEP_prior.orig_par */
    ISO_MP_DERIV = 5,   /* dt_deep | dm_deep: EEP_prior.deriv_prop */
    ISO_MP_NU_MAX = 6, ISO_MP_DELTA_NU = 7,
    ISO_MP_NCOLS = 8
};

typedef struct {
    int32_t n_stars;          /* N = 1, 2, 3 (starmodel.py:1398-1419) */
    int32_t eep_replaces_age; /* 1: evolution-track grid, params (mass, eep, feh, distance, AV) models.py:665;
                                 0: isochrone grid, params (eep_0.., age, feh, distance, AV) models.py:692 */
    int32_t index_order[5];   /* ic.param_index_order (models.py:669, 696) */
    int32_t n_bands;          /* observed bands, in BasicStarModel.bands order */
    int32_t band_col[ISO_MAX_BANDS]; /* column of each observed band in the BC pack */
    int32_t has_plax, has_nu_max, has_delta_nu;
    int32_t pad_;
    double spec_val[3], spec_unc[3]; /* Teff, logg, feh; NaN value = absent (likelihood.py:127) */
    double mag_val[ISO_MAX_BANDS], mag_unc[ISO_MAX_BANDS];
    double plax, plax_unc;
    double nu_max, nu_max_unc, delta_nu, delta_nu_unc;
    /* priors in BasicStarModel._priors (starmodel.py:1441-1448) */
    double eep_lo, eep_hi, eep_norm; /* EEP_prior bounds / _norm */
    int32_t eep_has_bounds;
    int32_t pad2_;
    iso_prior eep_orig;  /* EEP_prior.orig_prior (the object captured at construction, priors.py:411) */
    iso_prior mass, age, feh, distance, AV;
} iso_model;

/* Stage n star models to the device (catalog mode: one per star). */
int iso_models_stage(iso_ctx *ctx, const iso_model *h_models, int n_models, iso_models **out);
int iso_models_destroy(iso_ctx *ctx, iso_models *models);

/* Fused lnprior + lnlike + lnpost over rows of h_pars[N, ndim] (row-major, ndim = 4 + n_stars).
 *   model_pack: 3-D grid with the ISO_MP_* columns; bc_pack: 4-D grid whose columns are addressed by
 *   iso_model.band_col.  h_model_of_row == NULL: every row uses model 0; otherwise row i uses
 *   models[h_model_of_row[i]] (all models must share n_stars).  h_lnprior / h_lnlike may be NULL; when
 *   h_lnlike is given the likelihood is evaluated for every row (as BasicStarModel.lnlike would),
 *   otherwise rows whose prior is not finite skip it (StarModel.lnpost, starmodel.py:538-542). */
int iso_lnpost_batch(iso_ctx *ctx, const iso_grid *model_pack, const iso_grid *bc_pack, const iso_models *models,
                     const int32_t *h_model_of_row, const double *h_pars, int64_t N, double *h_lnpost,
                     double *h_lnprior, double *h_lnlike);
/* Same with device-resident buffers; asynchronous on the context's compute stream. */
int iso_lnpost_batch_device(iso_ctx *ctx, const iso_grid *model_pack, const iso_grid *bc_pack,
                            const iso_models *models, const int32_t *d_model_of_row, const double *d_pars,
                            int64_t N, double *d_lnpost, double *d_lnprior, double *d_lnlike);
/* BasicStarModel.mnest_prior (starmodel.py:1637-1640): cube[N, ndim] in place, u -> (hi - lo) u + lo (ndim <= 16;
 * no device allocation per call: the bounds travel in the kernel parameter block). */
int iso_mnest_prior(iso_ctx *ctx, const double *h_lo, const double *h_hi, int ndim, double *h_cube, int64_t N);
/* mnest_prior + mnest_loglike (starmodel.py:1637-1645) of a batch of live points in ONE launch: h_cube[N, ndim] holds
 * unit-cube points on entry and the mapped parameters on return (as mnest_prior leaves the cube), h_lnpost[N] their
 * lnpost (= mnest_loglike); h_lnprior / h_lnlike may be NULL.  h_lo / h_hi: BasicStarModel.bounds per parameter. */
int iso_mnest_lnpost_batch(iso_ctx *ctx, const iso_grid *model_pack, const iso_grid *bc_pack, const iso_models *models,
                           const double *h_lo, const double *h_hi, double *h_cube, int64_t N, double *h_lnpost,
                           double *h_lnprior, double *h_lnlike);
/* N uniform draws from the box [lo, hi] evaluated where they are made: row i is the unit-cube point Philox4x32-10(seed,
 * counter = row0 + i) mapped as above — nothing is shipped to the device per row (sample_from_prior /
 * nested-sampling initialisation, starmodel.py:838-884).  h_pars[N, ndim] (may be NULL) receives the points. */
int iso_lnpost_prior_draws(iso_ctx *ctx, const iso_grid *model_pack, const iso_grid *bc_pack, const iso_models *models,
                           const double *h_lo, const double *h_hi, uint64_t seed, int64_t row0, int64_t N, double *h_pars,
                           double *h_lnpost);
/* Device-buffer form of the two calls above (asynchronous on the compute stream): rng = 0 reads unit-cube points from
 * d_cube[N, ndim]; rng = 1 draws them (d_cube ignored).  d_pars (may be NULL, may alias d_cube) receives the mapped
 * parameters. */
int iso_lnpost_cube_device(iso_ctx *ctx, const iso_grid *model_pack, const iso_grid *bc_pack, const iso_models *models,
                           const double *h_lo, const double *h_hi, const double *d_cube, int rng, uint64_t seed,
                           int64_t row0, int64_t N, double *d_pars, double *d_lnpost);

/* ------------------------------------------------------------------------------------------------
 * on-device ensemble sampler (the emcee stretch move the reference drives through
 * emcee.EnsembleSampler(nwalkers, npars, mod.lnpost), starmodel.py:966) — SURVEY.md §8f-1
 * ---------------------------------------------------------------------------------------------- */
/* One persistent CTA per chain, one thread per walker of the active half (n_walkers even, <= 1024); the initial
 * lnpost of h_p0 is evaluated at creation (a NaN there is an error, as in emcee; -inf walkers are legal).  `models` holds one model (every chain samples it) or n_chains models
 * (catalog mode: chain c samples star c).  Randomness is Philox4x32-10 keyed by `seed`. */
int iso_sampler_create(iso_ctx *ctx, const iso_grid *model_pack, const iso_grid *bc_pack, const iso_models *models,
                       int n_chains, int n_walkers, const double *h_p0 /* [n_chains, n_walkers, ndim] */,
                       uint64_t seed, double stretch_a, iso_sampler **out);
/* Run n_steps full ensemble steps in ONE kernel launch.  h_chain ([n_steps / thin, n_chains, n_walkers, ndim]) and
 * h_lnprob ([n_steps / thin, n_chains, n_walkers]) may be NULL. */
int iso_sampler_run(iso_ctx *ctx, iso_sampler *s, int n_steps, int thin, double *h_chain, double *h_lnprob);
/* Current ensemble (h_pos [n_chains, n_walkers, ndim], h_lnprob [n_chains, n_walkers]; either may be NULL),
 * accepted proposals per chain (n_accepted[n_chains], may be NULL) and proposals made so far per chain. */
int iso_sampler_state(iso_ctx *ctx, iso_sampler *s, double *h_pos, double *h_lnprob, int64_t *n_accepted,
                      int64_t *n_proposed);
/* emcee's EnsembleSampler.reset(): zero the acceptance counters (and the proposal count they are divided by) and the
 * running moments; the walkers stay where they are.  The burn-in / production split of the reference's
 * fit_mcmc_old (starmodel.py:955-969). */
int iso_sampler_reset(iso_ctx *ctx, iso_sampler *s);
/* enable != 0: every kept (thinned) ensemble of the following runs is added to the running sums below (off by default:
 * with thin = 1 the accumulation costs ~10 % of a step). */
int iso_sampler_set_moments(iso_ctx *ctx, iso_sampler *s, int enable);
/* Running sums over every kept (thinned) ensemble since the last reset, per chain: [n_chains, 2 ndim + 1] =
 * (sum x_d, sum x_d^2, number of samples).  h_moments (host copy; synchronises) and d_moments (the device array itself,
 * e.g. as the send buffer of iso_allgather_f64 in a multi-GPU catalog fit) may each be NULL. */
int iso_sampler_moments(iso_ctx *ctx, iso_sampler *s, double *h_moments, const double **d_moments);
int iso_sampler_destroy(iso_ctx *ctx, iso_sampler *s);

/* ------------------------------------------------------------------------------------------------
 * multi-GPU: rows are sharded across ranks; the only exchange is an all-gather of results
 * (SURVEY.md §8e).  NCCL is loaded at run time (libnccl.so.2).
 * ---------------------------------------------------------------------------------------------- */
int iso_nccl_unique_id(void *id128 /* 128 bytes */);
int iso_nccl_init(iso_ctx *ctx, const void *id128, int rank, int nranks);
int iso_nccl_destroy(iso_ctx *ctx);
/* d_recv[nranks * n] <- concat over ranks of d_send[n]; asynchronous on the compute stream */
int iso_allgather_f64(iso_ctx *ctx, const double *d_send, int64_t n, double *d_recv);

/* Fused lnpost + all-gather over NVLink peer memory (one process per GPU, at most 8 ranks of one node): the same
 * exchange as iso_lnpost_batch_device followed by iso_allgather_f64, but the lnpost kernel itself stores every row's
 * result into the receive buffer of every rank through CUDA-IPC peer mappings, so no collective is launched and the
 * transfer overlaps the evaluation.  Setup: every rank creates a group, exports 128 bytes, the launcher all-gathers
 * them (any byte transport), every rank connects.  A step returns a pointer to this rank's gathered buffer
 * [nranks * rows_per_rank] (rank-major; valid until the step after next: the buffers alternate by step parity). */
typedef struct iso_peer_group iso_peer_group;
int iso_peer_create(iso_ctx *ctx, int rank, int nranks, int64_t rows_per_rank, iso_peer_group **out);
int iso_peer_export(iso_ctx *ctx, iso_peer_group *group, void *handle128 /* 128 bytes */);
int iso_peer_connect(iso_ctx *ctx, iso_peer_group *group, const void *handles /* [nranks][128], rank order */);
int iso_lnpost_allgather_device(iso_ctx *ctx, const iso_grid *model_pack, const iso_grid *bc_pack,
                                const iso_models *models, const int32_t *d_model_of_row, const double *d_pars,
                                int64_t N, iso_peer_group *group, const double **d_gathered);
/* The completion wait of a step is bounded (default 10 s of wall-clock time per step): when a rank does not publish a
 * step in time the wait gives up, and the NEXT call on the group — or iso_peer_check, which first synchronises the
 * stream — fails with ISO_E_TIMEOUT naming the missing rank.  The error is sticky: the group must be destroyed. */
int iso_peer_set_timeout(iso_ctx *ctx, iso_peer_group *group, double seconds);
int iso_peer_check(iso_ctx *ctx, iso_peer_group *group);
int iso_peer_destroy(iso_ctx *ctx, iso_peer_group *group);

/* ------------------------------------------------------------------------------------------------
 * ONE ensemble sharded over the GPUs of a node: every rank owns a block of each half, moves its block
 * of the active half per half-step, and each proposal gathers its partner walker from the owning rank's memory through
 * CUDA-IPC peer mappings inside the evaluation kernel; no collective launch.  A run ends (and every kept step of a
 * thinned chain is preceded) by one replication of all blocks, so every rank holds the whole ensemble between runs and
 * iso_ensemble_state reads locally.  Same stretch move and Philox stream as
 * iso_sampler_*: the chain is independent of the number of ranks (bit for bit).  Setup as for iso_peer_*: create on
 * every rank with the same h_p0 / seed, export 128 bytes, all-gather them, connect.  Every rank must call
 * iso_ensemble_run with the same arguments; h_chain [n_steps / thin, n_walkers, ndim] and h_lnprob
 * [n_steps / thin, n_walkers] (either may be NULL) receive the kept ensembles on every rank that asks.  A rank that
 * stops participating turns into ISO_E_TIMEOUT on its peers (bounded wait, default 10 s per half-step).
 * ---------------------------------------------------------------------------------------------- */
typedef struct iso_ensemble iso_ensemble;
int iso_ensemble_create(iso_ctx *ctx, const iso_grid *model_pack, const iso_grid *bc_pack, const iso_models *models,
                        int n_walkers, const double *h_p0 /* [n_walkers, ndim] */, uint64_t seed, double stretch_a,
                        int rank, int nranks, iso_ensemble **out);
int iso_ensemble_export(iso_ctx *ctx, iso_ensemble *e, void *handle128 /* 128 bytes */);
int iso_ensemble_connect(iso_ctx *ctx, iso_ensemble *e, const void *handles /* [nranks][128], rank order */);
int iso_ensemble_set_timeout(iso_ctx *ctx, iso_ensemble *e, double seconds);
int iso_ensemble_run(iso_ctx *ctx, iso_ensemble *e, int n_steps, int thin, double *h_chain, double *h_lnprob);
/* current ensemble (either may be NULL), proposals accepted by THIS rank and proposals made per ensemble so far */
int iso_ensemble_state(iso_ctx *ctx, iso_ensemble *e, double *h_pos, double *h_lnprob, int64_t *n_accepted_local,
                       int64_t *n_proposed);
int iso_ensemble_destroy(iso_ctx *ctx, iso_ensemble *e);

#ifdef __cplusplus
}
#endif
#endif /* ISOCHRONES_B200_H */
