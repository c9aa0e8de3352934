// isochrones_b200 — fused lnprior + lnlike + lnpost batch kernel (the row evaluation lives in iso_lnpost_row.cuh).
//
// One launch turns rows of parameter vectors into log-posterior values.
//
// Data movement per row and star: the model-grid cell is gathered ONCE — 8 corners x one 64-byte
// node of the 8-column model pack (Teff, logg, feh, Mbol for interp_mag; age|mass and dt_deep|dm_deep for the
// EEP prior; nu_max, delta_nu) as 2 x LDG.256 per corner — and the BC cell as 16 corners x one 32-byte sector per
// chunk of 4 packed bands.  The reference gathers the same model cell twice (once in EEP_prior, once in
// interp_mag) and a third time for asteroseismology.
#include "iso_lnpost_kernel.cuh"

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
static IsoGaussDev make_gauss(double val, double unc)
{
    IsoGaussDev g;
    g.val = val;
    g.c = ISO_LOG_ONE_OVER_ROOT_2PI + log(unc);
    g.h = 1.0 / (unc * unc);
    return g;
}

static int convert_model(iso_ctx *ctx, const iso_model &s, IsoModelDev &d)
{
    memset(&d, 0, sizeof(d));
    ISO_REQUIRE(ctx, s.n_stars >= 1 && s.n_stars <= ISO_MAX_STARS, "iso_model: n_stars must be 1, 2 or 3");
    ISO_REQUIRE(ctx, s.n_bands >= 0 && s.n_bands <= ISO_MAX_BANDS, "iso_model: too many bands");
    d.n_stars = s.n_stars;
    d.eep_replaces_age = s.eep_replaces_age ? 1 : 0;
    int seen = 0;
    for (int j = 0; j < 5; j++) {
        ISO_REQUIRE(ctx, s.index_order[j] >= 0 && s.index_order[j] < 5, "iso_model: bad index_order");
        seen |= 1 << s.index_order[j];
        d.index_order[j] = s.index_order[j];
    }
    ISO_REQUIRE(ctx, seen == 31, "iso_model: index_order is not a permutation");
    {   // the two orders the reference defines (models.py:669, 696); the kernel hard-wires them per grid kind
        static const int track_order[5] = {2, 0, 1, 3, 4}, iso_order[5] = {1, 2, 0, 3, 4};
        const int *want = d.eep_replaces_age ? track_order : iso_order;
        for (int j = 0; j < 5; j++)
            if (d.index_order[j] != want[j])
                return iso_set_error(ctx, ISO_E_UNSUPPORTED, "iso_model: index_order must be (2,0,1,3,4) for track grids and "
                                                             "(1,2,0,3,4) for isochrone grids");
    }
    ISO_REQUIRE(ctx, !(d.eep_replaces_age && d.n_stars > 1),
                "iso_model: multiple stars need an isochrone grid (starmodel.py:1396-1397)");
    for (int i = 0; i < 3; i++) {
        d.spec[i] = make_gauss(s.spec_val[i], s.spec_unc[i]);
        if (s.spec_val[i] == s.spec_val[i]) d.spec_mask |= 1 << i;
    }
    for (int b = 0; b < s.n_bands; b++) {
        int c = s.band_col[b];
        ISO_REQUIRE(ctx, c >= 0 && c < ISO_MAX_BANDS, "iso_model: band_col out of range");
        ISO_REQUIRE(ctx, !(d.obs_mask & (1 << c)), "iso_model: two bands map to the same BC-pack column");
        d.obs_mask |= 1 << c;
        d.mag[c] = make_gauss(s.mag_val[b], s.mag_unc[b]);
    }
    d.has_plax = s.has_plax ? 1 : 0;
    d.has_nu_max = s.has_nu_max ? 1 : 0;
    d.has_delta_nu = s.has_delta_nu ? 1 : 0;
    d.plax = make_gauss(s.plax, s.plax_unc);
    d.nu_max = make_gauss(s.nu_max, s.nu_max_unc);
    d.delta_nu = make_gauss(s.delta_nu, s.delta_nu);   // the reference passes delta_nu as its own sigma (starmodel.py:1612)
    d.eep_lo = s.eep_lo;
    d.eep_hi = s.eep_hi;
    d.eep_norm = s.eep_norm;
    d.eep_inv_norm = 1.0 / s.eep_norm;
    d.eep_has_bounds = s.eep_has_bounds ? 1 : 0;
    const iso_prior *src[6] = {&s.eep_orig, &s.mass, &s.age, &s.feh, &s.distance, &s.AV};
    iso_prior *dst[6] = {&d.eep_orig, &d.mass, &d.age, &d.feh, &d.distance, &d.AV};
    for (int i = 0; i < 6; i++) {
        // the prior of the parameter EEP replaces is only reached through eep_orig
        bool used = !((i == 1 && !d.eep_replaces_age) || (i == 2 && d.eep_replaces_age));
        *dst[i] = *src[i];
        if (!used && !iso_prior_valid(*src[i])) {
            memset(dst[i], 0, sizeof(iso_prior));
            dst[i]->self.kind = ISO_PRIOR_FLAT;
            continue;
        }
        ISO_REQUIRE(ctx, iso_prior_valid(*src[i]), "iso_model: unsupported prior kind (no CPU fallback exists)");
        iso_prior_fill(dst[i]);
    }
    // the default BasicStarModel prior classes (any bounds / constants) -> the specialised kernel profile
    const iso_prior &other = d.eep_replaces_age ? d.mass : d.age;
    bool def = d.feh.self.kind == ISO_PRIOR_FEH && d.distance.self.kind == ISO_PRIOR_POWERLAW &&
               d.AV.self.kind == ISO_PRIOR_FLAT && (d.AV.self.flags & ISO_PF_BOUNDED);
    if (d.eep_replaces_age)
        def = def && iso_prior_is_chabrier_like(other) && d.eep_orig.self.kind == ISO_PRIOR_FLATLOG;
    else
        def = def && other.self.kind == ISO_PRIOR_FLATLOG && iso_prior_is_chabrier_like(d.eep_orig);
    d.profile_default = def ? 1 : 0;
    d.eep_lnc[0] = d.eep_lnc[1] = 0.0;
    if (def && !d.eep_replaces_age) {
        // pdf(mass) / eep_norm = comp_i._pdf(mass) / comp_i._norm / norms[i] / _norm / eep_norm; the x-independent part:
        //   LogNormal (priors.py:272-275): 1 / (sqrt(2 pi) s scale);   PowerLaw (priors.py:469-471): (1 + alpha) / (hi^(1+alpha) - lo^(1+alpha))
        const iso_prior &op = d.eep_orig;
        const double common = op.inv_norm * d.eep_inv_norm;
        d.eep_lnc[0] = log(op.comp[0].k[1] * op.comp[0].k[2] * op.inv_norms[0] * common);
        d.eep_lnc[1] = log(op.comp[1].k[2] * op.inv_norms[1] * common);
    }
    return ISO_OK;
}

// axis tables that go to shared memory: every non-closed-form axis of the two grids
int iso_row_grids_fill(iso_ctx *ctx, const iso_grid *mp, const iso_grid *bp, IsoRowGrids *out, size_t *smem_bytes)
{
    // derived layout of the model pack the row kernels gather (built once, cached in the handle)
    int prc = ISO_MODEL_LAYOUT == ISO_LAYOUT_PAIR96 ? iso_grid_pair_pack(ctx, mp)
              : ISO_MODEL_LAYOUT == ISO_LAYOUT_NODE48 ? iso_grid_n48_pack(ctx, mp) : ISO_OK;
    if (prc != ISO_OK) return prc;
    out->mg = mp->dev;
    out->bg = bp->dev;
    int total = 0;
    const iso_grid *gr[2] = {mp, bp};
    for (int g = 0; g < 2; g++)
        for (int d = 0; d < ISO_MAX_DIM; d++) {
            out->smem_axis_off[g][d] = -1;   // closed-form axes need no table; iso_axis_locate never dereferences it
            if (d >= gr[g]->dev.ndim || gr[g]->dev.ax[d].arith) continue;
            out->smem_axis_off[g][d] = total;
            total += gr[g]->dev.ax[d].n;
        }
    out->smem_nodes = total;
    *smem_bytes = sizeof(double2) * (size_t)(total > 0 ? total : 1);
    ISO_REQUIRE(ctx, *smem_bytes <= 160 * 1024, "lnpost: axis tables do not fit in shared memory");
    return ISO_OK;
}

// unit-cube rows (BasicStarModel.mnest_prior fused into the evaluation; rng: cube points drawn on the device)
struct IsoCubeSpec {
    const double *lo, *hi;   // host arrays [ndim]
    double *d_pars_out;      // mapped parameters (device or page-locked host memory; NULL: not wanted)
    unsigned long long seed;
    long long row0;
    bool rng;
};

static int lnpost_launch(iso_ctx *ctx, cudaStream_t st, const iso_grid *mp, const iso_grid *bp, const iso_models *models,
                         const int32_t *d_model_of_row, const double *d_pars, int64_t N, double *d_lnpost, double *d_lnprior,
                         double *d_lnlike, const IsoPeerTargets *peers = nullptr, const IsoCubeSpec *cube = nullptr)
{
    IsoLnpostParams P;
    size_t smem = 0;
    int rc = iso_row_grids_fill(ctx, mp, bp, &P.G, &smem);
    if (rc != ISO_OK) return rc;
    P.model = models->h_first;
    IsoLnpostArgs &a = P.a;
    a.models = models->d_models;
    a.model_of_row = d_model_of_row;
    a.pars = d_pars;
    a.lnpost = d_lnpost;
    a.lnprior = d_lnprior;
    a.lnlike = d_lnlike;
    a.N = N;
    a.n_peers = 0;
    a.peer_off = 0;
    a.peer_rank = 0;
    a.peer_step = 0;
    a.peer_done = nullptr;
    a.peer_own_flags = nullptr;
    a.peer_timeout_ns = 0;
    a.peer_err = nullptr;
    a.claim = ctx->d_claim + ISO_CLAIM_STRIDE * (st == ctx->copy_stream[0] ? 1 : st == ctx->copy_stream[1] ? 2 : 0);
    for (int q = 0; q < ISO_MAX_PEERS; q++) {
        a.peer_out[q] = nullptr;
        a.peer_flags[q] = nullptr;
    }
    if (peers) {
        a.n_peers = peers->n;
        a.peer_off = peers->offset;
        a.peer_rank = peers->rank;
        a.peer_step = peers->step;
        a.peer_done = peers->done;
        a.peer_own_flags = peers->own_flags;
        a.peer_timeout_ns = peers->timeout_ns;
        a.peer_err = peers->err;
        for (int q = 0; q < peers->n; q++) {
            a.peer_out[q] = peers->out[q];
            a.peer_flags[q] = peers->flags[q];
        }
    }
    IsoLnpostFlags f;
    f.catalog = d_model_of_row != nullptr;
    f.profile_default = models->profile_default;
    f.track = models->track;
    f.peer = peers != nullptr;
#ifndef ISO_SEQ_MAX_CHUNKS
#define ISO_SEQ_MAX_CHUNKS 3
#endif
    // multi-star models: the star-sequential kernels for BC packs of 1..3 chunks of 4 bands, chunk-major beyond
    f.seq = bp->dev.ncols / 4 <= ISO_SEQ_MAX_CHUNKS ? bp->dev.ncols / 4 : 0;
    f.cube = cube != nullptr;
    a.pars_out = nullptr;
    a.cube_seed = 0;
    a.cube_row0 = 0;
    a.cube_rng = 0;
    for (int j = 0; j < ISO_MAX_STARS + 4; j++) a.cube_lo[j] = a.cube_w[j] = 0.0;
    if (cube) {
        const int ndim = 4 + models->n_stars;
        for (int j = 0; j < ndim; j++) {
            a.cube_lo[j] = cube->lo[j];
            a.cube_w[j] = cube->hi[j] - cube->lo[j];
        }
        a.pars_out = cube->d_pars_out;
        a.cube_seed = cube->seed;
        a.cube_row0 = cube->row0;
        a.cube_rng = cube->rng ? 1 : 0;
    }
    switch (models->n_stars) {
    case 1: rc = iso_lnpost_dispatch_1(ctx, st, P, smem, f); break;
    case 2: rc = iso_lnpost_dispatch_2(ctx, st, P, smem, f); break;
    case 3: rc = iso_lnpost_dispatch_3(ctx, st, P, smem, f); break;
    default: return iso_set_error(ctx, ISO_E_INVALID, "lnpost: bad n_stars");
    }
    if (rc != ISO_OK) return rc;
    ctx->launches++;
    ISO_CUDA(ctx, cudaGetLastError());
    return ISO_OK;
}

int iso_check_lnpost_handles(iso_ctx *ctx, const iso_grid *mp, const iso_grid *bp, const iso_models *models)
{
    ISO_REQUIRE(ctx, mp && bp && models, "lnpost: NULL handle");
    ISO_REQUIRE(ctx, mp->device == ctx->device && bp->device == ctx->device && models->device == ctx->device,
                "lnpost: handle belongs to another device");
    ISO_REQUIRE(ctx, mp->dev.ndim == 3 && mp->dev.ncols == ISO_MP_NCOLS,
                "lnpost: model_pack must be a 3-D grid with the 8 ISO_MP_* columns (iso_grid_repack)");
    ISO_REQUIRE(ctx, bp->dev.ndim == 4 && bp->dev.ncols % 4 == 0 && bp->dev.ncols <= ISO_MAX_BANDS,
                "lnpost: bc_pack must be a 4-D grid with 4, 8, 12 or 16 columns (iso_grid_repack)");
    ISO_REQUIRE(ctx, models->max_col < bp->dev.ncols, "lnpost: a model observes a band column the BC pack does not have");
    return ISO_OK;
}

struct LnpostUser {
    const iso_grid *mp, *bp;
    const iso_models *models;
    bool catalog, want_prior, want_like;
};

static int lnpost_pipe_launch(iso_ctx *ctx, cudaStream_t st, void *const *d, int64_t row0, int64_t n, void *user)
{
    (void)row0;
    LnpostUser *u = (LnpostUser *)user;
    return lnpost_launch(ctx, st, u->mp, u->bp, u->models, u->catalog ? (const int32_t *)d[1] : nullptr, (const double *)d[0],
                         n, (double *)d[2], u->want_prior ? (double *)d[3] : nullptr, u->want_like ? (double *)d[4] : nullptr);
}

// BasicStarModel.mnest_prior (starmodel.py:1637-1640): cube[i] = (hi - lo) * cube[i] + lo, unfused (bit-exact).  The
// bounds travel in the kernel parameter block: no device allocation or copy per call.
#define ISO_MNEST_MAX_DIM 16
struct IsoMnestBounds {
    double lo[ISO_MNEST_MAX_DIM], w[ISO_MNEST_MAX_DIM];
};

__global__ void iso_mnest_prior_kernel(double *cube, const __grid_constant__ IsoMnestBounds b, int ndim, long long total)
{
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(t % ndim);
        double lo = b.lo[0], w = b.w[0];
#pragma unroll
        for (int q = 1; q < ISO_MNEST_MAX_DIM; q++)   // constant-bank operands: no dynamic indexing of the parameter block
            if (q == j) {
                lo = b.lo[q];
                w = b.w[q];
            }
        cube[t] = __dadd_rn(__dmul_rn(w, cube[t]), lo);
    }
}

struct MnestUser {
    IsoMnestBounds b;
    int ndim;
};

static int mnest_launch(iso_ctx *ctx, cudaStream_t st, void *const *d, int64_t row0, int64_t n, void *user)
{
    (void)row0;
    MnestUser *u = (MnestUser *)user;
    long long total = n * u->ndim;
    int blocks = (int)((total + 255) / 256 < ctx->prop.multiProcessorCount * 8 ? (total + 255) / 256
                                                                                : ctx->prop.multiProcessorCount * 8);
    // in place: the output array aliases the input rows (d[1] receives a device-side copy below)
    // (the buffers are device memory, or page-locked host memory on the small-call path: cudaMemcpyDefault)
    if (d[1] != d[0]) ISO_CUDA(ctx, cudaMemcpyAsync(d[1], d[0], (size_t)total * 8, cudaMemcpyDefault, st));
    iso_mnest_prior_kernel<<<blocks, 256, 0, st>>>((double *)d[1], u->b, u->ndim, total);
    ctx->launches++;
    ISO_CUDA(ctx, cudaGetLastError());
    return ISO_OK;
}

// unit-cube rows through the fused kernel (host-pointer pipeline): d[0] cube in (absent when drawn on the device),
// d[1] mapped parameters out, d[2..4] lnpost / lnprior / lnlike
struct CubeUser {
    const iso_grid *mp, *bp;
    const iso_models *models;
    const double *lo, *hi;
    unsigned long long seed;
    long long row0;
    bool rng, want_pars, want_prior, want_like;
};

static int cube_pipe_launch(iso_ctx *ctx, cudaStream_t st, void *const *d, int64_t row0, int64_t n, void *user)
{
    CubeUser *u = (CubeUser *)user;
    IsoCubeSpec c{u->lo, u->hi, u->want_pars ? (double *)d[1] : nullptr, u->seed, u->row0 + row0, u->rng};
    return lnpost_launch(ctx, st, u->mp, u->bp, u->models, nullptr, (const double *)d[0], n, (double *)d[2],
                         u->want_prior ? (double *)d[3] : nullptr, u->want_like ? (double *)d[4] : nullptr, nullptr, &c);
}

static int check_cube_args(iso_ctx *ctx, const iso_models *models, const double *lo, const double *hi)
{
    ISO_REQUIRE(ctx, lo && hi, "unit-cube rows: NULL bounds");
    ISO_REQUIRE(ctx, models->n_models == 1, "unit-cube rows: one staged model (the bounds are the model's)");
    for (int j = 0; j < 4 + models->n_stars; j++)
        ISO_REQUIRE(ctx, lo[j] == lo[j] && hi[j] == hi[j], "unit-cube rows: NaN bound");
    return ISO_OK;
}

extern "C" {

int iso_models_stage(iso_ctx *ctx, const iso_model *h_models, int n_models, iso_models **out)
{
    if (!ctx) return iso_set_error(nullptr, ISO_E_INVALID, "iso_models_stage: ctx is NULL");
    ISO_REQUIRE(ctx, h_models && out && n_models >= 1, "iso_models_stage: bad argument");
    *out = nullptr;
    std::vector<IsoModelDev> dev((size_t)n_models);
    int max_col = -1;
    bool seismo = false, all_default = true;
    for (int i = 0; i < n_models; i++) {
        int rc = convert_model(ctx, h_models[i], dev[i]);
        if (rc != ISO_OK) return rc;
        ISO_REQUIRE(ctx, dev[i].n_stars == dev[0].n_stars, "iso_models_stage: all models must share n_stars");
        ISO_REQUIRE(ctx, dev[i].eep_replaces_age == dev[0].eep_replaces_age, "iso_models_stage: all models must share the grid kind");
        all_default = all_default && dev[i].profile_default;
        for (int c = 0; c < ISO_MAX_BANDS; c++)
            if (dev[i].obs_mask & (1 << c)) max_col = c > max_col ? c : max_col;
        seismo = seismo || dev[i].has_nu_max;
    }
    IsoDeviceGuard guard(ctx->device);
    iso_models *m = new iso_models();
    m->n_models = n_models;
    m->n_stars = dev[0].n_stars;
    m->device = ctx->device;
    m->max_col = max_col;
    m->needs_seismo = seismo;
    m->h_first = dev[0];
    m->profile_default = all_default;
    m->track = dev[0].eep_replaces_age != 0;
    cudaError_t e = cudaMalloc(&m->d_models, sizeof(IsoModelDev) * (size_t)n_models);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(m->d_models, dev.data(), sizeof(IsoModelDev) * (size_t)n_models, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        if (m->d_models) cudaFree(m->d_models);
        delete m;
        return iso_check_cuda(ctx, e, "iso_models_stage");
    }
    *out = m;
    return ISO_OK;
}

int iso_models_destroy(iso_ctx *ctx, iso_models *models)
{
    if (!models) return ISO_OK;
    IsoDeviceGuard guard(models->device);
    if (ctx) cudaStreamSynchronize(ctx->stream);
    if (models->d_models) cudaFree(models->d_models);
    delete models;
    return ISO_OK;
}

int iso_lnpost_batch_device(iso_ctx *ctx, const iso_grid *model_pack, const iso_grid *bc_pack, const iso_models *models,
                            const int32_t *d_model_of_row, const double *d_pars, int64_t N, double *d_lnpost,
                            double *d_lnprior, double *d_lnlike)
{
    if (!ctx) return iso_set_error(nullptr, ISO_E_INVALID, "iso_lnpost_batch_device: ctx is NULL");
    int rc = iso_check_lnpost_handles(ctx, model_pack, bc_pack, models);
    if (rc != ISO_OK) return rc;
    ISO_REQUIRE(ctx, N >= 0, "lnpost: negative N");
    if (N == 0) return ISO_OK;
    ISO_REQUIRE(ctx, d_pars && d_lnpost, "lnpost: NULL buffer");
    ISO_REQUIRE(ctx, d_model_of_row || models->n_models == 1, "lnpost: several models staged but no model_of_row given");
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    IsoDeviceGuard guard(ctx->device);
    return lnpost_launch(ctx, ctx->stream, model_pack, bc_pack, models, d_model_of_row, d_pars, N, d_lnpost, d_lnprior,
                         d_lnlike);
}

}  // extern "C"

// fused lnpost + all-gather launch (called by iso_peer.cu with the peer mappings of one exchange step)
int iso_lnpost_launch_peers(iso_ctx *ctx, const iso_grid *model_pack, const iso_grid *bc_pack, const iso_models *models,
                            const int32_t *d_model_of_row, const double *d_pars, int64_t N, const IsoPeerTargets *peers)
{
    int rc = iso_check_lnpost_handles(ctx, model_pack, bc_pack, models);
    if (rc != ISO_OK) return rc;
    ISO_REQUIRE(ctx, N >= 0 && peers && peers->n >= 1 && peers->n <= ISO_MAX_PEERS, "fused all-gather: bad argument");
    if (N == 0) return ISO_OK;
    ISO_REQUIRE(ctx, d_pars, "lnpost: NULL buffer");
    ISO_REQUIRE(ctx, d_model_of_row || models->n_models == 1, "lnpost: several models staged but no model_of_row given");
    return lnpost_launch(ctx, ctx->stream, model_pack, bc_pack, models, d_model_of_row, d_pars, N, nullptr, nullptr, nullptr,
                         peers);
}

extern "C" {

int iso_lnpost_batch(iso_ctx *ctx, const iso_grid *model_pack, const iso_grid *bc_pack, const iso_models *models,
                     const int32_t *h_model_of_row, const double *h_pars, int64_t N, double *h_lnpost, double *h_lnprior,
                     double *h_lnlike)
{
    if (!ctx) return iso_set_error(nullptr, ISO_E_INVALID, "iso_lnpost_batch: ctx is NULL");
    int rc = iso_check_lnpost_handles(ctx, model_pack, bc_pack, models);
    if (rc != ISO_OK) return rc;
    ISO_REQUIRE(ctx, N >= 0, "lnpost: negative N");
    if (N == 0) return ISO_OK;
    ISO_REQUIRE(ctx, h_pars && h_lnpost, "lnpost: NULL buffer");
    ISO_REQUIRE(ctx, h_model_of_row || models->n_models == 1, "lnpost: several models staged but no model_of_row given");
    if (h_model_of_row)
        for (int64_t i = 0; i < N; i++)
            ISO_REQUIRE(ctx, h_model_of_row[i] >= 0 && h_model_of_row[i] < models->n_models, "lnpost: model_of_row out of range");
    int ndim = 4 + models->n_stars;
    IsoPipeArray arr[5];
    arr[0] = IsoPipeArray{h_pars, nullptr, (int64_t)8 * ndim};
    arr[1] = IsoPipeArray{h_model_of_row, nullptr, 4};
    arr[2] = IsoPipeArray{nullptr, h_lnpost, 8};
    arr[3] = IsoPipeArray{nullptr, h_lnprior, 8};
    arr[4] = IsoPipeArray{nullptr, h_lnlike, 8};
    LnpostUser u{model_pack, bc_pack, models, h_model_of_row != nullptr, h_lnprior != nullptr, h_lnlike != nullptr};
    return iso_run_pipeline(ctx, N, arr, 5, lnpost_pipe_launch, &u);
}

int iso_mnest_prior(iso_ctx *ctx, const double *h_lo, const double *h_hi, int ndim, double *h_cube, int64_t N)
{
    if (!ctx) return iso_set_error(nullptr, ISO_E_INVALID, "iso_mnest_prior: ctx is NULL");
    ISO_REQUIRE(ctx, h_lo && h_hi && ndim >= 1 && ndim <= ISO_MNEST_MAX_DIM && N >= 0,
                "iso_mnest_prior: bad argument (at most 16 parameters)");
    if (N == 0) return ISO_OK;
    ISO_REQUIRE(ctx, h_cube, "iso_mnest_prior: cube is NULL");
    MnestUser u;
    memset(&u, 0, sizeof(u));
    u.ndim = ndim;
    for (int j = 0; j < ndim; j++) {
        u.b.lo[j] = h_lo[j];
        u.b.w[j] = h_hi[j] - h_lo[j];
    }
    // input and output are the same host array: two pipeline arrays over it (read, then write back); a call of at most
    // 512 rows (MultiNest transforms one live point at a time) is one launch on page-locked memory + one synchronisation
    IsoPipeArray arr[2];
    arr[0] = IsoPipeArray{h_cube, nullptr, (int64_t)8 * ndim};
    arr[1] = IsoPipeArray{nullptr, h_cube, (int64_t)8 * ndim};
    return iso_run_pipeline(ctx, N, arr, 2, mnest_launch, &u);
}

int iso_mnest_lnpost_batch(iso_ctx *ctx, const iso_grid *model_pack, const iso_grid *bc_pack, const iso_models *models,
                           const double *h_lo, const double *h_hi, double *h_cube, int64_t N, double *h_lnpost,
                           double *h_lnprior, double *h_lnlike)
{
    if (!ctx) return iso_set_error(nullptr, ISO_E_INVALID, "iso_mnest_lnpost_batch: ctx is NULL");
    int rc = iso_check_lnpost_handles(ctx, model_pack, bc_pack, models);
    if (rc != ISO_OK) return rc;
    ISO_REQUIRE(ctx, N >= 0, "iso_mnest_lnpost_batch: negative N");
    if (N == 0) return ISO_OK;
    ISO_REQUIRE(ctx, h_cube && h_lnpost, "iso_mnest_lnpost_batch: NULL buffer");
    rc = check_cube_args(ctx, models, h_lo, h_hi);
    if (rc != ISO_OK) return rc;
    const int ndim = 4 + models->n_stars;
    IsoPipeArray arr[5];
    arr[0] = IsoPipeArray{h_cube, nullptr, (int64_t)8 * ndim};
    arr[1] = IsoPipeArray{nullptr, h_cube, (int64_t)8 * ndim};   // mapped in place, as mnest_prior leaves the cube
    arr[2] = IsoPipeArray{nullptr, h_lnpost, 8};
    arr[3] = IsoPipeArray{nullptr, h_lnprior, 8};
    arr[4] = IsoPipeArray{nullptr, h_lnlike, 8};
    CubeUser u{model_pack, bc_pack, models, h_lo, h_hi, 0ULL, 0LL, false, true, h_lnprior != nullptr, h_lnlike != nullptr};
    return iso_run_pipeline(ctx, N, arr, 5, cube_pipe_launch, &u);
}

int iso_lnpost_prior_draws(iso_ctx *ctx, const iso_grid *model_pack, const iso_grid *bc_pack, const iso_models *models,
                           const double *h_lo, const double *h_hi, uint64_t seed, int64_t row0, int64_t N, double *h_pars,
                           double *h_lnpost)
{
    if (!ctx) return iso_set_error(nullptr, ISO_E_INVALID, "iso_lnpost_prior_draws: ctx is NULL");
    int rc = iso_check_lnpost_handles(ctx, model_pack, bc_pack, models);
    if (rc != ISO_OK) return rc;
    ISO_REQUIRE(ctx, N >= 0 && row0 >= 0, "iso_lnpost_prior_draws: bad argument");
    if (N == 0) return ISO_OK;
    ISO_REQUIRE(ctx, h_lnpost, "iso_lnpost_prior_draws: NULL buffer");
    rc = check_cube_args(ctx, models, h_lo, h_hi);
    if (rc != ISO_OK) return rc;
    const int ndim = 4 + models->n_stars;
    IsoPipeArray arr[5];
    arr[0] = IsoPipeArray{nullptr, nullptr, (int64_t)8 * ndim};   // no input rows: the cube points are drawn on the device
    arr[1] = IsoPipeArray{nullptr, h_pars, (int64_t)8 * ndim};
    arr[2] = IsoPipeArray{nullptr, h_lnpost, 8};
    arr[3] = IsoPipeArray{nullptr, nullptr, 8};
    arr[4] = IsoPipeArray{nullptr, nullptr, 8};
    CubeUser u{model_pack, bc_pack, models, h_lo, h_hi, (unsigned long long)seed, (long long)row0, true, h_pars != nullptr, false, false};
    return iso_run_pipeline(ctx, N, arr, 5, cube_pipe_launch, &u);
}

int iso_lnpost_cube_device(iso_ctx *ctx, const iso_grid *model_pack, const iso_grid *bc_pack, const iso_models *models,
                           const double *h_lo, const double *h_hi, const double *d_cube, int rng, uint64_t seed, int64_t row0,
                           int64_t N, double *d_pars, double *d_lnpost)
{
    if (!ctx) return iso_set_error(nullptr, ISO_E_INVALID, "iso_lnpost_cube_device: ctx is NULL");
    int rc = iso_check_lnpost_handles(ctx, model_pack, bc_pack, models);
    if (rc != ISO_OK) return rc;
    ISO_REQUIRE(ctx, N >= 0 && row0 >= 0, "iso_lnpost_cube_device: bad argument");
    if (N == 0) return ISO_OK;
    ISO_REQUIRE(ctx, d_lnpost && (rng || d_cube), "iso_lnpost_cube_device: NULL buffer");
    rc = check_cube_args(ctx, models, h_lo, h_hi);
    if (rc != ISO_OK) return rc;
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    IsoDeviceGuard guard(ctx->device);
    IsoCubeSpec c{h_lo, h_hi, d_pars, (unsigned long long)seed, (long long)row0, rng != 0};
    return lnpost_launch(ctx, ctx->stream, model_pack, bc_pack, models, nullptr, d_cube, N, d_lnpost, nullptr, nullptr, nullptr, &c);
}

}  // extern "C"
