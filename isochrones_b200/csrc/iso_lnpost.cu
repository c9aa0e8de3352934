// isochrones_b200 — fused lnprior + lnlike + lnpost batch kernel (the row evaluation lives in iso_lnpost_row.cuh).
//
// One launch turns rows of parameter vectors into log-posterior values.
//
// Data movement per row and star: the model-grid cell is gathered ONCE — 8 corners x one 64-byte
// node of the 8-column model pack (Teff, logg, feh, Mbol for interp_mag; age|mass and dt_deep|dm_deep for the
// EEP prior; nu_max, delta_nu) as 2 x LDG.256 per corner — and the BC cell as 16 corners x one 32-byte sector per
// chunk of 4 packed bands.  The reference gathers the same model cell twice (once in EEP_prior, once in
// interp_mag) and a third time for asteroseismology.
#include "iso_lnpost_row.cuh"

#ifndef ISO_LNPOST_THREADS
#define ISO_LNPOST_THREADS 256
#endif
#ifndef ISO_LNPOST_MIN_BLOCKS
#define ISO_LNPOST_MIN_BLOCKS 2
#endif
#ifndef ISO_LNPOST_MIN_BLOCKS_MULTI
#define ISO_LNPOST_MIN_BLOCKS_MULTI 2   // binary / triple models: 128 registers + a small spill beats 1 CTA/SM at 255
#endif
#ifndef ISO_LNPOST_PREFETCH
#define ISO_LNPOST_PREFETCH 1
#endif
#ifndef ISO_LNPOST_BLOCKS_PER_SM
#define ISO_LNPOST_BLOCKS_PER_SM 2   // persistent grid: exactly the CTAs that are resident (2 per SM), rows grid-strided
#endif

struct IsoLnpostArgs {
    const IsoModelDev *models;
    const int *model_of_row;   // catalog mode only
    const double *pars;        // [N, 4 + n_stars] row-major
    double *lnpost, *lnprior, *lnlike;   // [N]; lnprior / lnlike may be NULL
    long long N;
    // fused all-gather (PEER kernels, iso_peer.cu): row i of this rank is stored at peer_out[r][peer_off + i] in the
    // receive buffer of EVERY rank r (its own included) — plain stores over NVLink peer mappings
    double *peer_out[ISO_MAX_PEERS];
    long long peer_off;
    int n_peers;
    int peer_rank;
    // completion signal of the fused all-gather: the last CTA to finish publishes `peer_step` in this rank's slot of
    // every rank's flag array (release, system scope); peer_done counts finished CTAs and is reset by that CTA
    unsigned long long *peer_flags[ISO_MAX_PEERS];
    unsigned long long peer_step;
    unsigned *peer_done;
    unsigned long long *claim;   // dynamically scheduled kernels: [0] next unclaimed row beyond the first pass, [1] finished CTAs
};

// Everything the kernel reads besides the grids and the rows travels in the kernel parameter block (constant
// bank): the grid descriptors, the buffer pointers and — outside catalog mode — the star model itself, so that
// observation values and prior constants are constant-bank operands instead of memory loads.
struct IsoLnpostParams {
    IsoRowGrids G;
    IsoLnpostArgs a;
    IsoModelDev model;   // the single model (unused in catalog mode)
};

#ifndef ISO_LNPOST_DYN
#define ISO_LNPOST_DYN 1
#endif

// Row scheduling of the persistent grid:
//   DYN = 0 — static grid stride, the next row's parameters prefetched one iteration ahead;
//   DYN = 1 — the first pass is the static one, after it every warp claims 32-row chunks from an atomic counter
//             (a.claim[0]) so that the grid drains together whatever the rows cost (rows rejected by the grid-free
//             priors cost a tenth of a full row; rows whose gathers miss the L2 several times a cached one).  Batches
//             of at most one pass never touch the counters.  Round 2, B200, ms per 1e6 rows static -> dynamic:
//             grid-wide 0.175 -> 0.165, prior-like 0.139 -> 0.134, binary 0.293 -> 0.274, posterior-like 0.121 = 0.121;
//   DYN = 2 — the same with the claim and the row load issued one iteration ahead.
// The last CTA to finish resets the counters (a.claim[1] counts finished CTAs), so a launch finds them zero.
// SEQ: star-sequential evaluation of multi-star models whose BC pack is a single 4-band chunk (iso_lnpost_row.cuh).
template <int NSTARS, bool CATALOG, int PROFILE, bool TRACK, bool PEER = false, bool SEQ = false,
          int LAYOUT = ISO_MODEL_LAYOUT, int DYN = ISO_LNPOST_DYN>
__global__ void __launch_bounds__(ISO_LNPOST_THREADS, NSTARS == 1 ? ISO_LNPOST_MIN_BLOCKS : ISO_LNPOST_MIN_BLOCKS_MULTI)
iso_lnpost_kernel(const __grid_constant__ IsoLnpostParams P)
{
    constexpr int NDIMP = NSTARS + 4;
    const IsoLnpostArgs &a = P.a;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2 *s_nodes = reinterpret_cast<double2 *>(smem_raw);
    iso_stage_axis_tables(P.G, s_nodes);

    const bool want_prior = a.lnprior != nullptr, want_like = a.lnlike != nullptr;
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    auto eval_row = [&](long long row, const double (&p)[NDIMP]) {
        const IsoModelDev &m = CATALOG ? a.models[a.model_of_row[row]] : P.model;
        const IsoRowResult r = iso_lnpost_row<NSTARS, PROFILE, TRACK, LAYOUT, SEQ>(P.G, s_nodes, m, p, want_prior, want_like);
        if (want_prior) a.lnprior[row] = r.lnprior;
        if (want_like) a.lnlike[row] = r.lnlike;
        if (PEER) {
#pragma unroll
            for (int q = 0; q < ISO_MAX_PEERS; q++)
                if (q < a.n_peers) a.peer_out[q][a.peer_off + row] = r.lnpost;
        } else {
            a.lnpost[row] = r.lnpost;
        }
    };
    if (DYN == 0) {
#if ISO_LNPOST_PREFETCH
        // software pipelining of the row stream: the next row's parameters are requested before this row is evaluated,
        // so their HBM latency hides behind ~2000 instructions of work
        double pn[NDIMP];
        if (i < a.N) {
#pragma unroll
            for (int j = 0; j < NDIMP; j++) pn[j] = a.pars[i * NDIMP + j];
        }
#endif
        for (; i < a.N; i += stride) {
            double p[NDIMP];
#if ISO_LNPOST_PREFETCH
#pragma unroll
            for (int j = 0; j < NDIMP; j++) p[j] = pn[j];
            if (i + stride < a.N) {
#pragma unroll
                for (int j = 0; j < NDIMP; j++) pn[j] = a.pars[(i + stride) * NDIMP + j];
            }
#else
#pragma unroll
            for (int j = 0; j < NDIMP; j++) p[j] = a.pars[i * NDIMP + j];
#endif
            eval_row(i, p);
        }
    } else if (DYN == 1) {
        const int lane = threadIdx.x & 31;
        long long base = i - lane;   // warp-uniform: the static first pass
        while (base < a.N) {
            const long long row = base + lane;
            if (row < a.N) {
                double p[NDIMP];
#pragma unroll
                for (int j = 0; j < NDIMP; j++) p[j] = a.pars[row * NDIMP + j];
                eval_row(row, p);
            }
            if (a.N <= stride) break;   // single pass: nothing to claim, the counters stay untouched
            unsigned long long got = 0;
            if (lane == 0) got = atomicAdd(a.claim, 32ULL);
            base = stride + (long long)__shfl_sync(0xffffffffu, got, 0);
        }
    } else {
        const int lane = threadIdx.x & 31;
        long long base = i - lane;
        double pn[NDIMP];
        if (base + lane < a.N) {
#pragma unroll
            for (int j = 0; j < NDIMP; j++) pn[j] = a.pars[(base + lane) * NDIMP + j];
        }
        unsigned long long got = 0;
        if (lane == 0) got = atomicAdd(a.claim, 32ULL);
        while (base < a.N) {
            const long long row = base + lane;
            double p[NDIMP];
#pragma unroll
            for (int j = 0; j < NDIMP; j++) p[j] = pn[j];
            const long long nbase = stride + (long long)__shfl_sync(0xffffffffu, got, 0);
            if (nbase + lane < a.N) {
#pragma unroll
                for (int j = 0; j < NDIMP; j++) pn[j] = a.pars[(nbase + lane) * NDIMP + j];
            }
            if (lane == 0 && nbase < a.N) got = atomicAdd(a.claim, 32ULL);
            if (row < a.N) eval_row(row, p);
            base = nbase;
        }
    }
    if (DYN != 0 && !PEER && (DYN != 1 || a.N > stride)) {
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned long long ticket = atomicAdd(a.claim + 1, 1ULL);
            if (ticket == gridDim.x - 1) {   // every CTA has made its last claim: leave the counters zero
                a.claim[0] = 0;
                a.claim[1] = 0;
            }
        }
    }
    if (PEER) {
        // every thread's peer stores are ordered before its arrival; the last CTA then raises the flags
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned ticket = atomicAdd(a.peer_done, 1u);
            if (ticket == gridDim.x - 1) {
                *a.peer_done = 0;
                if (DYN != 0 && (DYN != 1 || a.N > stride)) {
                    a.claim[0] = 0;
                    a.claim[1] = 0;
                }
                __threadfence_system();
#pragma unroll
                for (int q = 0; q < ISO_MAX_PEERS; q++)
                    if (q < a.n_peers)
                        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(a.peer_flags[q] + a.peer_rank), "l"(a.peer_step)
                                     : "memory");
            }
        }
    }
}

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

static int lnpost_launch(iso_ctx *ctx, cudaStream_t st, const iso_grid *mp, const iso_grid *bp, const iso_models *models,
                         const int32_t *d_model_of_row, const double *d_pars, int64_t N, double *d_lnpost, double *d_lnprior,
                         double *d_lnlike, const IsoPeerTargets *peers = nullptr)
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
        for (int q = 0; q < peers->n; q++) {
            a.peer_out[q] = peers->out[q];
            a.peer_flags[q] = peers->flags[q];
        }
    }
    const bool peer = peers != nullptr;
    const bool catalog = d_model_of_row != nullptr;
    int64_t want = (N + ISO_LNPOST_THREADS - 1) / ISO_LNPOST_THREADS;
    int64_t cap = (int64_t)ctx->prop.multiProcessorCount * ISO_LNPOST_BLOCKS_PER_SM;
    int blocks = (int)(want < cap ? want : cap);
    if (blocks < 1) blocks = 1;
#define ISO_LAUNCH6(NS, CAT, PROF, TRK, PEER, SEQ)                                                                       \
    do {                                                                                                                 \
        if (smem > 48 * 1024)                                                                                            \
            ISO_CUDA(ctx, cudaFuncSetAttribute(iso_lnpost_kernel<NS, CAT, PROF, TRK, PEER, SEQ>,                         \
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                 \
        iso_lnpost_kernel<NS, CAT, PROF, TRK, PEER, SEQ><<<blocks, ISO_LNPOST_THREADS, smem, st>>>(P);                   \
    } while (0)
#define ISO_LAUNCH5(NS, CAT, PROF, TRK, PEER)                                                                            \
    do {                                                                                                                 \
        if (NS > 1 && seq) ISO_LAUNCH6(NS, CAT, PROF, TRK, PEER, (NS > 1));                                              \
        else ISO_LAUNCH6(NS, CAT, PROF, TRK, PEER, false);                                                               \
    } while (0)
#define ISO_LAUNCH4(NS, CAT, PROF, TRK)                                                                                  \
    do {                                                                                                                 \
        if (peer) ISO_LAUNCH5(NS, CAT, PROF, TRK, true);                                                                 \
        else ISO_LAUNCH5(NS, CAT, PROF, TRK, false);                                                                     \
    } while (0)
#define ISO_LAUNCH2(NS, TRK)                                                           \
    do {                                                                               \
        if (catalog) {                                                                 \
            if (def) ISO_LAUNCH4(NS, true, ISO_PROFILE_DEFAULT, TRK);                  \
            else ISO_LAUNCH4(NS, true, ISO_PROFILE_GENERIC, TRK);                      \
        } else {                                                                       \
            if (def) ISO_LAUNCH4(NS, false, ISO_PROFILE_DEFAULT, TRK);                 \
            else ISO_LAUNCH4(NS, false, ISO_PROFILE_GENERIC, TRK);                     \
        }                                                                              \
    } while (0)
    const bool def = models->profile_default;
    const bool track = models->track;
    const bool seq = bp->dev.ncols == 4;   // multi-star models with at most four bands: the star-sequential kernels
    switch (models->n_stars) {
    case 1:
        if (track) ISO_LAUNCH2(1, true);
        else ISO_LAUNCH2(1, false);
        break;
    case 2: ISO_LAUNCH2(2, false); break;
    case 3: ISO_LAUNCH2(3, false); break;
    default: return iso_set_error(ctx, ISO_E_INVALID, "lnpost: bad n_stars");
    }
#undef ISO_LAUNCH2
#undef ISO_LAUNCH4
#undef ISO_LAUNCH5
#undef ISO_LAUNCH6
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

__global__ void iso_mnest_prior_kernel(double *cube, const double *lo, const double *hi, int ndim, long long total)
{
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        int j = (int)(t % ndim);
        cube[t] = __dadd_rn(__dmul_rn(hi[j] - lo[j], cube[t]), lo[j]);   // starmodel.py:1637-1640 (unfused, bit-exact)
    }
}

struct MnestUser {
    const double *d_lo, *d_hi;
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
    iso_mnest_prior_kernel<<<blocks, 256, 0, st>>>((double *)d[1], u->d_lo, u->d_hi, u->ndim, total);
    ctx->launches++;
    ISO_CUDA(ctx, cudaGetLastError());
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
    ISO_REQUIRE(ctx, h_lo && h_hi && ndim >= 1 && ndim <= 64 && N >= 0, "iso_mnest_prior: bad argument");
    if (N == 0) return ISO_OK;
    ISO_REQUIRE(ctx, h_cube, "iso_mnest_prior: cube is NULL");
    IsoDeviceGuard guard(ctx->device);
    double *d_b = nullptr;
    ISO_CUDA(ctx, cudaMalloc(&d_b, sizeof(double) * 2 * ndim));
    cudaError_t e = cudaMemcpy(d_b, h_lo, sizeof(double) * ndim, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_b + ndim, h_hi, sizeof(double) * ndim, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        cudaFree(d_b);
        return iso_check_cuda(ctx, e, "iso_mnest_prior");
    }
    // input and output are the same host array: two pipeline arrays over it (read, then write back)
    IsoPipeArray arr[2];
    arr[0] = IsoPipeArray{h_cube, nullptr, (int64_t)8 * ndim};
    arr[1] = IsoPipeArray{nullptr, h_cube, (int64_t)8 * ndim};
    MnestUser u{d_b, d_b + ndim, ndim};
    int rc = iso_run_pipeline(ctx, N, arr, 2, mnest_launch, &u);
    cudaFree(d_b);
    return rc;
}

}  // extern "C"
