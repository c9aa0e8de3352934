// isochrones_b200 — on-device affine-invariant ensemble sampler (the emcee "stretch move", Goodman & Weare 2010).
//
// The reference drives its fits through emcee.EnsembleSampler(nwalkers, npars, mod.lnpost) (starmodel.py:966):
// per half-step emcee proposes nwalkers/2 points and calls lnpost on each from Python.  emcee itself is third-party
// and un-vendored (setup.py:60, unpinned); the algorithm restated here is its published stretch move:
//     z = ((a - 1) u + 1)^2 / a,   q = c_j - z (c_j - x_k),   accept iff  ln u' < (ndim - 1) ln z + lnpost(q) - lnpost(x_k)
// with c_j drawn uniformly from the complementary half of the ensemble, halves updated in turn.
//
// B200 mapping: ONE persistent CTA per chain, one thread per walker of the active half.  Walker positions and
// log-probabilities live in shared memory for the whole run, the two half-steps are separated by __syncthreads(),
// and lnpost is the same device function the batch kernel uses (iso_lnpost_row.cuh) — so a 256-walker x 2000-step
// fit is one kernel launch instead of 4000 batches of 128 rows, and many chains (catalog mode: one star per chain)
// fill the GPU.  Randomness is counter-based (Philox4x32-10 keyed by seed, counter = step / half / walker / chain),
// so a run is reproducible and independent of scheduling; tests replay the same stream on the host.
#include <math.h>
#include <stdlib.h>

#include "iso_lnpost_row.cuh"
#include "iso_stretch.cuh"
#include "iso_scratch.cuh"

#ifndef ISO_SAMPLER_SEG_STEPS
#define ISO_SAMPLER_SEG_STEPS 16   // steps a CTA advances a claimed chain by before it hands it back (multi-wave runs)
#endif

struct iso_sampler {
    const iso_grid *mp = nullptr, *bp = nullptr;
    const iso_models *models = nullptr;
    int device = 0;
    int n_chains = 0, n_walkers = 0, ndim = 0;
    uint64_t seed = 0;
    double a = 2.0;
    long long step = 0;             // full ensemble steps taken so far (RNG counter)
    double *d_pos = nullptr;        // [n_chains, n_walkers, ndim]
    double *d_lnprob = nullptr;     // [n_chains, n_walkers]
    unsigned long long *d_acc = nullptr;   // [n_chains] accepted proposals
    long long step_acc0 = 0;        // value of `step` when the acceptance counters were last zeroed (iso_sampler_reset)
    bool moments_on = false;        // accumulate d_mom at every kept step (iso_sampler_set_moments)
    int *d_queue = nullptr;         // work queue of multi-wave runs: [n_chains] progress, [n_chains] locks, [1] cursor
    double *d_mom = nullptr;        // [n_chains, 2 ndim + 1] running sums of the kept samples: sum x_d, sum x_d^2, count
};

struct IsoSamplerParams {
    IsoRowGrids G;
    const IsoModelDev *models;
    int n_models;
    int n_chains, n_walkers, n_steps, thin;
    long long step0;
    unsigned long long seed;
    double a;
    double *pos, *lnprob;
    double *chain_out, *lnprob_out;   // [n_steps / thin, n_chains, n_walkers, (ndim)] or NULL
    unsigned long long *accepted;
    double *moments;                  // [n_chains, 2 ndim + 1] (see iso_sampler) or NULL
    // work queue (seg_steps > 0): more chains than resident CTAs — a CTA claims a chain, advances it by one segment of
    // seg_steps steps, hands it back and claims the next, so that the last wave is as full as the first
    int seg_steps;
    int *progress;                    // [n_chains] steps of this run already taken
    int *lock;                        // [n_chains] 1 while a CTA holds the chain
    unsigned *cursor;                 // round-robin position of the next claim
    IsoModelDev model;                // the single model (n_models == 1)
};

// BOUNDS = 512: ensembles of up to 1024 walkers, 128 registers per thread, several CTAs per SM (many chains in flight);
// BOUNDS = 128: ensembles of up to 256 walkers when the chains do not fill the GPU anyway — one chain is bound by the
// dependent latency of a single row, and with 254 registers the compiler keeps more of a row's independent work in
// flight (a 256 x 2000 run: 29 ms instead of 33 ms).
// QUEUED: the work-queue form for more chains than resident CTAs (a template flag: the claim loop costs the 128-register
// instantiation a few more spilled values, which the direct form — one CTA per chain — should not pay).
template <int NSTARS, bool CATALOG, int PROFILE, bool TRACK, int BOUNDS, bool QUEUED = false>
__global__ void __launch_bounds__(BOUNDS, 1) iso_sampler_kernel(const __grid_constant__ IsoSamplerParams P)
{
    constexpr int NDIMP = NSTARS + 4;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2 *s_nodes = reinterpret_cast<double2 *>(smem_raw);
    double *s_pos = reinterpret_cast<double *>(smem_raw + sizeof(double2) * (P.G.smem_nodes > 0 ? P.G.smem_nodes : 1));
    double *s_lp = s_pos + (size_t)P.n_walkers * NDIMP;
    const int nhalf = P.n_walkers >> 1;
    const int t = threadIdx.x;
    __shared__ int s_claim[2];
    iso_stage_axis_tables(P.G, s_nodes);
    constexpr bool queued = QUEUED;
    for (;;) {
        int chain = blockIdx.x, s_begin = 0, s_end = P.n_steps;
        if (queued) {
            if (t == 0) {
                int c = -1, first = 0;
                for (int tries = 0; tries < P.n_chains; tries++) {
                    const int cand = (int)(atomicAdd(P.cursor, 1u) % (unsigned)P.n_chains);
                    if (*(volatile int *)(P.progress + cand) >= P.n_steps) continue;   // finished
                    if (atomicCAS(P.lock + cand, 0, 1) != 0) continue;                 // another CTA is advancing it
                    __threadfence();
                    first = *(volatile int *)(P.progress + cand);
                    if (first < P.n_steps) {
                        c = cand;
                        break;
                    }
                    atomicExch(P.lock + cand, 0);
                }
                s_claim[0] = c;
                s_claim[1] = first;
            }
            __syncthreads();
            chain = s_claim[0];
            if (chain < 0) break;   // every chain is finished or in another CTA's hands (which will finish it)
            s_begin = s_claim[1];
            s_end = min(P.n_steps, s_begin + P.seg_steps);
        }
        const IsoModelDev &m = CATALOG ? P.models[chain % P.n_models] : P.model;
        double *g_pos = P.pos + (size_t)chain * P.n_walkers * NDIMP;
        double *g_lp = P.lnprob + (size_t)chain * P.n_walkers;
        // the ensemble comes from L2 (ld.cg): a previous segment of this chain may have run on another SM
        for (int i = t; i < P.n_walkers * NDIMP; i += blockDim.x) s_pos[i] = __ldcg(g_pos + i);
        for (int i = t; i < P.n_walkers; i += blockDim.x) s_lp[i] = __ldcg(g_lp + i);
        __syncthreads();   // the ensemble in shared memory is complete before the first proposal reads it

        unsigned long long n_acc = 0;
        for (int s = s_begin; s < s_end; s++) {
            const unsigned long long gstep = (unsigned long long)(P.step0 + s);
#pragma unroll 1
            for (int half = 0; half < 2; half++) {
                if (t < nhalf) {
                    const int k = half * nhalf + t;            // walker being moved
                    const int other0 = (1 - half) * nhalf;     // first walker of the complementary half
                    double q[NDIMP], z, u_acc;
                    iso_stretch_propose<NDIMP>(P.seed, gstep, half, chain, k, other0, nhalf, P.a, s_pos, q, z, u_acc);
                    const IsoRowResult res = iso_lnpost_row<NSTARS, PROFILE, TRACK>(P.G, s_nodes, m, q, false, false);
                    if (iso_stretch_accept<NDIMP>(z, u_acc, res.lnpost, s_lp[k])) {
#pragma unroll
                        for (int d = 0; d < NDIMP; d++) s_pos[k * NDIMP + d] = q[d];
                        s_lp[k] = res.lnpost;
                        n_acc++;
                    }
                }
                __syncthreads();
            }
            if (P.thin > 0 && (s + 1) % P.thin == 0) {
                const long long keep = (s + 1) / P.thin - 1;
                if (P.chain_out) {
                    double *o = P.chain_out + ((size_t)keep * P.n_chains + chain) * P.n_walkers * NDIMP;
                    for (int i = t; i < P.n_walkers * NDIMP; i += blockDim.x) o[i] = s_pos[i];
                }
                if (P.lnprob_out) {
                    double *o = P.lnprob_out + ((size_t)keep * P.n_chains + chain) * P.n_walkers;
                    for (int i = t; i < P.n_walkers; i += blockDim.x) o[i] = s_lp[i];
                }
                if (P.moments) {
                    // running first / second moments of every kept ensemble (posterior mean and spread per chain without
                    // the samples ever leaving the GPU — what a catalog fit gathers per star): warp sums, one atomic per warp
                    double *mom = P.moments + (size_t)chain * (2 * NDIMP + 1);
#pragma unroll
                    for (int d = 0; d < NDIMP; d++) {
                        double s1 = 0.0, s2 = 0.0;
                        for (int w = t; w < P.n_walkers; w += blockDim.x) {
                            const double x = s_pos[w * NDIMP + d];
                            s1 += x;
                            s2 = fma(x, x, s2);
                        }
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) {
                            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
                            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
                        }
                        if ((t & 31) == 0) {
                            atomicAdd(mom + d, s1);
                            atomicAdd(mom + NDIMP + d, s2);
                        }
                    }
                    if (t == 0) atomicAdd(mom + 2 * NDIMP, (double)P.n_walkers);
                }
            }
        }
        for (int i = t; i < P.n_walkers * NDIMP; i += blockDim.x) g_pos[i] = s_pos[i];
        for (int i = t; i < P.n_walkers; i += blockDim.x) g_lp[i] = s_lp[i];
        if (n_acc) atomicAdd(P.accepted + chain, n_acc);
        if (!queued) break;
        __threadfence();    // the ensemble is in L2 before the chain is handed back
        __syncthreads();
        if (t == 0) {
            *(volatile int *)(P.progress + chain) = s_end;
            __threadfence();
            atomicExch(P.lock + chain, 0);
        }
    }
}

static void sampler_free(iso_sampler *s)
{
    if (!s) return;
    if (s->d_pos) cudaFree(s->d_pos);
    if (s->d_lnprob) cudaFree(s->d_lnprob);
    if (s->d_acc) cudaFree(s->d_acc);
    if (s->d_mom) cudaFree(s->d_mom);
    if (s->d_queue) cudaFree(s->d_queue);
    delete s;
}

extern "C" {

int iso_sampler_create(iso_ctx *ctx, const iso_grid *model_pack, const iso_grid *bc_pack, const iso_models *models,
                       int n_chains, int n_walkers, const double *h_p0, uint64_t seed, double stretch_a, iso_sampler **out)
{
    if (!ctx) return iso_set_error(nullptr, ISO_E_INVALID, "iso_sampler_create: ctx is NULL");
    ISO_REQUIRE(ctx, out, "iso_sampler_create: out is NULL");
    *out = nullptr;
    int rc = iso_check_lnpost_handles(ctx, model_pack, bc_pack, models);
    if (rc != ISO_OK) return rc;
    ISO_REQUIRE(ctx, n_chains >= 1 && h_p0, "iso_sampler_create: bad argument");
    ISO_REQUIRE(ctx, n_walkers >= 2 && n_walkers % 2 == 0 && n_walkers <= 1024,
                "iso_sampler_create: n_walkers must be even and at most 1024 (one thread per walker of a half)");
    ISO_REQUIRE(ctx, stretch_a > 1.0, "iso_sampler_create: stretch scale a must exceed 1");
    ISO_REQUIRE(ctx, models->n_models == 1 || models->n_models == n_chains,
                "iso_sampler_create: stage one model, or one model per chain (catalog mode)");
    IsoDeviceGuard guard(ctx->device);
    iso_sampler *s = new iso_sampler();
    s->mp = model_pack;
    s->bp = bc_pack;
    s->models = models;
    s->device = ctx->device;
    s->n_chains = n_chains;
    s->n_walkers = n_walkers;
    s->ndim = 4 + models->n_stars;
    s->seed = seed;
    s->a = stretch_a;
    const size_t n_rows = (size_t)n_chains * n_walkers;
    cudaError_t e = cudaMalloc(&s->d_pos, n_rows * s->ndim * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&s->d_lnprob, n_rows * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&s->d_acc, n_chains * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemsetAsync(s->d_acc, 0, n_chains * sizeof(unsigned long long), ctx->stream);
    const size_t mom_bytes = (size_t)n_chains * (2 * s->ndim + 1) * sizeof(double);
    if (e == cudaSuccess) e = cudaMalloc(&s->d_mom, mom_bytes);
    if (e == cudaSuccess) e = cudaMalloc(&s->d_queue, (2 * (size_t)n_chains + 1) * sizeof(int));
    if (e == cudaSuccess) e = cudaMemsetAsync(s->d_mom, 0, mom_bytes, ctx->stream);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(s->d_pos, h_p0, n_rows * s->ndim * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
    int *d_mor = nullptr;
    if (e == cudaSuccess && models->n_models > 1) {   // chain c uses model c
        std::vector<int> mor(n_rows);
        for (size_t i = 0; i < n_rows; i++) mor[i] = (int)(i / n_walkers);
        e = cudaMalloc(&d_mor, n_rows * sizeof(int));
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_mor, mor.data(), n_rows * sizeof(int), cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    }
    if (e != cudaSuccess) {
        if (d_mor) cudaFree(d_mor);
        sampler_free(s);
        return iso_check_cuda(ctx, e, "iso_sampler_create");
    }
    // lnpost of the initial ensemble through the batch kernel
    rc = iso_lnpost_batch_device(ctx, model_pack, bc_pack, models, d_mor, s->d_pos, (int64_t)n_rows, s->d_lnprob, nullptr, nullptr);
    e = cudaStreamSynchronize(ctx->stream);
    if (d_mor) cudaFree(d_mor);
    // a NaN initial lnpost (a walker whose star falls off the BC grid) can never be left — every proposal compares
    // against NaN — and emcee refuses such an ensemble ("The initial lnprob was NaN"); -inf walkers are legal
    long long first_nan = -1;
    if (rc == ISO_OK && e == cudaSuccess) {
        std::vector<double> lp(n_rows);
        e = cudaMemcpy(lp.data(), s->d_lnprob, n_rows * sizeof(double), cudaMemcpyDeviceToHost);
        for (size_t i = 0; i < n_rows && first_nan < 0; i++)
            if (lp[i] != lp[i]) first_nan = (long long)i;
    }
    if (rc != ISO_OK || e != cudaSuccess) {
        sampler_free(s);
        return rc != ISO_OK ? rc : iso_check_cuda(ctx, e, "iso_sampler_create");
    }
    if (first_nan >= 0) {
        sampler_free(s);
        return iso_set_error(ctx, ISO_E_INVALID, "iso_sampler_create: the initial lnpost of walker %lld of chain %lld is NaN "
                                                 "(outside the bolometric-correction grid); draw another starting point",
                             first_nan % n_walkers, first_nan / n_walkers);
    }
    *out = s;
    return ISO_OK;
}

int iso_sampler_run(iso_ctx *ctx, iso_sampler *s, int n_steps, int thin, double *h_chain, double *h_lnprob)
{
    if (!ctx) return iso_set_error(nullptr, ISO_E_INVALID, "iso_sampler_run: ctx is NULL");
    ISO_REQUIRE(ctx, s && n_steps >= 0 && thin >= 1, "iso_sampler_run: bad argument");
    ISO_REQUIRE(ctx, s->device == ctx->device, "iso_sampler_run: sampler belongs to another device");
    if (n_steps == 0) return ISO_OK;
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    IsoDeviceGuard guard(ctx->device);
    IsoSamplerParams P;
    size_t smem = 0;
    int rc = iso_row_grids_fill(ctx, s->mp, s->bp, &P.G, &smem);
    if (rc != ISO_OK) return rc;
    smem += (size_t)s->n_walkers * (s->ndim + 1) * sizeof(double);
    ISO_REQUIRE(ctx, smem <= 200 * 1024, "iso_sampler_run: ensemble does not fit in shared memory");
    const long long n_keep = n_steps / thin;
    const size_t rows = (size_t)s->n_chains * s->n_walkers;
    double *d_chain = nullptr, *d_lp = nullptr;
    if (h_chain && n_keep > 0) ISO_CUDA(ctx, iso_scratch_alloc(ctx, (void **)&d_chain, (size_t)n_keep * rows * s->ndim * sizeof(double)));
    if (h_lnprob && n_keep > 0) {
        cudaError_t e = iso_scratch_alloc(ctx, (void **)&d_lp, (size_t)n_keep * rows * sizeof(double));
        if (e != cudaSuccess) {
            iso_scratch_free(ctx, d_chain);
            return iso_check_cuda(ctx, e, "iso_sampler_run");
        }
    }
    P.models = s->models->d_models;
    P.n_models = s->models->n_models;
    P.model = s->models->h_first;
    P.n_chains = s->n_chains;
    P.n_walkers = s->n_walkers;
    P.n_steps = n_steps;
    P.thin = thin;
    P.step0 = s->step;
    P.seed = s->seed;
    P.a = s->a;
    P.pos = s->d_pos;
    P.lnprob = s->d_lnprob;
    P.chain_out = d_chain;
    P.lnprob_out = d_lp;
    P.accepted = s->d_acc;
    P.moments = (s->moments_on && n_keep > 0) ? s->d_mom : nullptr;
    P.seg_steps = 0;
    P.progress = s->d_queue;
    P.lock = s->d_queue + s->n_chains;
    P.cursor = (unsigned *)(s->d_queue + 2 * (size_t)s->n_chains);
    int threads = ((s->n_walkers / 2 + 31) / 32) * 32;
    // few chains of small ensembles: the latency-optimised instantiation (see the kernel's comment)
    const bool few_small = threads <= 128 && s->n_chains <= ctx->prop.multiProcessorCount;
    const bool catalog = s->models->n_models > 1;
    const bool def = s->models->profile_default, track = s->models->track;
    cudaError_t e = cudaSuccess;
    // More chains than the GPU holds at once, and not a whole number of waves: a persistent grid of exactly the resident
    // CTAs works through (chain, segment) items from a queue, so the makespan is total work / capacity instead of whole
    // waves of whole runs (1250 chains on 592 resident CTAs: 2.1 rounds instead of 3).  Hopping between chains costs
    // ~10 % (measured: 1184 chains = exactly two waves, 5.1 ms direct, 5.9 ms queued), so it is used only when the last
    // wave would be less than ~3/4 full.  Results do not depend on the schedule: the random stream is
    // a function of (chain, step, walker) only.
#define ISO_SLAUNCH5(NS, CAT, PROF, TRK, BND)                                                                          \
    do {                                                                                                              \
        if (smem > 48 * 1024)                                                                                         \
            e = cudaFuncSetAttribute(iso_sampler_kernel<NS, CAT, PROF, TRK, BND>,                                     \
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                         \
        int per_sm = 0;                                                                                               \
        if (e == cudaSuccess && BND == 512 && smem > 48 * 1024)                                                       \
            e = cudaFuncSetAttribute(iso_sampler_kernel<NS, CAT, PROF, TRK, BND, (BND == 512)>,                       \
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                         \
        if (e == cudaSuccess && BND == 512)                                                                           \
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, iso_sampler_kernel<NS, CAT, PROF, TRK, BND, (BND == 512)>, \
                                                              threads, smem);                                         \
        const int capacity = per_sm * ctx->prop.multiProcessorCount;                                                  \
        P.seg_steps = 0;                                                                                              \
        /* whole waves of whole runs waste ceil(w) / w - 1 of the time (w = chains / capacity); the queue costs ~10 % */ \
        const double waves = capacity > 0 ? (double)s->n_chains / capacity : 0.0;                                     \
        if (e == cudaSuccess && waves > 1.0 && ceil(waves) / waves > 1.12 && n_steps >= 2 * ISO_SAMPLER_SEG_STEPS) {  \
            P.seg_steps = n_steps / 16 > ISO_SAMPLER_SEG_STEPS ? n_steps / 16 : ISO_SAMPLER_SEG_STEPS;                \
            e = cudaMemsetAsync(s->d_queue, 0, (2 * (size_t)s->n_chains + 1) * sizeof(int), ctx->stream);             \
            if (e == cudaSuccess)                                                                                     \
                iso_sampler_kernel<NS, CAT, PROF, TRK, BND, (BND == 512)><<<capacity, threads, smem, ctx->stream>>>(P); \
        } else if (e == cudaSuccess) {                                                                                \
            iso_sampler_kernel<NS, CAT, PROF, TRK, BND, false><<<s->n_chains, threads, smem, ctx->stream>>>(P);       \
        }                                                                                                             \
    } while (0)
#define ISO_SLAUNCH4(NS, CAT, PROF, TRK)                                        \
    do {                                                                       \
        if (few_small) ISO_SLAUNCH5(NS, CAT, PROF, TRK, 128);                  \
        else ISO_SLAUNCH5(NS, CAT, PROF, TRK, 512);                            \
    } while (0)
#define ISO_SLAUNCH2(NS, TRK)                                                    \
    do {                                                                         \
        if (catalog) {                                                           \
            if (def) ISO_SLAUNCH4(NS, true, ISO_PROFILE_DEFAULT, TRK);           \
            else ISO_SLAUNCH4(NS, true, ISO_PROFILE_GENERIC, TRK);               \
        } else {                                                                 \
            if (def) ISO_SLAUNCH4(NS, false, ISO_PROFILE_DEFAULT, TRK);          \
            else ISO_SLAUNCH4(NS, false, ISO_PROFILE_GENERIC, TRK);              \
        }                                                                        \
    } while (0)
    switch (s->models->n_stars) {
    case 1:
        if (track) ISO_SLAUNCH2(1, true);
        else ISO_SLAUNCH2(1, false);
        break;
    case 2: ISO_SLAUNCH2(2, false); break;
    default: ISO_SLAUNCH2(3, false); break;
    }
#undef ISO_SLAUNCH2
#undef ISO_SLAUNCH4
#undef ISO_SLAUNCH5
    ctx->launches++;
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e == cudaSuccess && d_chain)
        e = cudaMemcpyAsync(h_chain, d_chain, (size_t)n_keep * rows * s->ndim * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess && d_lp)
        e = cudaMemcpyAsync(h_lnprob, d_lp, (size_t)n_keep * rows * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    iso_scratch_free(ctx, d_chain);
    iso_scratch_free(ctx, d_lp);
    if (e != cudaSuccess) return iso_check_cuda(ctx, e, "iso_sampler_run");
    s->step += n_steps;
    return ISO_OK;
}

int iso_sampler_state(iso_ctx *ctx, iso_sampler *s, double *h_pos, double *h_lnprob, int64_t *n_accepted, int64_t *n_proposed)
{
    if (!ctx) return iso_set_error(nullptr, ISO_E_INVALID, "iso_sampler_state: ctx is NULL");
    ISO_REQUIRE(ctx, s, "iso_sampler_state: sampler is NULL");
    IsoDeviceGuard guard(ctx->device);
    const size_t rows = (size_t)s->n_chains * s->n_walkers;
    if (h_pos) ISO_CUDA(ctx, cudaMemcpyAsync(h_pos, s->d_pos, rows * s->ndim * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (h_lnprob) ISO_CUDA(ctx, cudaMemcpyAsync(h_lnprob, s->d_lnprob, rows * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    std::vector<unsigned long long> acc(s->n_chains);
    ISO_CUDA(ctx, cudaMemcpyAsync(acc.data(), s->d_acc, s->n_chains * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    ISO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (n_accepted)   // [n_chains]
        for (int c = 0; c < s->n_chains; c++) n_accepted[c] = (int64_t)acc[c];
    if (n_proposed) *n_proposed = (int64_t)(s->step - s->step_acc0) * s->n_walkers;   // per chain, since the last reset
    return ISO_OK;
}

int iso_sampler_reset(iso_ctx *ctx, iso_sampler *s)
{
    if (!ctx) return iso_set_error(nullptr, ISO_E_INVALID, "iso_sampler_reset: ctx is NULL");
    ISO_REQUIRE(ctx, s, "iso_sampler_reset: sampler is NULL");
    IsoDeviceGuard guard(s->device);
    ISO_CUDA(ctx, cudaMemsetAsync(s->d_acc, 0, s->n_chains * sizeof(unsigned long long), ctx->stream));
    ISO_CUDA(ctx, cudaMemsetAsync(s->d_mom, 0, (size_t)s->n_chains * (2 * s->ndim + 1) * sizeof(double), ctx->stream));
    s->step_acc0 = s->step;
    return ISO_OK;
}

int iso_sampler_set_moments(iso_ctx *ctx, iso_sampler *s, int enable)
{
    if (!ctx) return iso_set_error(nullptr, ISO_E_INVALID, "iso_sampler_set_moments: ctx is NULL");
    ISO_REQUIRE(ctx, s, "iso_sampler_set_moments: sampler is NULL");
    s->moments_on = enable != 0;
    return ISO_OK;
}

int iso_sampler_moments(iso_ctx *ctx, iso_sampler *s, double *h_moments, const double **d_moments)
{
    if (!ctx) return iso_set_error(nullptr, ISO_E_INVALID, "iso_sampler_moments: ctx is NULL");
    ISO_REQUIRE(ctx, s, "iso_sampler_moments: sampler is NULL");
    IsoDeviceGuard guard(s->device);
    if (d_moments) *d_moments = s->d_mom;
    if (h_moments) {
        ISO_CUDA(ctx, cudaMemcpyAsync(h_moments, s->d_mom, (size_t)s->n_chains * (2 * s->ndim + 1) * sizeof(double),
                                      cudaMemcpyDeviceToHost, ctx->stream));
        ISO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return ISO_OK;
}

int iso_sampler_destroy(iso_ctx *ctx, iso_sampler *s)
{
    if (!s) return ISO_OK;
    IsoDeviceGuard guard(s->device);
    if (ctx) cudaStreamSynchronize(ctx->stream);
    sampler_free(s);
    return ISO_OK;
}

}  // extern "C"
