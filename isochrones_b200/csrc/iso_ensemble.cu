// isochrones_b200 — ONE ensemble of walkers sharded over the GPUs of a node ("with the on-device
// stretch-move sampler, one all-gather of the accepted half-ensemble + their lnprob per half-step").
//
// The reference's samplers evaluate the walkers of a half-step one Python call at a time (emcee, starmodel.py:966) or
// spread live points over MPI ranks (MultiNest, starmodel.py:755-797).  Here every rank (one process per GPU) OWNS a
// contiguous block of each half of the ensemble; in a half-step a rank proposes and evaluates its block of the active
// half, and every proposal GATHERS its partner walker c_j — drawn uniformly from the complementary half — straight from
// the HBM of the rank that owns it: plain loads through CUDA-IPC peer mappings, carried by NVLink.  The exchange is
// fused into the evaluation kernel — no collective is launched, and only the 8 * ndim bytes a proposal actually needs
// travel (a first version pushed every accepted walker into every rank's copy instead: 7 x 48 B of scattered 8-byte
// remote stores per accepted walker made the half-step SLOWER with more GPUs — 0.131 ms at N = 8 against 0.098 ms on
// one GPU for 2^20 walkers, profiles/README.md).  Accepted walkers are written to the owner's own memory only.
// Completion is the flag protocol of iso_peer.cu: the last CTA of a half-step publishes the half-step number in this
// rank's slot of every rank's flag array (system-scope release), and the NEXT half-step kernel starts by waiting
// (bounded, ISO_E_TIMEOUT) until every rank has published the previous one.  That one wait orders both directions: the
// partners a rank is about to read are final (their owners finished the previous half-step), and a rank starts
// overwriting its block of a half only after every peer has finished reading it.  At the end of a run (and at every
// kept step of a thinned chain) each rank pushes its two blocks into every peer's copy with coalesced stores
// (iso_ensemble_share_kernel), so between runs every rank holds the whole ensemble: iso_ensemble_state reads locally.
// A run starts with a token of its own: the peers may not store into a copy its host may still be reading.
//
// The proposal is iso_stretch.cuh's — the same code, the same Philox counters (half-step, walker) as the one-GPU
// persistent sampler — so the chain does not depend on the number of ranks: world-size-N and single-GPU runs agree bit
// for bit (tests/test_gpu_ensemble.py).  Ensembles are not limited by shared memory here (1e5-1e6 walkers are fine).
#include <string.h>

#include <vector>

#include "iso_lnpost_row.cuh"
#include "iso_stretch.cuh"
#include "iso_scratch.cuh"

struct iso_ensemble {
    const iso_grid *mp = nullptr, *bp = nullptr;
    const iso_models *models = nullptr;
    int device = 0, rank = 0, nranks = 1;
    int n_walkers = 0, ndim = 0;
    uint64_t seed = 0;
    double a = 2.0;
    long long step = 0;                    // full ensemble steps taken
    long long step_acc0 = 0;
    unsigned long long published = 0;      // half-steps published so far (flag value)
    double *d_state = nullptr;             // [n_walkers, ndim] positions followed by [n_walkers] lnpost (own allocation)
    unsigned long long *d_flags = nullptr; // [ISO_MAX_PEERS]
    double *peer_state[ISO_MAX_PEERS] = {nullptr};
    unsigned long long *peer_flags[ISO_MAX_PEERS] = {nullptr};
    unsigned *d_done = nullptr;
    unsigned long long *d_acc = nullptr;   // proposals accepted by THIS rank
    bool connected = false;
    unsigned long long timeout_ns = 10ULL * 1000 * 1000 * 1000;
    unsigned *h_err = nullptr, *d_err = nullptr;
};

struct IsoEnsembleParams {
    IsoRowGrids G;
    IsoModelDev model;
    double *state;                         // this rank's copy: pos [n_walkers, ndim], then lnpost [n_walkers]
    double *peer_state[ISO_MAX_PEERS];     // every rank's copy; rank r's blocks of it are the authoritative ones
    unsigned long long *peer_flags[ISO_MAX_PEERS];
    const unsigned long long *own_flags;
    unsigned long long wait_for, publish;  // half-step numbers: wait until every rank published `wait_for`, then publish
    unsigned long long seed, gstep, timeout_ns;
    unsigned *done, *err;
    unsigned long long *accepted;
    double a;
    int n_walkers, half, first, count;     // this rank moves walkers [half * nhalf + first, .. + count)
    int n_peers, rank;
    int blk_base, blk_extra;               // block sizes of a half: ranks < blk_extra own blk_base + 1 walkers, the rest blk_base
};

__device__ __forceinline__ unsigned long long iso_ens_globaltimer()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// every thread's writes are ordered before its CTA's arrival (the CTA barrier, then the system-scope fences of warp 0:
// cumulative, as in a grid-wide barrier); the last CTA to arrive publishes `publish` in this rank's slot on every rank
__device__ __forceinline__ void iso_ensemble_publish(const IsoEnsembleParams &P)
{
    __syncthreads();
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        unsigned last = 0;
        if (lane == 0) {
            __threadfence_system();
            last = atomicAdd(P.done, 1u) == gridDim.x - 1;
            if (last) *P.done = 0;
        }
        last = __shfl_sync(0xffffffffu, last, 0);
        if (last) {
            // lane r serves rank r: one fence, then relaxed stores side by side (release stores one after the other would
            // each wait for the previous one's NVLink round trip)
            __threadfence_system();
            if (lane < P.n_peers)
                asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(P.peer_flags[lane] + P.rank), "l"(P.publish) : "memory");
        }
    }
}

// bounded wait (lanes 0 .. n_peers - 1 of every CTA) until every rank has published `wait_for`
__device__ __forceinline__ void iso_ensemble_await(const IsoEnsembleParams &P)
{
    if (P.wait_for > 0) {
        const int r = threadIdx.x;
        if (r < P.n_peers) {
            const unsigned long long t0 = iso_ens_globaltimer();
            unsigned spins = 0;
            for (;;) {
                unsigned long long v;
                asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(P.own_flags + r) : "memory");
                if (v >= P.wait_for) break;
                if ((++spins & 255u) == 0 && iso_ens_globaltimer() - t0 > P.timeout_ns) {
                    atomicCAS_system(P.err, 0u, (unsigned)(r + 1) | ((unsigned)P.wait_for << 8));
                    break;
                }
            }
        }
        __syncthreads();
    }
}

// A walker's position record (NDIMP doubles, 8-byte aligned) read as whole 32-byte sectors: one 256-bit load per sector
// the record touches (2, for NDIMP = 7 sometimes 3) instead of NDIMP 8-byte loads — over NVLink every load instruction
// of a scattered gather is a read request of its own.  L2-only (.cg): the record was written by another kernel, maybe on
// another GPU.  Sectors lie inside the rank's state allocation (the lnpost array follows the positions).
template <int NDIMP>
__device__ __forceinline__ void iso_gather_record(const double *rec, double (&c)[NDIMP])
{
    const unsigned long long addr = (unsigned long long)rec;
    const int sh = (int)((addr & 31ULL) >> 3);   // doubles between the sector boundary and the record
    const double *base = reinterpret_cast<const double *>(addr & ~31ULL);
    constexpr int NSEC = (NDIMP + 3 + 3) / 4;
    double v[4 * NSEC];
#pragma unroll
    for (int sct = 0; sct < NSEC; sct++) {
        v[4 * sct] = v[4 * sct + 1] = v[4 * sct + 2] = v[4 * sct + 3] = 0.0;
        if (4 * sct < sh + NDIMP)
            asm volatile("ld.global.cg.v4.f64 {%0,%1,%2,%3}, [%4];"
                         : "=d"(v[4 * sct]), "=d"(v[4 * sct + 1]), "=d"(v[4 * sct + 2]), "=d"(v[4 * sct + 3])
                         : "l"(base + 4 * sct));
    }
#pragma unroll
    for (int d = 0; d < NDIMP; d++) {
        double t = v[d];
        if (d + 1 < 4 * NSEC) t = sh == 1 ? v[d + 1] : t;
        if (d + 2 < 4 * NSEC) t = sh == 2 ? v[d + 2] : t;
        if (d + 3 < 4 * NSEC) t = sh == 3 ? v[d + 3] : t;
        c[d] = t;
    }
}

template <int NSTARS, int PROFILE, bool TRACK>
__global__ void __launch_bounds__(256, 2) iso_ensemble_half_kernel(const __grid_constant__ IsoEnsembleParams P)
{
    constexpr int NDIMP = NSTARS + 4;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2 *s_nodes = reinterpret_cast<double2 *>(smem_raw);
    iso_stage_axis_tables(P.G, s_nodes);
    // every rank has finished (and published) the previous half-step: the partners we are about to read are final, and
    // nobody still reads the block we are about to overwrite
    iso_ensemble_await(P);
    const int nhalf = P.n_walkers >> 1;
    const int other0 = (1 - P.half) * nhalf;
    double *pos = P.state;
    double *lp = P.state + (size_t)P.n_walkers * NDIMP;
    const int big = P.blk_extra * (P.blk_base + 1);   // walkers of a half owned by the ranks with one walker more
    unsigned long long n_acc = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P.count; i += gridDim.x * blockDim.x) {
        const int k = P.half * nhalf + P.first + i;
        int j_rel;
        double z, u_acc;
        iso_stretch_draw(P.seed, P.gstep, P.half, 0, k, nhalf, P.a, j_rel, z, u_acc);
        // the partner lives in the copy of the rank that owns block(j_rel) of the complementary half
        const int owner = P.n_peers == 1 ? 0 : (j_rel < big ? j_rel / (P.blk_base + 1) : P.blk_extra + (j_rel - big) / P.blk_base);
        const double *partner = P.peer_state[owner] + (size_t)(other0 + j_rel) * NDIMP;
        double c[NDIMP], x[NDIMP], q[NDIMP];
        iso_gather_record<NDIMP>(partner, c);
#pragma unroll
        for (int d = 0; d < NDIMP; d++) x[d] = pos[(size_t)k * NDIMP + d];
        iso_stretch_point<NDIMP>(c, x, z, q);
        const IsoRowResult res = iso_lnpost_row<NSTARS, PROFILE, TRACK>(P.G, s_nodes, P.model, q, false, false);
        if (iso_stretch_accept<NDIMP>(z, u_acc, res.lnpost, lp[k])) {
#pragma unroll
            for (int d = 0; d < NDIMP; d++) pos[(size_t)k * NDIMP + d] = q[d];
            lp[k] = res.lnpost;
            n_acc++;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n_acc += __shfl_xor_sync(0xffffffffu, n_acc, o);
    if ((threadIdx.x & 31) == 0 && n_acc) atomicAdd(P.accepted, n_acc);
    iso_ensemble_publish(P);
}

// Replication: this rank's blocks of both halves (positions + lnpost) go into every peer's copy — contiguous runs,
// coalesced stores — and the rank then publishes.  Launched once every rank has published the last half-step.
__global__ void __launch_bounds__(256) iso_ensemble_share_kernel(const __grid_constant__ IsoEnsembleParams P, int ndim)
{
    iso_ensemble_await(P);
    const int nhalf = P.n_walkers >> 1;
    const double *lp = P.state + (size_t)P.n_walkers * ndim;
    const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x, nthr = (long long)gridDim.x * blockDim.x;
    for (int r = 0; r < P.n_peers; r++) {
        if (r == P.rank) continue;
        double *dst = P.peer_state[r];
        for (int half = 0; half < 2; half++) {
            const size_t k0 = (size_t)half * nhalf + P.first;
            for (long long i = tid; i < (long long)P.count * ndim; i += nthr) dst[k0 * ndim + i] = P.state[k0 * ndim + i];
            for (long long i = tid; i < P.count; i += nthr) dst[(size_t)P.n_walkers * ndim + k0 + i] = lp[k0 + i];
        }
    }
    iso_ensemble_publish(P);
}

// final wait of a run (and the wait before a kept ensemble is copied out): same bounded spin as the kernel's prologue
__global__ void iso_ensemble_wait_kernel(const unsigned long long *own_flags, int n, unsigned long long step,
                                         unsigned long long timeout_ns, unsigned *err)
{
    const int r = threadIdx.x;
    if (r < n) {
        const unsigned long long t0 = iso_ens_globaltimer();
        unsigned spins = 0;
        for (;;) {
            unsigned long long v;
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(own_flags + r) : "memory");
            if (v >= step) break;
            if ((++spins & 255u) == 0 && iso_ens_globaltimer() - t0 > timeout_ns) {
                atomicCAS_system(err, 0u, (unsigned)(r + 1) | ((unsigned)step << 8));
                break;
            }
        }
    }
    __syncthreads();
    __threadfence_system();
}

// start-of-run token: a rank's copy may be read by its host (iso_ensemble_state) between runs, so the peers may only
// start storing into it again once this rank has entered the next run
__global__ void iso_ensemble_signal_kernel(IsoEnsembleParams P)
{
    const int r = threadIdx.x;
    if (r < P.n_peers) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(P.peer_flags[r] + P.rank), "l"(P.publish) : "memory");
}

static void ensemble_free(iso_ensemble *e)
{
    if (!e) return;
    for (int r = 0; r < e->nranks; r++) {
        if (r == e->rank) continue;
        if (e->peer_state[r]) cudaIpcCloseMemHandle(e->peer_state[r]);
        if (e->peer_flags[r]) cudaIpcCloseMemHandle(e->peer_flags[r]);
    }
    if (e->d_state) cudaFree(e->d_state);
    if (e->d_flags) cudaFree(e->d_flags);
    if (e->d_done) cudaFree(e->d_done);
    if (e->d_acc) cudaFree(e->d_acc);
    if (e->h_err) cudaFreeHost(e->h_err);
    delete e;
}

static int ensemble_timed_out(iso_ctx *ctx, const iso_ensemble *e)
{
    const unsigned v = *(volatile unsigned *)e->h_err;
    if (!v) return ISO_OK;
    return iso_set_error(ctx, ISO_E_TIMEOUT, "sharded ensemble: rank %u did not publish half-step %u (mod 2^24) within %.1f s; "
                                             "the ensemble copies have diverged",
                         (v & 0xffu) - 1u, v >> 8, (double)e->timeout_ns * 1e-9);
}

// balanced contiguous blocks of the active half: rank r moves walkers [first, first + count) of it
static void ensemble_block(int nhalf, int nranks, int rank, int *first, int *count)
{
    const int base = nhalf / nranks, extra = nhalf % nranks;
    *count = base + (rank < extra ? 1 : 0);
    *first = rank * base + (rank < extra ? rank : extra);
}

extern "C" {

int iso_ensemble_create(iso_ctx *ctx, const iso_grid *model_pack, const iso_grid *bc_pack, const iso_models *models,
                        int n_walkers, const double *h_p0, uint64_t seed, double stretch_a, int rank, int nranks,
                        iso_ensemble **out)
{
    if (!ctx) return iso_set_error(nullptr, ISO_E_INVALID, "iso_ensemble_create: ctx is NULL");
    ISO_REQUIRE(ctx, out, "iso_ensemble_create: out is NULL");
    *out = nullptr;
    int rc = iso_check_lnpost_handles(ctx, model_pack, bc_pack, models);
    if (rc != ISO_OK) return rc;
    ISO_REQUIRE(ctx, h_p0 && n_walkers >= 2 && n_walkers % 2 == 0, "iso_ensemble_create: n_walkers must be even and >= 2");
    ISO_REQUIRE(ctx, nranks >= 1 && nranks <= ISO_MAX_PEERS && rank >= 0 && rank < nranks,
                "iso_ensemble_create: bad rank / nranks (at most 8 ranks of one node)");
    ISO_REQUIRE(ctx, models->n_models == 1, "iso_ensemble_create: one star model per ensemble");
    ISO_REQUIRE(ctx, stretch_a > 1.0, "iso_ensemble_create: stretch scale a must exceed 1");
    IsoDeviceGuard guard(ctx->device);
    iso_ensemble *e = new iso_ensemble();
    e->mp = model_pack;
    e->bp = bc_pack;
    e->models = models;
    e->device = ctx->device;
    e->rank = rank;
    e->nranks = nranks;
    e->n_walkers = n_walkers;
    e->ndim = 4 + models->n_stars;
    e->seed = seed;
    e->a = stretch_a;
    const size_t pos_bytes = (size_t)n_walkers * e->ndim * sizeof(double);
    cudaError_t ce = cudaMalloc(&e->d_state, pos_bytes + (size_t)n_walkers * sizeof(double));
    if (ce == cudaSuccess) ce = cudaMalloc(&e->d_flags, sizeof(unsigned long long) * ISO_MAX_PEERS);
    if (ce == cudaSuccess) ce = cudaMalloc(&e->d_done, sizeof(unsigned));
    if (ce == cudaSuccess) ce = cudaMalloc(&e->d_acc, sizeof(unsigned long long));
    if (ce == cudaSuccess) ce = cudaHostAlloc((void **)&e->h_err, sizeof(unsigned), cudaHostAllocMapped);
    if (ce == cudaSuccess) {
        *e->h_err = 0;
        ce = cudaHostGetDevicePointer((void **)&e->d_err, e->h_err, 0);
    }
    if (ce == cudaSuccess) ce = cudaMemsetAsync(e->d_flags, 0, sizeof(unsigned long long) * ISO_MAX_PEERS, ctx->stream);
    if (ce == cudaSuccess) ce = cudaMemsetAsync(e->d_done, 0, sizeof(unsigned), ctx->stream);
    if (ce == cudaSuccess) ce = cudaMemsetAsync(e->d_acc, 0, sizeof(unsigned long long), ctx->stream);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(e->d_state, h_p0, pos_bytes, cudaMemcpyHostToDevice, ctx->stream);
    if (ce != cudaSuccess) {
        ensemble_free(e);
        return iso_check_cuda(ctx, ce, "iso_ensemble_create");
    }
    // every rank evaluates the whole initial ensemble itself (identical inputs, identical code: identical copies)
    rc = iso_lnpost_batch_device(ctx, model_pack, bc_pack, models, nullptr, e->d_state, n_walkers,
                                 e->d_state + (size_t)n_walkers * e->ndim, nullptr, nullptr);
    std::vector<double> lp((size_t)n_walkers);
    if (rc == ISO_OK) {
        ce = cudaMemcpyAsync(lp.data(), e->d_state + (size_t)n_walkers * e->ndim, (size_t)n_walkers * sizeof(double),
                             cudaMemcpyDeviceToHost, ctx->stream);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(ctx->stream);
    }
    if (rc != ISO_OK || ce != cudaSuccess) {
        ensemble_free(e);
        return rc != ISO_OK ? rc : iso_check_cuda(ctx, ce, "iso_ensemble_create");
    }
    for (int i = 0; i < n_walkers; i++)
        if (lp[i] != lp[i]) {
            ensemble_free(e);
            return iso_set_error(ctx, ISO_E_INVALID, "iso_ensemble_create: the initial lnpost of walker %d is NaN (outside the "
                                                     "bolometric-correction grid); draw another starting point", i);
        }
    e->peer_state[rank] = e->d_state;
    e->peer_flags[rank] = e->d_flags;
    e->connected = nranks == 1;
    *out = e;
    return ISO_OK;
}

int iso_ensemble_export(iso_ctx *ctx, iso_ensemble *e, void *handle128)
{
    if (!ctx) return iso_set_error(nullptr, ISO_E_INVALID, "iso_ensemble_export: ctx is NULL");
    ISO_REQUIRE(ctx, e && handle128, "iso_ensemble_export: NULL argument");
    IsoDeviceGuard guard(e->device);
    cudaIpcMemHandle_t h[2];
    ISO_CUDA(ctx, cudaIpcGetMemHandle(&h[0], e->d_state));
    ISO_CUDA(ctx, cudaIpcGetMemHandle(&h[1], e->d_flags));
    memcpy(handle128, h, sizeof(h));
    return ISO_OK;
}

int iso_ensemble_connect(iso_ctx *ctx, iso_ensemble *e, const void *handles)
{
    if (!ctx) return iso_set_error(nullptr, ISO_E_INVALID, "iso_ensemble_connect: ctx is NULL");
    ISO_REQUIRE(ctx, e && handles, "iso_ensemble_connect: NULL argument");
    ISO_REQUIRE(ctx, !e->connected, "iso_ensemble_connect: already connected");
    IsoDeviceGuard guard(e->device);
    for (int r = 0; r < e->nranks; r++) {
        if (r == e->rank) continue;
        cudaIpcMemHandle_t h[2];
        memcpy(h, (const char *)handles + (size_t)r * 128, sizeof(h));
        void *p = nullptr;
        ISO_CUDA(ctx, cudaIpcOpenMemHandle(&p, h[0], cudaIpcMemLazyEnablePeerAccess));
        e->peer_state[r] = (double *)p;
        ISO_CUDA(ctx, cudaIpcOpenMemHandle(&p, h[1], cudaIpcMemLazyEnablePeerAccess));
        e->peer_flags[r] = (unsigned long long *)p;
    }
    e->connected = true;
    return ISO_OK;
}

int iso_ensemble_set_timeout(iso_ctx *ctx, iso_ensemble *e, double seconds)
{
    if (!ctx) return iso_set_error(nullptr, ISO_E_INVALID, "iso_ensemble_set_timeout: ctx is NULL");
    ISO_REQUIRE(ctx, e && seconds > 0.0 && seconds < 1e6, "iso_ensemble_set_timeout: bad argument");
    e->timeout_ns = (unsigned long long)(seconds * 1e9);
    return ISO_OK;
}

int iso_ensemble_run(iso_ctx *ctx, iso_ensemble *e, int n_steps, int thin, double *h_chain, double *h_lnprob)
{
    if (!ctx) return iso_set_error(nullptr, ISO_E_INVALID, "iso_ensemble_run: ctx is NULL");
    ISO_REQUIRE(ctx, e && n_steps >= 0 && thin >= 1, "iso_ensemble_run: bad argument");
    ISO_REQUIRE(ctx, e->connected, "iso_ensemble_run: ensemble not connected to its peers");
    ISO_REQUIRE(ctx, e->device == ctx->device, "iso_ensemble_run: ensemble belongs to another device");
    int rc = ensemble_timed_out(ctx, e);
    if (rc != ISO_OK) return rc;
    if (n_steps == 0) return ISO_OK;
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    IsoDeviceGuard guard(ctx->device);
    IsoEnsembleParams P;
    size_t smem = 0;
    rc = iso_row_grids_fill(ctx, e->mp, e->bp, &P.G, &smem);
    if (rc != ISO_OK) return rc;
    P.model = e->models->h_first;
    P.state = e->d_state;
    P.own_flags = e->d_flags;
    for (int r = 0; r < ISO_MAX_PEERS; r++) {
        P.peer_state[r] = r < e->nranks ? e->peer_state[r] : nullptr;
        P.peer_flags[r] = r < e->nranks ? e->peer_flags[r] : nullptr;
    }
    P.seed = e->seed;
    P.timeout_ns = e->timeout_ns;
    P.done = e->d_done;
    P.err = e->d_err;
    P.accepted = e->d_acc;
    P.a = e->a;
    P.n_walkers = e->n_walkers;
    P.n_peers = e->nranks;
    P.rank = e->rank;
    const int nhalf = e->n_walkers / 2;
    ensemble_block(nhalf, e->nranks, e->rank, &P.first, &P.count);
    P.blk_base = nhalf / e->nranks;
    P.blk_extra = nhalf % e->nranks;
    const long long n_keep = n_steps / thin;
    const size_t pos_n = (size_t)e->n_walkers * e->ndim;
    double *d_chain = nullptr, *d_lp = nullptr;
    if (h_chain && n_keep > 0) ISO_CUDA(ctx, iso_scratch_alloc(ctx, (void **)&d_chain, (size_t)n_keep * pos_n * sizeof(double)));
    if (h_lnprob && n_keep > 0) {
        cudaError_t ce = iso_scratch_alloc(ctx, (void **)&d_lp, (size_t)n_keep * e->n_walkers * sizeof(double));
        if (ce != cudaSuccess) {
            iso_scratch_free(ctx, d_chain);
            return iso_check_cuda(ctx, ce, "iso_ensemble_run");
        }
    }
    const bool def = e->models->profile_default, track = e->models->track;
    int blocks = (P.count + 255) / 256;
    const int cap = ctx->prop.multiProcessorCount * 2;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;   // a rank without walkers in the half still waits and publishes
    cudaError_t ce = cudaSuccess;
    P.half = 0;
    P.gstep = 0;
    P.wait_for = 0;
    P.publish = ++e->published;   // the start-of-run token
    iso_ensemble_signal_kernel<<<1, 32, 0, ctx->stream>>>(P);
    ctx->launches++;
#define ISO_ELAUNCH(NS, PROF, TRK)                                                                                       \
    do {                                                                                                                 \
        if (smem > 48 * 1024)                                                                                            \
            ce = cudaFuncSetAttribute(iso_ensemble_half_kernel<NS, PROF, TRK>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                      (int)smem);                                                                        \
        if (ce == cudaSuccess) iso_ensemble_half_kernel<NS, PROF, TRK><<<blocks, 256, smem, ctx->stream>>>(P);           \
    } while (0)
    // replication + completion wait: every rank's blocks are in every copy when the wait kernel returns
    int share_blocks = (int)(((long long)P.count * e->ndim + 255) / 256);
    if (share_blocks > cap) share_blocks = cap;
    if (share_blocks < 1) share_blocks = 1;
    auto replicate = [&]() {
        if (e->nranks > 1) {
            P.wait_for = e->published;
            P.publish = ++e->published;
            iso_ensemble_share_kernel<<<share_blocks, 256, 0, ctx->stream>>>(P, e->ndim);
            ctx->launches++;
        }
        iso_ensemble_wait_kernel<<<1, 32, 0, ctx->stream>>>(e->d_flags, e->nranks, e->published, e->timeout_ns, e->d_err);
        ctx->launches++;
    };
    for (int s = 0; s < n_steps && ce == cudaSuccess; s++) {
        P.gstep = (unsigned long long)(e->step + s);
        const bool keep = (s + 1) % thin == 0 && (d_chain || d_lp);
        const long long kept = (s + 1) / thin - 1;
        for (int half = 0; half < 2 && ce == cudaSuccess; half++) {
            P.half = half;
            P.wait_for = e->published;
            P.publish = ++e->published;
            switch (e->models->n_stars) {
            case 1:
                if (track) {
                    if (def) ISO_ELAUNCH(1, ISO_PROFILE_DEFAULT, true);
                    else ISO_ELAUNCH(1, ISO_PROFILE_GENERIC, true);
                } else {
                    if (def) ISO_ELAUNCH(1, ISO_PROFILE_DEFAULT, false);
                    else ISO_ELAUNCH(1, ISO_PROFILE_GENERIC, false);
                }
                break;
            case 2:
                if (def) ISO_ELAUNCH(2, ISO_PROFILE_DEFAULT, false);
                else ISO_ELAUNCH(2, ISO_PROFILE_GENERIC, false);
                break;
            default:
                if (def) ISO_ELAUNCH(3, ISO_PROFILE_DEFAULT, false);
                else ISO_ELAUNCH(3, ISO_PROFILE_GENERIC, false);
                break;
            }
            ctx->launches++;
        }
        if (ce == cudaSuccess && keep) {
            // a kept ensemble is whole only after a replication; the copy is then a plain device-to-device one (nobody
            // writes into this rank's copy before this rank publishes its next half-step)
            replicate();
            if (d_chain)
                ce = cudaMemcpyAsync(d_chain + (size_t)kept * pos_n, e->d_state, pos_n * sizeof(double), cudaMemcpyDeviceToDevice,
                                     ctx->stream);
            if (ce == cudaSuccess && d_lp)
                ce = cudaMemcpyAsync(d_lp + (size_t)kept * e->n_walkers, e->d_state + pos_n, (size_t)e->n_walkers * sizeof(double),
                                     cudaMemcpyDeviceToDevice, ctx->stream);
        }
    }
#undef ISO_ELAUNCH
    if (ce == cudaSuccess) {
        replicate();
        ce = cudaGetLastError();
    }
    if (ce == cudaSuccess && d_chain)
        ce = cudaMemcpyAsync(h_chain, d_chain, (size_t)n_keep * pos_n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
    if (ce == cudaSuccess && d_lp)
        ce = cudaMemcpyAsync(h_lnprob, d_lp, (size_t)n_keep * e->n_walkers * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(ctx->stream);
    iso_scratch_free(ctx, d_chain);
    iso_scratch_free(ctx, d_lp);
    if (ce != cudaSuccess) return iso_check_cuda(ctx, ce, "iso_ensemble_run");
    e->step += n_steps;
    return ensemble_timed_out(ctx, e);
}

int iso_ensemble_state(iso_ctx *ctx, iso_ensemble *e, double *h_pos, double *h_lnprob, int64_t *n_accepted_local,
                       int64_t *n_proposed)
{
    if (!ctx) return iso_set_error(nullptr, ISO_E_INVALID, "iso_ensemble_state: ctx is NULL");
    ISO_REQUIRE(ctx, e, "iso_ensemble_state: ensemble is NULL");
    IsoDeviceGuard guard(e->device);
    const size_t pos_n = (size_t)e->n_walkers * e->ndim;
    if (h_pos) ISO_CUDA(ctx, cudaMemcpyAsync(h_pos, e->d_state, pos_n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (h_lnprob)
        ISO_CUDA(ctx, cudaMemcpyAsync(h_lnprob, e->d_state + pos_n, (size_t)e->n_walkers * sizeof(double), cudaMemcpyDeviceToHost,
                                      ctx->stream));
    unsigned long long acc = 0;
    ISO_CUDA(ctx, cudaMemcpyAsync(&acc, e->d_acc, sizeof(acc), cudaMemcpyDeviceToHost, ctx->stream));
    ISO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (n_accepted_local) *n_accepted_local = (int64_t)acc;
    if (n_proposed) *n_proposed = (int64_t)(e->step - e->step_acc0) * e->n_walkers;
    return ensemble_timed_out(ctx, e);
}

int iso_ensemble_destroy(iso_ctx *ctx, iso_ensemble *e)
{
    if (!e) return ISO_OK;
    IsoDeviceGuard guard(e->device);
    if (ctx) cudaStreamSynchronize(ctx->stream);
    ensemble_free(e);
    return ISO_OK;
}

}  // extern "C"
