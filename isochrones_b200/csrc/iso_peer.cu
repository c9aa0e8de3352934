// isochrones_b200 — fused lnpost + all-gather over NVLink peer memory (one process per GPU).
//
// The only exchange of the sharded lnpost path is the all-gather of per-row results a sampler's acceptance step
// needs.  With NCCL that is a second operation after the kernel (iso_allgather_f64); here the lnpost
// kernel itself stores every result into the receive buffer of EVERY rank — ordinary 8-byte stores through CUDA-IPC
// peer mappings, carried by NVLink / NVSwitch — so the transfer overlaps the evaluation row by row and no collective
// is launched.  Completion is a flag exchange: the last CTA of the kernel writes the step number into this rank's slot
// of every rank's flag array (system-scope release), and a one-warp kernel then waits until all slots of the rank's own
// array show the step (acquire).
//
// Receive buffers are double-buffered by step parity: a rank that has passed the wait of step s may start writing
// step s + 1 into its peers while a slower peer still reads step s from the other buffer; it cannot reach step s + 2
// before that peer has signalled s + 1, which (stream order) is after the peer finished with step s.
#include <string.h>

#include "iso_lnpost_row.cuh"

struct iso_peer_group {
    int device = 0, rank = 0, nranks = 1;
    int64_t pad = 0;                      // rows per rank in a receive buffer
    double *d_recv = nullptr;             // [2][nranks * pad]   (own allocation)
    unsigned long long *d_flags = nullptr;   // [nranks]         (own allocation)
    double *peer_recv[ISO_MAX_PEERS] = {nullptr};               // peer-mapped (own entry = d_recv)
    unsigned long long *peer_flags[ISO_MAX_PEERS] = {nullptr};
    unsigned *d_done = nullptr;           // CTA arrival counter of the fused kernel
    bool connected = false;
    unsigned long long step = 0;
    // bounded wait: a rank that does not publish a step within timeout_ns is reported instead of hanging the stream
    unsigned long long timeout_ns = 10ULL * 1000 * 1000 * 1000;
    unsigned *h_err = nullptr;            // page-locked, device-mapped: 0, or (missing rank + 1) | (step << 8) of the first timeout
    unsigned *d_err = nullptr;            // device alias of h_err
};

// the fused kernel's last CTA publishes the step number in this rank's slot of every rank's flag array (release,
// system scope); this kernel waits until every rank has published it in ours
//
// The spin is bounded: lane r gives up when rank r has not published the step within `timeout_ns` of wall-clock time
// (%globaltimer) and records (r + 1) | (step << 8) in the group's error word (page-locked host memory, first timeout
// wins); the host turns that into ISO_E_TIMEOUT on its next call / iso_peer_check.  A dead or stalled rank therefore
// costs one timeout, not a hung stream.
__device__ __forceinline__ unsigned long long iso_globaltimer()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__global__ void iso_peer_wait_kernel(const unsigned long long *own_flags, int n, unsigned long long step,
                                     unsigned long long timeout_ns, unsigned *err)
{
    const int r = threadIdx.x;
    if (r < n) {
        const unsigned long long t0 = iso_globaltimer();
        unsigned spins = 0;
        for (;;) {
            unsigned long long v;
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(own_flags + r) : "memory");
            if (v >= step) break;
            if ((++spins & 255u) == 0 && iso_globaltimer() - t0 > timeout_ns) {
                atomicCAS_system(err, 0u, (unsigned)(r + 1) | ((unsigned)step << 8));
                break;
            }
        }
    }
    __syncthreads();
    __threadfence_system();
}

// a rank with no rows in a step still has to publish the step (nobody may wait for it forever)
__global__ void iso_peer_signal_kernel(IsoPeerTargets t)
{
    const int r = threadIdx.x;
    if (r < t.n) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(t.flags[r] + t.rank), "l"(t.step) : "memory");
}

static void peer_free(iso_peer_group *g)
{
    if (!g) return;
    for (int r = 0; r < g->nranks; r++) {
        if (r == g->rank) continue;
        if (g->peer_recv[r]) cudaIpcCloseMemHandle(g->peer_recv[r]);
        if (g->peer_flags[r]) cudaIpcCloseMemHandle(g->peer_flags[r]);
    }
    if (g->d_recv) cudaFree(g->d_recv);
    if (g->d_flags) cudaFree(g->d_flags);
    if (g->d_done) cudaFree(g->d_done);
    if (g->h_err) cudaFreeHost(g->h_err);
    delete g;
}

// the error word of a group -> ISO_E_TIMEOUT (sticky: the receive buffers of a timed-out step are incomplete)
static int peer_timed_out(iso_ctx *ctx, const iso_peer_group *g)
{
    const unsigned e = *(volatile unsigned *)g->h_err;
    if (!e) return ISO_OK;
    return iso_set_error(ctx, ISO_E_TIMEOUT, "fused all-gather: rank %u did not publish step %u (mod 2^24) within %.1f s; the "
                                             "gathered buffer of that step is incomplete",
                         (e & 0xffu) - 1u, e >> 8, (double)g->timeout_ns * 1e-9);
}

extern "C" {

int iso_peer_set_timeout(iso_ctx *ctx, iso_peer_group *g, double seconds)
{
    if (!ctx) return iso_set_error(nullptr, ISO_E_INVALID, "iso_peer_set_timeout: ctx is NULL");
    ISO_REQUIRE(ctx, g && seconds > 0.0 && seconds < 1e6, "iso_peer_set_timeout: bad argument");
    g->timeout_ns = (unsigned long long)(seconds * 1e9);
    return ISO_OK;
}

int iso_peer_check(iso_ctx *ctx, iso_peer_group *g)
{
    if (!ctx) return iso_set_error(nullptr, ISO_E_INVALID, "iso_peer_check: ctx is NULL");
    ISO_REQUIRE(ctx, g, "iso_peer_check: group is NULL");
    IsoDeviceGuard guard(g->device);
    ISO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return peer_timed_out(ctx, g);
}

int iso_peer_create(iso_ctx *ctx, int rank, int nranks, int64_t rows_per_rank, iso_peer_group **out)
{
    if (!ctx) return iso_set_error(nullptr, ISO_E_INVALID, "iso_peer_create: ctx is NULL");
    ISO_REQUIRE(ctx, out && nranks >= 1 && nranks <= ISO_MAX_PEERS && rank >= 0 && rank < nranks && rows_per_rank >= 1,
                "iso_peer_create: bad argument (at most 8 ranks)");
    *out = nullptr;
    IsoDeviceGuard guard(ctx->device);
    iso_peer_group *g = new iso_peer_group();
    g->device = ctx->device;
    g->rank = rank;
    g->nranks = nranks;
    g->pad = rows_per_rank;
    const size_t bytes = (size_t)2 * nranks * rows_per_rank * sizeof(double);
    cudaError_t e = cudaMalloc(&g->d_recv, bytes);
    if (e == cudaSuccess) e = cudaMalloc(&g->d_flags, sizeof(unsigned long long) * ISO_MAX_PEERS);
    if (e == cudaSuccess) e = cudaMalloc(&g->d_done, sizeof(unsigned));
    if (e == cudaSuccess) e = cudaHostAlloc((void **)&g->h_err, sizeof(unsigned), cudaHostAllocMapped);
    if (e == cudaSuccess) {
        *g->h_err = 0;
        e = cudaHostGetDevicePointer((void **)&g->d_err, g->h_err, 0);
    }
    if (e == cudaSuccess) e = cudaMemsetAsync(g->d_done, 0, sizeof(unsigned), ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(g->d_recv, 0, bytes, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(g->d_flags, 0, sizeof(unsigned long long) * ISO_MAX_PEERS, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        peer_free(g);
        return iso_check_cuda(ctx, e, "iso_peer_create");
    }
    g->peer_recv[rank] = g->d_recv;
    g->peer_flags[rank] = g->d_flags;
    g->connected = nranks == 1;
    *out = g;
    return ISO_OK;
}

int iso_peer_export(iso_ctx *ctx, iso_peer_group *g, void *handle128)
{
    if (!ctx) return iso_set_error(nullptr, ISO_E_INVALID, "iso_peer_export: ctx is NULL");
    ISO_REQUIRE(ctx, g && handle128, "iso_peer_export: NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "two IPC handles travel in 128 bytes");
    IsoDeviceGuard guard(g->device);
    cudaIpcMemHandle_t h[2];
    ISO_CUDA(ctx, cudaIpcGetMemHandle(&h[0], g->d_recv));
    ISO_CUDA(ctx, cudaIpcGetMemHandle(&h[1], g->d_flags));
    memcpy(handle128, h, sizeof(h));
    return ISO_OK;
}

int iso_peer_connect(iso_ctx *ctx, iso_peer_group *g, const void *handles)
{
    if (!ctx) return iso_set_error(nullptr, ISO_E_INVALID, "iso_peer_connect: ctx is NULL");
    ISO_REQUIRE(ctx, g && handles, "iso_peer_connect: NULL argument");
    ISO_REQUIRE(ctx, !g->connected, "iso_peer_connect: already connected");
    IsoDeviceGuard guard(g->device);
    for (int r = 0; r < g->nranks; r++) {
        if (r == g->rank) continue;
        cudaIpcMemHandle_t h[2];
        memcpy(h, (const char *)handles + (size_t)r * 128, sizeof(h));
        void *p = nullptr;
        ISO_CUDA(ctx, cudaIpcOpenMemHandle(&p, h[0], cudaIpcMemLazyEnablePeerAccess));
        g->peer_recv[r] = (double *)p;
        ISO_CUDA(ctx, cudaIpcOpenMemHandle(&p, h[1], cudaIpcMemLazyEnablePeerAccess));
        g->peer_flags[r] = (unsigned long long *)p;
    }
    g->connected = true;
    return ISO_OK;
}

int iso_lnpost_allgather_device(iso_ctx *ctx, const iso_grid *model_pack, const iso_grid *bc_pack, const iso_models *models,
                                const int32_t *d_model_of_row, const double *d_pars, int64_t N, iso_peer_group *g,
                                const double **d_gathered)
{
    if (!ctx) return iso_set_error(nullptr, ISO_E_INVALID, "iso_lnpost_allgather_device: ctx is NULL");
    ISO_REQUIRE(ctx, g && g->connected, "iso_lnpost_allgather_device: peer group not connected");
    ISO_REQUIRE(ctx, g->device == ctx->device, "iso_lnpost_allgather_device: peer group belongs to another device");
    ISO_REQUIRE(ctx, N >= 0 && N <= g->pad, "iso_lnpost_allgather_device: more rows than the group was created for");
    {
        int trc = peer_timed_out(ctx, g);
        if (trc != ISO_OK) return trc;
    }
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    IsoDeviceGuard guard(ctx->device);
    const unsigned long long step = ++g->step;
    const size_t half = (size_t)(step & 1) * (size_t)g->nranks * (size_t)g->pad;
    IsoPeerTargets t;
    t.n = g->nranks;
    t.rank = g->rank;
    t.step = step;
    t.done = g->d_done;
    t.offset = (long long)g->rank * g->pad;
    t.own_flags = ISO_PEER_INKERNEL_WAIT ? g->d_flags : nullptr;
    t.timeout_ns = g->timeout_ns;
    t.err = g->d_err;
    for (int r = 0; r < ISO_MAX_PEERS; r++) {
        t.out[r] = r < g->nranks ? g->peer_recv[r] + half : nullptr;
        t.flags[r] = r < g->nranks ? g->peer_flags[r] : nullptr;
    }
    if (N > 0) {
        int rc = iso_lnpost_launch_peers(ctx, model_pack, bc_pack, models, d_model_of_row, d_pars, N, &t);
        if (rc != ISO_OK) {
            --g->step;   // nothing was launched: the step did not happen
            return rc;
        }
    } else {
        iso_peer_signal_kernel<<<1, 32, 0, ctx->stream>>>(t);
        ctx->launches += 1;
    }
    if (!ISO_PEER_INKERNEL_WAIT || N == 0) {
        iso_peer_wait_kernel<<<1, 32, 0, ctx->stream>>>(g->d_flags, g->nranks, step, g->timeout_ns, g->d_err);
        ctx->launches += 1;
    }
    ISO_CUDA(ctx, cudaGetLastError());
    if (d_gathered) *d_gathered = g->d_recv + half;
    return ISO_OK;
}

int iso_peer_destroy(iso_ctx *ctx, iso_peer_group *g)
{
    if (!g) return ISO_OK;
    IsoDeviceGuard guard(g->device);
    if (ctx) cudaStreamSynchronize(ctx->stream);
    peer_free(g);
    return ISO_OK;
}

}  // extern "C"
