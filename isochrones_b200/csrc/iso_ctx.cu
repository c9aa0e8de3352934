// isochrones_b200 — context, memory, timing and error plumbing of the C ABI (include/isochrones_b200.h).
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "iso_common.cuh"
#include "iso_scratch.cuh"

static thread_local std::string g_last_error;  // failures of calls that had no context

int iso_set_error(iso_ctx *ctx, int code, const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) ctx->last_error = buf;
    g_last_error = buf;
    return code;
}

int iso_check_cuda(iso_ctx *ctx, cudaError_t e, const char *what)
{
    if (e == cudaSuccess) return ISO_OK;
    cudaGetLastError();  // clear the sticky flag of non-fatal errors
    int code = (e == cudaErrorMemoryAllocation) ? ISO_E_NOMEM : ISO_E_CUDA;
    return iso_set_error(ctx, code, "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
}

extern "C" {

int iso_abi_version(void) { return ISO_ABI_VERSION; }

int64_t iso_struct_size(int which)
{
    switch (which) {
    case 0: return (int64_t)sizeof(iso_prior_leaf);
    case 1: return (int64_t)sizeof(iso_prior);
    case 2: return (int64_t)sizeof(iso_model);
    default: return -1;
    }
}

int iso_device_count(int *count)
{
    if (!count) return iso_set_error(nullptr, ISO_E_INVALID, "iso_device_count: count is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        *count = 0;
        return iso_check_cuda(nullptr, e, "cudaGetDeviceCount");
    }
    *count = n;
    return ISO_OK;
}

const char *iso_last_error(iso_ctx *ctx) { return ctx ? ctx->last_error.c_str() : g_last_error.c_str(); }

int iso_ctx_create(int device, iso_ctx **out)
{
    if (!out) return iso_set_error(nullptr, ISO_E_INVALID, "iso_ctx_create: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return iso_check_cuda(nullptr, e, "cudaGetDeviceCount (no CUDA device: there is no CPU fallback)");
    if (device < 0 || device >= n)
        return iso_set_error(nullptr, n == 0 ? ISO_E_CUDA : ISO_E_INVALID, "iso_ctx_create: device %d of %d", device, n);
    iso_ctx *ctx = new iso_ctx();
    ctx->device = device;
    IsoDeviceGuard guard(device);
#define CTX_TRY(call)                                     \
    do {                                                  \
        cudaError_t e2 = (call);                          \
        if (e2 != cudaSuccess) {                          \
            int rc = iso_check_cuda(nullptr, e2, #call);  \
            delete ctx;                                   \
            return rc;                                    \
        }                                                 \
    } while (0)
    CTX_TRY(cudaGetDeviceProperties(&ctx->prop, device));
    if (ctx->prop.major < 10) {
        int rc = iso_set_error(nullptr, ISO_E_UNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a only",
                               device, ctx->prop.major, ctx->prop.minor);
        delete ctx;
        return rc;
    }
    CTX_TRY(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) {
        CTX_TRY(cudaStreamCreateWithFlags(&ctx->copy_stream[i], cudaStreamNonBlocking));
        CTX_TRY(cudaEventCreateWithFlags(&ctx->ev_copy[i], cudaEventDisableTiming));
        CTX_TRY(cudaEventCreateWithFlags(&ctx->ev_done[i], cudaEventDisableTiming));
    }
    CTX_TRY(cudaEventCreate(&ctx->ev_start));
    CTX_TRY(cudaEventCreate(&ctx->ev_stop));
    {   // scratch allocations (iso_scratch_alloc) come from the device's default pool; let it keep freed blocks
        cudaMemPool_t pool;
        CTX_TRY(cudaDeviceGetDefaultMemPool(&pool, device));
        uint64_t keep = 1ull << 30;
        CTX_TRY(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    }
    CTX_TRY(cudaMalloc(&ctx->d_claim, ISO_CLAIM_SLOTS * ISO_CLAIM_STRIDE * sizeof(unsigned long long)));
    CTX_TRY(cudaMemset(ctx->d_claim, 0, ISO_CLAIM_SLOTS * ISO_CLAIM_STRIDE * sizeof(unsigned long long)));
#undef CTX_TRY
    *out = ctx;
    return ISO_OK;
}

int iso_ctx_destroy(iso_ctx *ctx)
{
    if (!ctx) return ISO_OK;
    IsoDeviceGuard guard(ctx->device);
    cudaDeviceSynchronize();
    iso_nccl_destroy(ctx);
    for (int i = 0; i < 2; i++) {
        if (ctx->d_stage[i]) cudaFree(ctx->d_stage[i]);
        if (ctx->h_stage[i]) cudaFreeHost(ctx->h_stage[i]);
        if (ctx->copy_stream[i]) cudaStreamDestroy(ctx->copy_stream[i]);
        if (ctx->ev_copy[i]) cudaEventDestroy(ctx->ev_copy[i]);
        if (ctx->ev_done[i]) cudaEventDestroy(ctx->ev_done[i]);
    }
    if (ctx->ev_start) cudaEventDestroy(ctx->ev_start);
    if (ctx->ev_stop) cudaEventDestroy(ctx->ev_stop);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->d_small) cudaFree(ctx->d_small);
    if (ctx->d_claim) cudaFree(ctx->d_claim);
    delete ctx;
    return ISO_OK;
}

int iso_ctx_sync(iso_ctx *ctx)
{
    if (!ctx) return iso_set_error(nullptr, ISO_E_INVALID, "iso_ctx_sync: ctx is NULL");
    IsoDeviceGuard guard(ctx->device);
    ISO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ISO_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream[0]));
    ISO_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream[1]));
    return ISO_OK;
}

int iso_ctx_info(iso_ctx *ctx, char *name, int *sm_count, int64_t *l2_bytes, int64_t *hbm_bytes, int *cc)
{
    if (!ctx) return iso_set_error(nullptr, ISO_E_INVALID, "iso_ctx_info: ctx is NULL");
    if (name) {
        strncpy(name, ctx->prop.name, 255);
        name[255] = 0;
    }
    if (sm_count) *sm_count = ctx->prop.multiProcessorCount;
    if (l2_bytes) *l2_bytes = ctx->prop.l2CacheSize;
    if (hbm_bytes) *hbm_bytes = (int64_t)ctx->prop.totalGlobalMem;
    if (cc) *cc = ctx->prop.major * 10 + ctx->prop.minor;
    return ISO_OK;
}

}  // extern "C"

cudaError_t iso_scratch_alloc(iso_ctx *ctx, void **d_ptr, size_t bytes)
{
    *d_ptr = nullptr;
    cudaError_t e = cudaMallocAsync(d_ptr, bytes > 0 ? bytes : 1, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);   // allocated for every stream, not only ours
    return e;
}

void iso_scratch_free(iso_ctx *ctx, void *d_ptr)
{
    if (d_ptr) cudaFreeAsync(d_ptr, ctx->stream);
}

extern "C" {

int iso_dev_alloc(iso_ctx *ctx, int64_t bytes, void **d_ptr)
{
    if (!ctx || !d_ptr || bytes < 0) return iso_set_error(ctx, ISO_E_INVALID, "iso_dev_alloc: bad argument");
    IsoDeviceGuard guard(ctx->device);
    *d_ptr = nullptr;
    ISO_CUDA(ctx, cudaMalloc(d_ptr, (size_t)(bytes > 0 ? bytes : 1)));
    return ISO_OK;
}

int iso_dev_free(iso_ctx *ctx, void *d_ptr)
{
    if (!ctx) return iso_set_error(nullptr, ISO_E_INVALID, "iso_dev_free: ctx is NULL");
    IsoDeviceGuard guard(ctx->device);
    if (d_ptr) ISO_CUDA(ctx, cudaFree(d_ptr));
    return ISO_OK;
}

int iso_host_alloc(iso_ctx *ctx, int64_t bytes, void **h_ptr)
{
    if (!ctx || !h_ptr || bytes < 0) return iso_set_error(ctx, ISO_E_INVALID, "iso_host_alloc: bad argument");
    IsoDeviceGuard guard(ctx->device);
    *h_ptr = nullptr;
    ISO_CUDA(ctx, cudaMallocHost(h_ptr, (size_t)(bytes > 0 ? bytes : 1)));
    return ISO_OK;
}

int iso_host_free(iso_ctx *ctx, void *h_ptr)
{
    if (!ctx) return iso_set_error(nullptr, ISO_E_INVALID, "iso_host_free: ctx is NULL");
    IsoDeviceGuard guard(ctx->device);
    if (h_ptr) ISO_CUDA(ctx, cudaFreeHost(h_ptr));
    return ISO_OK;
}

int iso_memcpy_h2d(iso_ctx *ctx, void *d_dst, const void *h_src, int64_t bytes)
{
    if (!ctx || (bytes > 0 && (!d_dst || !h_src)) || bytes < 0)
        return iso_set_error(ctx, ISO_E_INVALID, "iso_memcpy_h2d: bad argument");
    IsoDeviceGuard guard(ctx->device);
    ISO_CUDA(ctx, cudaMemcpyAsync(d_dst, h_src, (size_t)bytes, cudaMemcpyHostToDevice, ctx->stream));
    ISO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ISO_OK;
}

int iso_memcpy_d2h(iso_ctx *ctx, void *h_dst, const void *d_src, int64_t bytes)
{
    if (!ctx || (bytes > 0 && (!h_dst || !d_src)) || bytes < 0)
        return iso_set_error(ctx, ISO_E_INVALID, "iso_memcpy_d2h: bad argument");
    IsoDeviceGuard guard(ctx->device);
    ISO_CUDA(ctx, cudaMemcpyAsync(h_dst, d_src, (size_t)bytes, cudaMemcpyDeviceToHost, ctx->stream));
    ISO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ISO_OK;
}

int iso_memset(iso_ctx *ctx, void *d_dst, int value, int64_t bytes)
{
    if (!ctx || (bytes > 0 && !d_dst) || bytes < 0) return iso_set_error(ctx, ISO_E_INVALID, "iso_memset: bad argument");
    IsoDeviceGuard guard(ctx->device);
    ISO_CUDA(ctx, cudaMemsetAsync(d_dst, value, (size_t)bytes, ctx->stream));
    return ISO_OK;
}

int iso_timer_start(iso_ctx *ctx)
{
    if (!ctx) return iso_set_error(nullptr, ISO_E_INVALID, "iso_timer_start: ctx is NULL");
    IsoDeviceGuard guard(ctx->device);
    ISO_CUDA(ctx, cudaEventRecord(ctx->ev_start, ctx->stream));
    return ISO_OK;
}

int iso_timer_stop(iso_ctx *ctx, float *ms)
{
    if (!ctx || !ms) return iso_set_error(ctx, ISO_E_INVALID, "iso_timer_stop: bad argument");
    IsoDeviceGuard guard(ctx->device);
    ISO_CUDA(ctx, cudaEventRecord(ctx->ev_stop, ctx->stream));
    ISO_CUDA(ctx, cudaEventSynchronize(ctx->ev_stop));
    ISO_CUDA(ctx, cudaEventElapsedTime(ms, ctx->ev_start, ctx->ev_stop));
    return ISO_OK;
}

int iso_launch_count(iso_ctx *ctx, int64_t *count)
{
    if (!ctx || !count) return iso_set_error(ctx, ISO_E_INVALID, "iso_launch_count: bad argument");
    *count = ctx->launches;
    return ISO_OK;
}

}  // extern "C"

// Grow-on-demand staging buffers (device + pinned host) used by the host-pointer entry points.
int iso_stage_reserve(iso_ctx *ctx, int slot, int64_t dev_bytes, int64_t host_bytes)
{
    if (dev_bytes > ctx->d_stage_bytes[slot]) {
        if (ctx->d_stage[slot]) {
            ISO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            ISO_CUDA(ctx, cudaFree(ctx->d_stage[slot]));
            ctx->d_stage[slot] = nullptr;
            ctx->d_stage_bytes[slot] = 0;
        }
        ISO_CUDA(ctx, cudaMalloc(&ctx->d_stage[slot], (size_t)dev_bytes));
        ctx->d_stage_bytes[slot] = dev_bytes;
    }
    if (host_bytes > ctx->h_stage_bytes[slot]) {
        if (ctx->h_stage[slot]) {
            ISO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            ISO_CUDA(ctx, cudaFreeHost(ctx->h_stage[slot]));
            ctx->h_stage[slot] = nullptr;
            ctx->h_stage_bytes[slot] = 0;
        }
        ISO_CUDA(ctx, cudaMallocHost(&ctx->h_stage[slot], (size_t)host_bytes));
        ctx->h_stage_bytes[slot] = host_bytes;
    }
    return ISO_OK;
}
