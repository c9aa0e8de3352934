// isochrones_b200 — multi-GPU exchange: rows are sharded across ranks with no data-path collective; the only
// exchange is the all-gather of per-row results the sampler's acceptance step needs.
// NCCL is bound at run time (dlopen of libnccl.so.2) so the library itself has no link-time dependency on it.
#include <dlfcn.h>
#include <string.h>

#include "iso_common.cuh"

typedef struct { char internal[128]; } iso_ncclUniqueId;
typedef void *iso_ncclComm_t;
enum { ISO_NCCL_FLOAT64 = 8 };   // ncclDataType_t::ncclFloat64

struct NcclApi {
    void *handle = nullptr;
    int (*GetUniqueId)(iso_ncclUniqueId *) = nullptr;
    int (*CommInitRank)(iso_ncclComm_t *, int, iso_ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(iso_ncclComm_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, iso_ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    int (*GetVersion)(int *) = nullptr;
};

static NcclApi g_nccl;

static int load_nccl(iso_ctx *ctx)
{
    if (g_nccl.handle) return ISO_OK;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (int i = 0; i < 2 && !h; i++) h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
    if (!h) return iso_set_error(ctx, ISO_E_NCCL, "cannot load libnccl.so.2: %s", dlerror());
    NcclApi api;
    api.handle = h;
    api.GetUniqueId = (int (*)(iso_ncclUniqueId *))dlsym(h, "ncclGetUniqueId");
    api.CommInitRank = (int (*)(iso_ncclComm_t *, int, iso_ncclUniqueId, int))dlsym(h, "ncclCommInitRank");
    api.CommDestroy = (int (*)(iso_ncclComm_t))dlsym(h, "ncclCommDestroy");
    api.AllGather = (int (*)(const void *, void *, size_t, int, iso_ncclComm_t, cudaStream_t))dlsym(h, "ncclAllGather");
    api.GetErrorString = (const char *(*)(int))dlsym(h, "ncclGetErrorString");
    api.GetVersion = (int (*)(int *))dlsym(h, "ncclGetVersion");
    if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllGather || !api.GetErrorString)
        return iso_set_error(ctx, ISO_E_NCCL, "libnccl.so.2 lacks a required symbol");
    g_nccl = api;
    return ISO_OK;
}

#define ISO_NCCL(ctx, call)                                                                             \
    do {                                                                                                \
        int r__ = (call);                                                                               \
        if (r__ != 0) return iso_set_error((ctx), ISO_E_NCCL, "%s: %s", #call, g_nccl.GetErrorString(r__)); \
    } while (0)

extern "C" {

int iso_nccl_unique_id(void *id128)
{
    if (!id128) return iso_set_error(nullptr, ISO_E_INVALID, "iso_nccl_unique_id: NULL");
    int rc = load_nccl(nullptr);
    if (rc != ISO_OK) return rc;
    iso_ncclUniqueId id;
    ISO_NCCL(nullptr, g_nccl.GetUniqueId(&id));
    memcpy(id128, &id, sizeof(id));
    return ISO_OK;
}

int iso_nccl_init(iso_ctx *ctx, const void *id128, int rank, int nranks)
{
    if (!ctx) return iso_set_error(nullptr, ISO_E_INVALID, "iso_nccl_init: ctx is NULL");
    ISO_REQUIRE(ctx, id128 && nranks >= 1 && rank >= 0 && rank < nranks, "iso_nccl_init: bad argument");
    ISO_REQUIRE(ctx, !ctx->nccl_comm, "iso_nccl_init: communicator already initialised");
    int rc = load_nccl(ctx);
    if (rc != ISO_OK) return rc;
    IsoDeviceGuard guard(ctx->device);
    ISO_CUDA(ctx, cudaSetDevice(ctx->device));
    iso_ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    iso_ncclComm_t comm = nullptr;
    ISO_NCCL(ctx, g_nccl.CommInitRank(&comm, nranks, id, rank));
    ctx->nccl_comm = comm;
    ctx->nccl_rank = rank;
    ctx->nccl_nranks = nranks;
    return ISO_OK;
}

int iso_nccl_destroy(iso_ctx *ctx)
{
    if (!ctx || !ctx->nccl_comm) return ISO_OK;
    IsoDeviceGuard guard(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    g_nccl.CommDestroy((iso_ncclComm_t)ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
    ctx->nccl_nranks = 1;
    ctx->nccl_rank = 0;
    return ISO_OK;
}

int iso_allgather_f64(iso_ctx *ctx, const double *d_send, int64_t n, double *d_recv)
{
    if (!ctx) return iso_set_error(nullptr, ISO_E_INVALID, "iso_allgather_f64: ctx is NULL");
    ISO_REQUIRE(ctx, n >= 0 && (n == 0 || (d_send && d_recv)), "iso_allgather_f64: bad argument");
    if (n == 0) return ISO_OK;
    IsoDeviceGuard guard(ctx->device);
    if (!ctx->nccl_comm) {   // single rank without a communicator: the gather is a device copy
        ISO_REQUIRE(ctx, ctx->nccl_nranks == 1, "iso_allgather_f64: no communicator");
        if (d_send != d_recv)
            ISO_CUDA(ctx, cudaMemcpyAsync(d_recv, d_send, (size_t)n * 8, cudaMemcpyDeviceToDevice, ctx->stream));
        return ISO_OK;
    }
    ISO_NCCL(ctx, g_nccl.AllGather(d_send, d_recv, (size_t)n, ISO_NCCL_FLOAT64, (iso_ncclComm_t)ctx->nccl_comm, ctx->stream));
    return ISO_OK;
}

}  // extern "C"
