// isochrones_b200 — vector evaluation of one prior object: lnpdf(x) / __call__(x) of priors.py (see iso_prior.cuh).
#include "iso_common.cuh"
#include "iso_prior.cuh"
#include "iso_scratch.cuh"

__global__ void iso_prior_eval_kernel(const iso_prior *__restrict__ prior, int which, const double *__restrict__ x,
                                      double *__restrict__ out, long long N)
{
    __shared__ iso_prior p;
    {
        const int *src = reinterpret_cast<const int *>(prior);
        int *dst = reinterpret_cast<int *>(&p);
        for (int t = threadIdx.x; t < (int)(sizeof(iso_prior) / sizeof(int)); t += blockDim.x) dst[t] = src[t];
    }
    __syncthreads();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x)
        out[i] = which == 0 ? iso_prior_lnpdf_dyn(&p, x[i]) : iso_prior_call_dyn(&p, x[i]);
}

struct PriorUser {
    const iso_prior *d_prior;
    int which;
};

static int prior_launch(iso_ctx *ctx, cudaStream_t st, void *const *d, int64_t row0, int64_t n, void *user)
{
    (void)row0;
    PriorUser *u = (PriorUser *)user;
    int64_t want = (n + 255) / 256, cap = (int64_t)ctx->prop.multiProcessorCount * 8;
    int blocks = (int)(want < cap ? want : cap);
    iso_prior_eval_kernel<<<blocks, 256, 0, st>>>(u->d_prior, u->which, (const double *)d[0], (double *)d[1], n);
    ctx->launches++;
    ISO_CUDA(ctx, cudaGetLastError());
    return ISO_OK;
}

extern "C" int iso_prior_eval(iso_ctx *ctx, const iso_prior *prior, int which, const double *h_x, int64_t N, double *h_out)
{
    if (!ctx) return iso_set_error(nullptr, ISO_E_INVALID, "iso_prior_eval: ctx is NULL");
    ISO_REQUIRE(ctx, prior && (which == 0 || which == 1) && N >= 0, "iso_prior_eval: bad argument");
    ISO_REQUIRE(ctx, iso_prior_valid(*prior), "iso_prior_eval: unsupported prior kind (no CPU fallback exists)");
    if (N == 0) return ISO_OK;
    ISO_REQUIRE(ctx, h_x && h_out, "iso_prior_eval: NULL buffer");
    IsoDeviceGuard guard(ctx->device);
    iso_prior filled = *prior;
    iso_prior_fill(&filled);
    iso_prior *d_prior = nullptr;
    ISO_CUDA(ctx, iso_scratch_alloc(ctx, (void **)&d_prior, sizeof(iso_prior)));
    cudaError_t e = cudaMemcpy(d_prior, &filled, sizeof(iso_prior), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        iso_scratch_free(ctx, d_prior);
        return iso_check_cuda(ctx, e, "iso_prior_eval");
    }
    IsoPipeArray arr[2] = {IsoPipeArray{h_x, nullptr, 8}, IsoPipeArray{nullptr, h_out, 8}};
    PriorUser u{d_prior, which};
    int rc = iso_run_pipeline(ctx, N, arr, 2, prior_launch, &u);   // returns with every stream of the pipeline waited for
    iso_scratch_free(ctx, d_prior);
    return rc;
}
