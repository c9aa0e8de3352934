// isochrones_b200 — staging of the dense model / bolometric-correction grids in HBM.
//
// Replaces the host-resident DFInterpolator.grid / .index_columns of the reference (interp.py:571-614).
// Layout: float64 [n_nodes + ISO_PAD_NODES][ncols], nodes in the reference's C order, columns
// innermost; the trailing padding nodes are all-zero and are the target of corner reads that the reference's
// unchecked indexing would send beyond the end of its array (interp.py:266-291; weight is always 0 there).
#include <math.h>
#include <string.h>

#include "iso_common.cuh"

// ncols_out columns per node; cols[c] < 0 (or c >= ncols_sel) gives a zero column
__global__ void iso_repack_kernel(const double *__restrict__ src, int src_ncols, long long n_nodes, double *__restrict__ dst,
                                  int ncols_out, const int *__restrict__ cols, int ncols_sel)
{
    long long total = (n_nodes + ISO_PAD_NODES) * (long long)ncols_out;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        long long node = t / ncols_out;
        int c = (int)(t - node * ncols_out);
        double v = 0.0;
        if (node < n_nodes && c < ncols_sel) {
            int sc = cols[c];
            if (sc >= 0) v = src[node * src_ncols + sc];
        }
        dst[t] = v;
    }
}

// EEP-pair records of an 8-column model pack (ISO_PP_*): record r = columns 0..5 of node r, then of node r + 1.
// Node r + 1 follows the reference's flat (unchecked) index arithmetic: past the last node it is the all-zero
// padding node, and the padding records themselves are all-zero.
__global__ void iso_pair_pack_kernel(const double *__restrict__ src, long long n_nodes, double *__restrict__ dst)
{
    const long long total = (n_nodes + ISO_PAD_NODES) * (long long)ISO_PP_STRIDE;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long rec = t / ISO_PP_STRIDE;
        const int r = (int)(t - rec * ISO_PP_STRIDE);
        const int half = r / ISO_PP_NCOLS, c = r - half * ISO_PP_NCOLS;
        dst[t] = rec < n_nodes ? src[(rec + half) * ISO_MP_NCOLS + c] : 0.0;   // rec + 1 <= n_nodes: a staged (zero) node
    }
}

int iso_grid_pair_pack(iso_ctx *ctx, const iso_grid *model_pack)
{
    iso_grid *g = const_cast<iso_grid *>(model_pack);   // a cache inside the handle; callers hold the context lock
    if (g->d_pair) return ISO_OK;
    ISO_REQUIRE(ctx, g->dev.ndim == 3 && g->dev.ncols == ISO_MP_NCOLS, "pair pack: not a model pack");
    IsoDeviceGuard guard(g->device);
    const size_t bytes = (size_t)(g->dev.n_nodes + ISO_PAD_NODES) * ISO_PP_STRIDE * sizeof(double);
    double *d = nullptr;
    ISO_CUDA(ctx, cudaMalloc(&d, bytes));
    iso_pair_pack_kernel<<<ctx->prop.multiProcessorCount * 8, 256, 0, ctx->stream>>>(g->d_grid, g->dev.n_nodes, d);
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        cudaFree(d);
        return iso_check_cuda(ctx, e, "iso_grid_pair_pack");
    }
    g->d_pair = d;
    g->dev.gp = d;
    return ISO_OK;
}

// 48-byte nodes of an 8-column model pack: columns 0..5 of every node, followed by ISO_N48_PAD_NODES zero nodes
__global__ void iso_n48_pack_kernel(const double *__restrict__ src, long long n_nodes, double *__restrict__ dst)
{
    const long long total = (n_nodes + ISO_N48_PAD_NODES) * (long long)ISO_PP_NCOLS;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long node = t / ISO_PP_NCOLS;
        const int c = (int)(t - node * ISO_PP_NCOLS);
        dst[t] = node < n_nodes ? src[node * ISO_MP_NCOLS + c] : 0.0;
    }
}

int iso_grid_n48_pack(iso_ctx *ctx, const iso_grid *model_pack)
{
    iso_grid *g = const_cast<iso_grid *>(model_pack);   // a cache inside the handle; callers hold the context lock
    if (g->d_n48) return ISO_OK;
    ISO_REQUIRE(ctx, g->dev.ndim == 3 && g->dev.ncols == ISO_MP_NCOLS, "48-byte node pack: not a model pack");
    IsoDeviceGuard guard(g->device);
    const size_t bytes = (size_t)(g->dev.n_nodes + ISO_N48_PAD_NODES) * ISO_PP_NCOLS * sizeof(double);
    double *d = nullptr;
    ISO_CUDA(ctx, cudaMalloc(&d, bytes));
    iso_n48_pack_kernel<<<ctx->prop.multiProcessorCount * 8, 256, 0, ctx->stream>>>(g->d_grid, g->dev.n_nodes, d);
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        cudaFree(d);
        return iso_check_cuda(ctx, e, "iso_grid_n48_pack");
    }
    g->d_n48 = d;
    g->dev.g48 = d;
    return ISO_OK;
}

static int fill_axes(iso_ctx *ctx, iso_grid *g)
{
    // concatenated (a[i], 1 / (a[i+1] - a[i])) table + closed-form descriptors
    std::vector<double2> nodes;
    for (int d = 0; d < g->dev.ndim; d++) {
        const std::vector<double> &a = g->h_axes[d];
        IsoAxisDev &ax = g->dev.ax[d];
        int n = (int)a.size();
        ax.n = n;
        ax.off = (int)nodes.size();
        ax.pad_ = 0;
        ax.a0 = a[0];
        ax.alast = a[n - 1];
        ax.step = n > 1 ? a[1] - a[0] : 1.0;
        ax.inv_step = 1.0 / ax.step;
        int arith = n > 1 ? 1 : 0;
        for (int i = 0; i < n && arith; i++)
            if (fma((double)i, ax.step, ax.a0) != a[i]) arith = 0;
        ax.arith = arith;
        for (int i = 0; i < n; i++) {
            double2 nd;
            nd.x = a[i];
            nd.y = (i + 1 < n) ? 1.0 / (a[i + 1] - a[i]) : 0.0;
            nodes.push_back(nd);
        }
        g->dev.n[d] = n;
    }
    for (int d = g->dev.ndim; d < ISO_MAX_DIM; d++) {
        memset(&g->dev.ax[d], 0, sizeof(IsoAxisDev));
        g->dev.n[d] = 1;
    }
    g->dev.nodes_total = (int)nodes.size();
    ISO_CUDA(ctx, cudaMalloc(&g->d_nodes, nodes.size() * sizeof(double2)));
    ISO_CUDA(ctx, cudaMemcpyAsync(g->d_nodes, nodes.data(), nodes.size() * sizeof(double2), cudaMemcpyHostToDevice,
                                  ctx->stream));
    ISO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    g->dev.nodes = g->d_nodes;
    return ISO_OK;
}

static void free_grid(iso_grid *g)
{
    if (!g) return;
    if (g->d_grid) cudaFree(g->d_grid);
    if (g->d_nodes) cudaFree(g->d_nodes);
    if (g->d_pair) cudaFree(g->d_pair);
    if (g->d_n48) cudaFree(g->d_n48);
    delete g;
}

extern "C" {

int iso_grid_stage(iso_ctx *ctx, const double *h_grid, int ndim, const int64_t *shape, const double *const *h_axes,
                   iso_grid **out)
{
    if (!ctx) return iso_set_error(nullptr, ISO_E_INVALID, "iso_grid_stage: ctx is NULL");
    ISO_REQUIRE(ctx, out && h_grid && shape && h_axes, "iso_grid_stage: NULL argument");
    ISO_REQUIRE(ctx, ndim >= 2 && ndim <= ISO_MAX_DIM, "iso_grid_stage: ndim must be 2, 3 or 4");
    *out = nullptr;
    long long n_nodes = 1;
    for (int d = 0; d < ndim; d++) {
        ISO_REQUIRE(ctx, shape[d] >= 1 && shape[d] < (1 << 30), "iso_grid_stage: bad axis length");
        ISO_REQUIRE(ctx, h_axes[d] != nullptr, "iso_grid_stage: NULL axis");
        n_nodes *= shape[d];
    }
    int64_t ncols = shape[ndim];
    ISO_REQUIRE(ctx, ncols >= 1 && ncols <= 4096, "iso_grid_stage: bad column count");
    ISO_REQUIRE(ctx, n_nodes + ISO_PAD_NODES < (1LL << 31), "iso_grid_stage: more than 2^31 nodes");
    for (int d = 0; d < ndim; d++) {
        for (int64_t i = 0; i < shape[d]; i++) {
            double v = h_axes[d][i];
            // a NaN node would make the reference's searchsorted loop forever (interp.py:10-35)
            ISO_REQUIRE(ctx, v == v, "iso_grid_stage: NaN in an axis");
            ISO_REQUIRE(ctx, i == 0 || h_axes[d][i - 1] < v, "iso_grid_stage: axis is not strictly increasing");
        }
    }
    IsoDeviceGuard guard(ctx->device);
    iso_grid *g = new iso_grid();
    g->device = ctx->device;
    memset(&g->dev, 0, sizeof(g->dev));
    g->dev.ndim = ndim;
    g->dev.ncols = (int)ncols;
    g->dev.n_nodes = n_nodes;
    for (int d = 0; d <= ISO_MAX_DIM; d++) g->shape[d] = 0;
    for (int d = 0; d < ndim; d++) {
        g->shape[d] = shape[d];
        g->h_axes[d].assign(h_axes[d], h_axes[d] + shape[d]);
    }
    g->shape[ndim] = ncols;
    size_t bytes = (size_t)(n_nodes + ISO_PAD_NODES) * ncols * sizeof(double);
    cudaError_t e = cudaMalloc(&g->d_grid, bytes);
    if (e != cudaSuccess) {
        free_grid(g);
        return iso_check_cuda(ctx, e, "cudaMalloc(grid)");
    }
    size_t body = (size_t)n_nodes * ncols * sizeof(double);
    e = cudaMemcpyAsync(g->d_grid, h_grid, body, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync((char *)g->d_grid + body, 0, bytes - body, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        free_grid(g);
        return iso_check_cuda(ctx, e, "copy grid to device");
    }
    g->dev.g = g->d_grid;
    int rc = fill_axes(ctx, g);
    if (rc != ISO_OK) {
        free_grid(g);
        return rc;
    }
    *out = g;
    return ISO_OK;
}

int iso_grid_repack(iso_ctx *ctx, const iso_grid *src, const int32_t *cols, int ncols, int ncols_out, iso_grid **out)
{
    if (!ctx) return iso_set_error(nullptr, ISO_E_INVALID, "iso_grid_repack: ctx is NULL");
    ISO_REQUIRE(ctx, src && cols && out, "iso_grid_repack: NULL argument");
    ISO_REQUIRE(ctx, src->device == ctx->device, "iso_grid_repack: grid belongs to another device");
    ISO_REQUIRE(ctx, ncols >= 1 && ncols_out >= ncols && ncols_out <= 4096, "iso_grid_repack: bad column counts");
    for (int c = 0; c < ncols; c++)
        ISO_REQUIRE(ctx, cols[c] < src->dev.ncols, "iso_grid_repack: column index out of range");
    *out = nullptr;
    IsoDeviceGuard guard(ctx->device);
    iso_grid *g = new iso_grid();
    g->device = ctx->device;
    g->dev = src->dev;
    g->dev.ncols = ncols_out;
    g->dev.g = nullptr;
    g->dev.gp = nullptr;
    g->dev.g48 = nullptr;
    g->dev.nodes = nullptr;
    for (int d = 0; d <= ISO_MAX_DIM; d++) g->shape[d] = src->shape[d];
    g->shape[src->dev.ndim] = ncols_out;
    for (int d = 0; d < src->dev.ndim; d++) g->h_axes[d] = src->h_axes[d];
    size_t bytes = (size_t)(src->dev.n_nodes + ISO_PAD_NODES) * ncols_out * sizeof(double);
    int *d_cols = nullptr;
    cudaError_t e = cudaMalloc(&g->d_grid, bytes);
    if (e == cudaSuccess) e = cudaMalloc(&d_cols, ncols * sizeof(int));
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_cols, cols, ncols * sizeof(int), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        int blocks = ctx->prop.multiProcessorCount * 8;
        iso_repack_kernel<<<blocks, 256, 0, ctx->stream>>>(src->d_grid, src->dev.ncols, src->dev.n_nodes, g->d_grid,
                                                           ncols_out, d_cols, ncols);
        ctx->launches++;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (d_cols) cudaFree(d_cols);
    if (e != cudaSuccess) {
        free_grid(g);
        return iso_check_cuda(ctx, e, "iso_grid_repack");
    }
    g->dev.g = g->d_grid;
    int rc = fill_axes(ctx, g);
    if (rc != ISO_OK) {
        free_grid(g);
        return rc;
    }
    *out = g;
    return ISO_OK;
}

int iso_grid_destroy(iso_ctx *ctx, iso_grid *grid)
{
    if (!grid) return ISO_OK;
    IsoDeviceGuard guard(grid->device);
    if (ctx) cudaStreamSynchronize(ctx->stream);
    free_grid(grid);
    return ISO_OK;
}

int iso_grid_shape(const iso_grid *grid, int *ndim, int64_t *shape)
{
    if (!grid || !ndim || !shape) return iso_set_error(nullptr, ISO_E_INVALID, "iso_grid_shape: NULL argument");
    *ndim = grid->dev.ndim;
    for (int d = 0; d <= ISO_MAX_DIM; d++) shape[d] = grid->shape[d];
    return ISO_OK;
}

}  // extern "C"
