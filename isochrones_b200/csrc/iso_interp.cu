// isochrones_b200 — batched multilinear interpolation kernels.
//
//   iso_interp_values  replaces interp_values_2d/3d/4d (interp.py:341-392), i.e. the serial loop over
//                      interp_value_* (:208-338), find_indices_* (:63-205) and searchsorted (:10-35);
//   iso_interp_mags    replaces interp_mags (mags.py:64-124) over interp_mag (mags.py:8-61).
//
// One thread per point.  The 2^ndim corner offsets and weights are computed once per point and kept in
// registers; columns are the outer loop, corners the inner one, so each output accumulates in the reference's
// corner order (zero-weight corners are read and multiplied: NaN * 0 = NaN must propagate, interp.py:286-291).
#include "iso_common.cuh"

#define ISO_INTERP_THREADS 256

struct IsoInterpArgs {
    const double *x[ISO_MAX_DIM];   // device coordinate arrays, each [N]
    const int *icols;               // device [ncols]
    double *out;                    // device [N, ncols]
    long long N;
    int ncols;
};

template <int NDIM>
__global__ void __launch_bounds__(ISO_INTERP_THREADS)
iso_interp_values_kernel(IsoGridDev g, IsoInterpArgs a)
{
    const double nan = iso_nan();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < a.N; i += (long long)gridDim.x * blockDim.x) {
        double x[NDIM], y[NDIM];
        int idx[NDIM];
#pragma unroll
        for (int d = 0; d < NDIM; d++) x[d] = a.x[d][i];
        double *o = a.out + i * a.ncols;
        if (!iso_locate<NDIM>(g, g.nodes, x, idx, y)) {
            for (int c = 0; c < a.ncols; c++) o[c] = nan;
            continue;
        }
        unsigned node[1 << NDIM];
        double w[1 << NDIM];
        iso_corners<NDIM>(g, idx, y, node, w);
        for (int c = 0; c < a.ncols; c++) {
            int ic = __ldg(a.icols + c);
            double acc = 0.0;
#pragma unroll
            for (int j = 0; j < (1 << NDIM); j++) acc = fma(__ldg(g.g + (size_t)node[j] * g.ncols + ic), w[j], acc);
            o[c] = acc;
        }
    }
}

struct IsoMagsArgs {
    const double *pars[5];          // device, parameter-major: pars[j][i]  (mags.py:86-87)
    int index_order[5];
    int i_Teff, i_logg, i_feh, i_Mbol;
    const int *bc_cols;             // device [n_bands]
    int n_bands;
    long long N;
    double *Teff, *logg, *feh, *mags;   // device [N], [N], [N], [N, n_bands]
};

// Packed fast path of interp_mags: the model grid is a model pack whose first four columns are (Teff, logg, feh, Mbol)
// and the BC grid a BC pack whose columns 0 .. n_bands-1 are the requested bands (what the Python mirror always
// passes).  A corner is then ONE 32-byte sector of the model pack and one sector per chunk of 4 bands of the BC pack:
// 8 + 16 vector loads per point instead of 32 + 16 n_bands scalar ones.  Same corner order, same FMAs.
__global__ void __launch_bounds__(ISO_INTERP_THREADS)
iso_interp_mags_packed_kernel(IsoGridDev model, IsoGridDev bc, IsoMagsArgs a)
{
    const double nan = iso_nan();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < a.N; i += (long long)gridDim.x * blockDim.x) {
        double p[5];
#pragma unroll
        for (int j = 0; j < 5; j++) p[j] = a.pars[j][i];
        double q[5];
#pragma unroll
        for (int d = 0; d < 5; d++) {
            int io = a.index_order[d];
            q[d] = io == 0 ? p[0] : io == 1 ? p[1] : io == 2 ? p[2] : io == 3 ? p[3] : p[4];
        }
        double props[4] = {nan, nan, nan, nan};
        {
            double x[3] = {q[0], q[1], q[2]}, y[3];
            int idx[3];
            if (iso_locate<3>(model, model.nodes, x, idx, y)) {
                unsigned node[8];
                double w[8];
                iso_corners<3>(model, idx, y, node, w);
                double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const iso_d4 r = iso_ldg256(model.g + (size_t)node[j] * ISO_MP_NCOLS);
                    v0 = fma(r.x, w[j], v0);
                    v1 = fma(r.y, w[j], v1);
                    v2 = fma(r.z, w[j], v2);
                    v3 = fma(r.w, w[j], v3);
                }
                props[0] = v0;
                props[1] = v1;
                props[2] = v2;
                props[3] = v3;
            }
        }
        a.Teff[i] = props[0];
        a.logg[i] = props[1];
        a.feh[i] = props[2];
        const double dist_mod = 5.0 * log10(q[3] / 10.0);
        double *m = a.mags + i * a.n_bands;
        double x4[4] = {props[0], props[1], props[2], q[4]}, y4[4];
        int idx4[4];
        if (!iso_locate<4>(bc, bc.nodes, x4, idx4, y4)) {
            for (int b = 0; b < a.n_bands; b++) m[b] = nan;
            continue;
        }
        unsigned node[16];
        double w[16];
        iso_corners<4>(bc, idx4, y4, node, w);
        for (int ch = 0; 4 * ch < a.n_bands; ch++) {
            double b0 = 0.0, b1 = 0.0, b2 = 0.0, b3 = 0.0;
#pragma unroll
            for (int j = 0; j < 16; j++) {
                const iso_d4 r = iso_ldg256(bc.g + (size_t)node[j] * bc.ncols + 4 * ch);
                b0 = fma(r.x, w[j], b0);
                b1 = fma(r.y, w[j], b1);
                b2 = fma(r.z, w[j], b2);
                b3 = fma(r.w, w[j], b3);
            }
            const double mb = props[3] + dist_mod;
            const double out4[4] = {mb - b0, mb - b1, mb - b2, mb - b3};
#pragma unroll
            for (int b = 0; b < 4; b++)
                if (4 * ch + b < a.n_bands) m[4 * ch + b] = out4[b];
        }
    }
}

__global__ void __launch_bounds__(ISO_INTERP_THREADS)
iso_interp_mags_kernel(IsoGridDev model, IsoGridDev bc, IsoMagsArgs a)
{
    const double nan = iso_nan();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < a.N; i += (long long)gridDim.x * blockDim.x) {
        double p[5];
#pragma unroll
        for (int j = 0; j < 5; j++) p[j] = a.pars[j][i];
        double q[5];   // q[d] = pars[index_order[d]]
#pragma unroll
        for (int d = 0; d < 5; d++) {
            int io = a.index_order[d];
            q[d] = io == 0 ? p[0] : io == 1 ? p[1] : io == 2 ? p[2] : io == 3 ? p[3] : p[4];
        }
        // star_props = interp_value_3d(..., [i_Teff, i_logg, i_feh, i_Mbol])   mags.py:35-47
        double props[4] = {nan, nan, nan, nan};
        {
            double x[3] = {q[0], q[1], q[2]}, y[3];
            int idx[3];
            if (iso_locate<3>(model, model.nodes, x, idx, y)) {
                unsigned node[8];
                double w[8];
                iso_corners<3>(model, idx, y, node, w);
                const int cols[4] = {a.i_Teff, a.i_logg, a.i_feh, a.i_Mbol};
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    double acc = 0.0;
#pragma unroll
                    for (int j = 0; j < 8; j++) acc = fma(__ldg(model.g + (size_t)node[j] * model.ncols + cols[c]), w[j], acc);
                    props[c] = acc;
                }
            }
        }
        a.Teff[i] = props[0];
        a.logg[i] = props[1];
        a.feh[i] = props[2];
        // bc = interp_value_4d(Teff, logg, feh, AV, ...)   mags.py:49-50 — the BC lookup uses the interpolated
        // surface feh; mags = Mbol + 5 log10(d / 10) - bc   mags.py:52-59
        double dist_mod = 5.0 * log10(q[3] / 10.0);
        double *m = a.mags + i * a.n_bands;
        double x4[4] = {props[0], props[1], props[2], q[4]}, y4[4];
        int idx4[4];
        if (!iso_locate<4>(bc, bc.nodes, x4, idx4, y4)) {
            for (int b = 0; b < a.n_bands; b++) m[b] = nan;
            continue;
        }
        unsigned node[16];
        double w[16];
        iso_corners<4>(bc, idx4, y4, node, w);
        for (int b = 0; b < a.n_bands; b++) {
            int ic = __ldg(a.bc_cols + b);
            double acc = 0.0;
#pragma unroll
            for (int j = 0; j < 16; j++) acc = fma(__ldg(bc.g + (size_t)node[j] * bc.ncols + ic), w[j], acc);
            m[b] = props[3] + dist_mod - acc;
        }
    }
}

// interp_eeps (interp.py:488-499) over interp_eep (:502-558): (age, feh, mass) -> EEP on an evolution-track grid.
// The per-track age arrays of the reference (StellarModelGrid.get_array_grids, models.py:171-205) are the `i_age`
// column of the staged (feh, mass, eep) grid; lengths[t] is the number of leading populated EEPs of track t.
struct IsoEepArgs {
    const double *age, *feh, *mass;   // device [N]
    const int *lengths;               // device [n_feh * n_mass]
    double *out;                      // device [N]
    long long N;
    int i_age;
};

__global__ void __launch_bounds__(ISO_INTERP_THREADS) iso_interp_eeps_kernel(IsoGridDev g, IsoEepArgs a)
{
    const double nan = iso_nan();
    const int n0 = g.n[0], n1 = g.n[1], n_eep = g.n[2];
    const long long n_tracks = (long long)n0 * n1;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < a.N; i += (long long)gridDim.x * blockDim.x) {
        const double x = a.age[i], x0 = a.feh[i], x1 = a.mass[i];
        if (x != x || !iso_in_bounds(g.ax[0], x0) || !iso_in_bounds(g.ax[1], x1)) {   // NaN in / out of bounds -> NaN
            a.out[i] = nan;
            continue;
        }
        double d0, d1;
        const int i0 = iso_axis_locate(g.ax[0], g.nodes + g.ax[0].off, x0, d0);
        const int i1 = iso_axis_locate(g.ax[1], g.nodes + g.ax[1].off, x1, d1);
        // the reference's unchecked track indices: 00, 01, 10, 11 (interp.py:515-518)
        const long long ind[4] = {(long long)i0 * n1 + i1, (long long)i0 * n1 + (i1 + 1), (long long)(i0 + 1) * n1 + i1,
                                  (long long)(i0 + 1) * n1 + (i1 + 1)};
        double eep[4];
        int ie[4], len[4];
        bool bad = false;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            len[k] = ind[k] < n_tracks ? a.lengths[ind[k]] : 0;   // a track beyond the array (weight 0) counts as empty
            const double *arr = g.g + (size_t)ind[k] * n_eep * g.ncols + a.i_age;
            int lo = 0, n = len[k];                               // searchsorted: number of ages < x (an exact hit returns its index)
            while (n > 0) {
                const int half = n >> 1;
                const bool less = __ldg(arr + (size_t)(lo + half) * g.ncols) < x;
                lo = less ? lo + half + 1 : lo;
                n = less ? n - half - 1 : half;
            }
            ie[k] = lo;
            bad = bad || (lo > n_eep - 1);                        // max_i_eep (interp.py:526-528)
            eep[k] = (double)(lo + 1);                            // EEP = index + 1
        }
        if (bad) {
            a.out[i] = nan;
            continue;
        }
        if (ie[0] >= len[0]) eep[0] = eep[1];                     // sequential, as written (interp.py:540-551)
        if (ie[1] >= len[1]) eep[1] = eep[0];
        if (ie[2] >= len[2]) eep[2] = eep[3];
        if (ie[3] >= len[3]) eep[3] = eep[2];
        const double eep_0 = __dadd_rn(__dmul_rn(1.0 - d1, eep[0]), __dmul_rn(d1, eep[1]));
        const double eep_1 = __dadd_rn(__dmul_rn(1.0 - d1, eep[2]), __dmul_rn(d1, eep[3]));
        a.out[i] = __dadd_rn(__dmul_rn(1.0 - d0, eep_0), __dmul_rn(d0, eep_1));
    }
}

static int grid_blocks(iso_ctx *ctx, int64_t n, int threads)
{
    int64_t want = (n + threads - 1) / threads;
    int64_t cap = (int64_t)ctx->prop.multiProcessorCount * 16;
    return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

struct InterpUser {
    const iso_grid *grid;
    const int *d_icols;
    int ncols;
};

static int interp_launch(iso_ctx *ctx, cudaStream_t st, void *const *d, int64_t row0, int64_t n, void *user)
{
    (void)row0;
    InterpUser *u = (InterpUser *)user;
    IsoInterpArgs a;
    int ndim = u->grid->dev.ndim;
    for (int k = 0; k < ISO_MAX_DIM; k++) a.x[k] = k < ndim ? (const double *)d[k] : nullptr;
    a.icols = u->d_icols;
    a.out = (double *)d[ndim];
    a.N = n;
    a.ncols = u->ncols;
    int blocks = grid_blocks(ctx, n, ISO_INTERP_THREADS);
    if (ndim == 2) iso_interp_values_kernel<2><<<blocks, ISO_INTERP_THREADS, 0, st>>>(u->grid->dev, a);
    else if (ndim == 3) iso_interp_values_kernel<3><<<blocks, ISO_INTERP_THREADS, 0, st>>>(u->grid->dev, a);
    else iso_interp_values_kernel<4><<<blocks, ISO_INTERP_THREADS, 0, st>>>(u->grid->dev, a);
    ctx->launches++;
    ISO_CUDA(ctx, cudaGetLastError());
    return ISO_OK;
}

struct MagsUser {
    const iso_grid *model, *bc;
    IsoMagsArgs proto;
    bool packed = false;   // packed fast path applies (set by mags_is_packed)
};

// the packed layout the fast kernel needs: model pack with (Teff, logg, feh, Mbol) = columns 0..3, BC pack (column
// count a multiple of 4) with the requested bands in columns 0 .. n_bands-1
static bool mags_is_packed(const iso_grid *model, const iso_grid *bc, int i_Teff, int i_logg, int i_feh, int i_Mbol,
                           const int32_t *bc_cols, int n_bands)
{
    if (model->dev.ncols != ISO_MP_NCOLS || i_Teff != ISO_MP_TEFF || i_logg != ISO_MP_LOGG || i_feh != ISO_MP_FEH ||
        i_Mbol != ISO_MP_MBOL)
        return false;
    if (bc->dev.ncols % 4 != 0 || n_bands > bc->dev.ncols) return false;
    for (int b = 0; b < n_bands; b++)
        if (bc_cols[b] != b) return false;
    return true;
}

static int mags_launch(iso_ctx *ctx, cudaStream_t st, void *const *d, int64_t row0, int64_t n, void *user)
{
    (void)row0;
    MagsUser *u = (MagsUser *)user;
    IsoMagsArgs a = u->proto;
    for (int j = 0; j < 5; j++) a.pars[j] = (const double *)d[j];
    a.Teff = (double *)d[5];
    a.logg = (double *)d[6];
    a.feh = (double *)d[7];
    a.mags = (double *)d[8];
    a.N = n;
    int blocks = grid_blocks(ctx, n, ISO_INTERP_THREADS);
    if (u->packed) iso_interp_mags_packed_kernel<<<blocks, ISO_INTERP_THREADS, 0, st>>>(u->model->dev, u->bc->dev, a);
    else iso_interp_mags_kernel<<<blocks, ISO_INTERP_THREADS, 0, st>>>(u->model->dev, u->bc->dev, a);
    ctx->launches++;
    ISO_CUDA(ctx, cudaGetLastError());
    return ISO_OK;
}

struct EepUser {
    const iso_grid *grid;
    const int *d_lengths;
    int i_age;
};

static int eep_launch(iso_ctx *ctx, cudaStream_t st, void *const *d, int64_t row0, int64_t n, void *user)
{
    (void)row0;
    EepUser *u = (EepUser *)user;
    IsoEepArgs a;
    a.age = (const double *)d[0];
    a.feh = (const double *)d[1];
    a.mass = (const double *)d[2];
    a.out = (double *)d[3];
    a.lengths = u->d_lengths;
    a.N = n;
    a.i_age = u->i_age;
    int blocks = grid_blocks(ctx, n, ISO_INTERP_THREADS);
    iso_interp_eeps_kernel<<<blocks, ISO_INTERP_THREADS, 0, st>>>(u->grid->dev, a);
    ctx->launches++;
    ISO_CUDA(ctx, cudaGetLastError());
    return ISO_OK;
}

// small device array of ints that lives for the duration of one host call
struct DevInts {
    int *d = nullptr;
    ~DevInts() { if (d) cudaFree(d); }
    int upload(iso_ctx *ctx, const int32_t *h, int n)
    {
        ISO_CUDA(ctx, cudaMalloc(&d, sizeof(int) * (n > 0 ? n : 1)));
        if (n > 0) ISO_CUDA(ctx, cudaMemcpy(d, h, sizeof(int) * n, cudaMemcpyHostToDevice));
        return ISO_OK;
    }
};

extern "C" {

int iso_interp_values(iso_ctx *ctx, const iso_grid *grid, const double *const *h_x, int64_t N, const int32_t *icols,
                      int ncols, double *h_out)
{
    if (!ctx) return iso_set_error(nullptr, ISO_E_INVALID, "iso_interp_values: ctx is NULL");
    ISO_REQUIRE(ctx, grid && h_x && icols, "iso_interp_values: NULL argument");
    ISO_REQUIRE(ctx, grid->device == ctx->device, "iso_interp_values: grid belongs to another device");
    ISO_REQUIRE(ctx, N >= 0 && ncols >= 1, "iso_interp_values: bad sizes");
    ISO_REQUIRE(ctx, N == 0 || h_out, "iso_interp_values: out is NULL");
    for (int c = 0; c < ncols; c++)
        ISO_REQUIRE(ctx, icols[c] >= 0 && icols[c] < grid->dev.ncols, "iso_interp_values: column index out of range");
    int ndim = grid->dev.ndim;
    for (int d = 0; d < ndim; d++) ISO_REQUIRE(ctx, N == 0 || h_x[d], "iso_interp_values: NULL coordinate array");
    if (N == 0) return ISO_OK;
    IsoDeviceGuard guard(ctx->device);
    DevInts cols;
    int rc = cols.upload(ctx, icols, ncols);
    if (rc != ISO_OK) return rc;
    IsoPipeArray arr[ISO_MAX_DIM + 1];
    for (int d = 0; d < ndim; d++) arr[d] = IsoPipeArray{h_x[d], nullptr, 8};
    arr[ndim] = IsoPipeArray{nullptr, h_out, (int64_t)8 * ncols};
    InterpUser u{grid, cols.d, ncols};
    return iso_run_pipeline(ctx, N, arr, ndim + 1, interp_launch, &u);
}

int iso_interp_mags_cols(iso_ctx *ctx, const iso_grid *model, const iso_grid *bc, const int32_t index_order[5], int i_Teff,
                         int i_logg, int i_feh, int i_Mbol, const int32_t *bc_cols, int n_bands,
                         const double *const *h_par /* 5 pointers, each [N] */, int64_t N, double *h_Teff, double *h_logg,
                         double *h_feh, double *h_mags)
{
    if (!ctx) return iso_set_error(nullptr, ISO_E_INVALID, "iso_interp_mags: ctx is NULL");
    ISO_REQUIRE(ctx, model && bc && index_order, "iso_interp_mags: NULL argument");
    ISO_REQUIRE(ctx, model->device == ctx->device && bc->device == ctx->device, "iso_interp_mags: grid on another device");
    ISO_REQUIRE(ctx, model->dev.ndim == 3 && bc->dev.ndim == 4, "iso_interp_mags: model grid must be 3-D, BC grid 4-D");
    ISO_REQUIRE(ctx, N >= 0 && n_bands >= 0, "iso_interp_mags: bad sizes");
    ISO_REQUIRE(ctx, n_bands == 0 || bc_cols, "iso_interp_mags: bc_cols is NULL");
    int mc = model->dev.ncols;
    ISO_REQUIRE(ctx, i_Teff >= 0 && i_Teff < mc && i_logg >= 0 && i_logg < mc && i_feh >= 0 && i_feh < mc && i_Mbol >= 0 &&
                         i_Mbol < mc, "iso_interp_mags: model column index out of range");
    for (int b = 0; b < n_bands; b++)
        ISO_REQUIRE(ctx, bc_cols[b] >= 0 && bc_cols[b] < bc->dev.ncols, "iso_interp_mags: band column out of range");
    for (int j = 0; j < 5; j++)
        ISO_REQUIRE(ctx, index_order[j] >= 0 && index_order[j] < 5, "iso_interp_mags: bad index_order");
    if (N == 0) return ISO_OK;
    ISO_REQUIRE(ctx, h_par && h_Teff && h_logg && h_feh && (n_bands == 0 || h_mags), "iso_interp_mags: NULL buffer");
    for (int j = 0; j < 5; j++) ISO_REQUIRE(ctx, h_par[j], "iso_interp_mags: NULL parameter array");
    IsoDeviceGuard guard(ctx->device);
    DevInts cols;
    int rc = cols.upload(ctx, bc_cols, n_bands);
    if (rc != ISO_OK) return rc;
    IsoPipeArray arr[9];
    for (int j = 0; j < 5; j++) arr[j] = IsoPipeArray{h_par[j], nullptr, 8};
    arr[5] = IsoPipeArray{nullptr, h_Teff, 8};
    arr[6] = IsoPipeArray{nullptr, h_logg, 8};
    arr[7] = IsoPipeArray{nullptr, h_feh, 8};
    arr[8] = IsoPipeArray{nullptr, n_bands ? h_mags : nullptr, (int64_t)8 * (n_bands > 0 ? n_bands : 1)};
    MagsUser u;
    u.model = model;
    u.bc = bc;
    for (int j = 0; j < 5; j++) u.proto.index_order[j] = index_order[j];
    u.proto.i_Teff = i_Teff;
    u.proto.i_logg = i_logg;
    u.proto.i_feh = i_feh;
    u.proto.i_Mbol = i_Mbol;
    u.proto.bc_cols = cols.d;
    u.proto.n_bands = n_bands;
    u.packed = mags_is_packed(model, bc, i_Teff, i_logg, i_feh, i_Mbol, bc_cols, n_bands);
    return iso_run_pipeline(ctx, N, arr, 9, mags_launch, &u);
}

int iso_interp_mags(iso_ctx *ctx, const iso_grid *model, const iso_grid *bc, const int32_t index_order[5], int i_Teff,
                    int i_logg, int i_feh, int i_Mbol, const int32_t *bc_cols, int n_bands, const double *h_pars,
                    int64_t N, double *h_Teff, double *h_logg, double *h_feh, double *h_mags)
{
    if (!ctx) return iso_set_error(nullptr, ISO_E_INVALID, "iso_interp_mags: ctx is NULL");
    ISO_REQUIRE(ctx, N <= 0 || h_pars, "iso_interp_mags: NULL buffer");
    const double *par[5];
    for (int j = 0; j < 5; j++) par[j] = h_pars ? h_pars + (size_t)j * (size_t)(N > 0 ? N : 0) : nullptr;
    return iso_interp_mags_cols(ctx, model, bc, index_order, i_Teff, i_logg, i_feh, i_Mbol, bc_cols, n_bands, par, N, h_Teff,
                                h_logg, h_feh, h_mags);
}

// column list of a device-buffer call -> the context's scratch (stream-ordered before the kernel that reads it)
static int small_upload(iso_ctx *ctx, const int32_t *h, int n, int offset, const int **d)
{
    ISO_REQUIRE(ctx, n >= 0 && offset >= 0 && offset + n <= 1024, "device entry point: too many columns");
    if (!ctx->d_small) ISO_CUDA(ctx, cudaMalloc(&ctx->d_small, 1024 * sizeof(int)));
    if (n > 0) ISO_CUDA(ctx, cudaMemcpyAsync(ctx->d_small + offset, h, sizeof(int) * n, cudaMemcpyHostToDevice, ctx->stream));
    *d = ctx->d_small + offset;
    return ISO_OK;
}

int iso_interp_values_device(iso_ctx *ctx, const iso_grid *grid, const double *const *d_x, int64_t N, const int32_t *icols,
                             int ncols, double *d_out)
{
    if (!ctx) return iso_set_error(nullptr, ISO_E_INVALID, "iso_interp_values_device: ctx is NULL");
    ISO_REQUIRE(ctx, grid && d_x && icols, "iso_interp_values_device: NULL argument");
    ISO_REQUIRE(ctx, grid->device == ctx->device, "iso_interp_values_device: grid belongs to another device");
    ISO_REQUIRE(ctx, N >= 0 && ncols >= 1, "iso_interp_values_device: bad sizes");
    for (int c = 0; c < ncols; c++)
        ISO_REQUIRE(ctx, icols[c] >= 0 && icols[c] < grid->dev.ncols, "iso_interp_values_device: column index out of range");
    if (N == 0) return ISO_OK;
    ISO_REQUIRE(ctx, d_out, "iso_interp_values_device: out is NULL");
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    IsoDeviceGuard guard(ctx->device);
    const int *d_cols = nullptr;
    int rc = small_upload(ctx, icols, ncols, 0, &d_cols);
    if (rc != ISO_OK) return rc;
    void *d[ISO_MAX_DIM + 1];
    for (int k = 0; k < grid->dev.ndim; k++) {
        ISO_REQUIRE(ctx, d_x[k], "iso_interp_values_device: NULL coordinate array");
        d[k] = const_cast<double *>(d_x[k]);
    }
    d[grid->dev.ndim] = d_out;
    InterpUser u{grid, d_cols, ncols};
    return interp_launch(ctx, ctx->stream, d, 0, N, &u);
}

int iso_interp_mags_device(iso_ctx *ctx, const iso_grid *model, const iso_grid *bc, const int32_t index_order[5], int i_Teff,
                           int i_logg, int i_feh, int i_Mbol, const int32_t *bc_cols, int n_bands, const double *const *d_par,
                           int64_t N, double *d_Teff, double *d_logg, double *d_feh, double *d_mags)
{
    if (!ctx) return iso_set_error(nullptr, ISO_E_INVALID, "iso_interp_mags_device: ctx is NULL");
    ISO_REQUIRE(ctx, model && bc && index_order && d_par, "iso_interp_mags_device: NULL argument");
    ISO_REQUIRE(ctx, model->device == ctx->device && bc->device == ctx->device, "iso_interp_mags_device: grid on another device");
    ISO_REQUIRE(ctx, model->dev.ndim == 3 && bc->dev.ndim == 4, "iso_interp_mags_device: model grid must be 3-D, BC grid 4-D");
    ISO_REQUIRE(ctx, N >= 0 && n_bands >= 0 && (n_bands == 0 || bc_cols), "iso_interp_mags_device: bad sizes");
    int mc = model->dev.ncols;
    ISO_REQUIRE(ctx, i_Teff >= 0 && i_Teff < mc && i_logg >= 0 && i_logg < mc && i_feh >= 0 && i_feh < mc && i_Mbol >= 0 &&
                         i_Mbol < mc, "iso_interp_mags_device: model column index out of range");
    for (int b = 0; b < n_bands; b++)
        ISO_REQUIRE(ctx, bc_cols[b] >= 0 && bc_cols[b] < bc->dev.ncols, "iso_interp_mags_device: band column out of range");
    for (int j = 0; j < 5; j++)
        ISO_REQUIRE(ctx, index_order[j] >= 0 && index_order[j] < 5 && d_par[j], "iso_interp_mags_device: bad index_order / NULL array");
    if (N == 0) return ISO_OK;
    ISO_REQUIRE(ctx, d_Teff && d_logg && d_feh && (n_bands == 0 || d_mags), "iso_interp_mags_device: NULL buffer");
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);
    IsoDeviceGuard guard(ctx->device);
    const int *d_cols = nullptr;
    int rc = small_upload(ctx, bc_cols, n_bands, 512, &d_cols);
    if (rc != ISO_OK) return rc;
    MagsUser u;
    u.model = model;
    u.bc = bc;
    for (int j = 0; j < 5; j++) u.proto.index_order[j] = index_order[j];
    u.proto.i_Teff = i_Teff;
    u.proto.i_logg = i_logg;
    u.proto.i_feh = i_feh;
    u.proto.i_Mbol = i_Mbol;
    u.proto.bc_cols = d_cols;
    u.proto.n_bands = n_bands;
    u.packed = mags_is_packed(model, bc, i_Teff, i_logg, i_feh, i_Mbol, bc_cols, n_bands);
    void *d[9];
    for (int j = 0; j < 5; j++) d[j] = const_cast<double *>(d_par[j]);
    d[5] = d_Teff;
    d[6] = d_logg;
    d[7] = d_feh;
    d[8] = d_mags;
    return mags_launch(ctx, ctx->stream, d, 0, N, &u);
}

int iso_interp_eeps(iso_ctx *ctx, const iso_grid *track_grid, int i_age, const int32_t *h_lengths, const double *h_age,
                    const double *h_feh, const double *h_mass, int64_t N, double *h_eep)
{
    if (!ctx) return iso_set_error(nullptr, ISO_E_INVALID, "iso_interp_eeps: ctx is NULL");
    ISO_REQUIRE(ctx, track_grid && h_lengths, "iso_interp_eeps: NULL argument");
    ISO_REQUIRE(ctx, track_grid->device == ctx->device, "iso_interp_eeps: grid belongs to another device");
    ISO_REQUIRE(ctx, track_grid->dev.ndim == 3, "iso_interp_eeps: needs a 3-D (feh, mass, eep) grid");
    ISO_REQUIRE(ctx, i_age >= 0 && i_age < track_grid->dev.ncols, "iso_interp_eeps: age column out of range");
    ISO_REQUIRE(ctx, N >= 0, "iso_interp_eeps: negative N");
    const int n_tracks = track_grid->dev.n[0] * track_grid->dev.n[1];
    for (int t = 0; t < n_tracks; t++)
        ISO_REQUIRE(ctx, h_lengths[t] >= 0 && h_lengths[t] <= track_grid->dev.n[2], "iso_interp_eeps: bad track length");
    if (N == 0) return ISO_OK;
    ISO_REQUIRE(ctx, h_age && h_feh && h_mass && h_eep, "iso_interp_eeps: NULL buffer");
    IsoDeviceGuard guard(ctx->device);
    DevInts lengths;
    int rc = lengths.upload(ctx, h_lengths, n_tracks);
    if (rc != ISO_OK) return rc;
    IsoPipeArray arr[4] = {IsoPipeArray{h_age, nullptr, 8}, IsoPipeArray{h_feh, nullptr, 8}, IsoPipeArray{h_mass, nullptr, 8},
                           IsoPipeArray{nullptr, h_eep, 8}};
    EepUser u{track_grid, lengths.d, i_age};
    return iso_run_pipeline(ctx, N, arr, 4, eep_launch, &u);
}

}  // extern "C"
