// isochrones_b200 — shared host/device definitions of the CUDA library (sm_100a only).
//
// Data layout in HBM:
//   grid   : float64 [n_nodes + ISO_PAD_NODES][ncols]   nodes in the reference's C order (axis 0 slowest), columns
//            innermost, followed by two all-zero padding nodes (target of out-of-array corner reads);
//   axes   : double2 per node = (a[i], 1 / (a[i+1] - a[i]))  (last node: (a[n-1], 0)).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>
#include <string>
#include <vector>

#include "../../include/isochrones_b200.h"

#define ISO_LOG_ONE_OVER_ROOT_2PI (-0.91893853320467274178)
#define ISO_PAD_NODES 2   // all-zero nodes appended to every staged grid
// EEP-pair records of a model pack (what the fused row kernels gather): record r holds the six columns every row
// needs (Teff, logg, feh, Mbol, age|mass, dt_deep|dm_deep) of flat node r AND of flat node r + 1 — the two
// EEP-adjacent corners of a cell — in 96 bytes = 3 sectors, instead of 2 x 64-byte nodes = 4 sectors.
#define ISO_PP_NCOLS 6
#define ISO_PP_STRIDE 12
// Layouts of the model cell the fused row kernels can gather (all hold the same numbers; a batch picks one):
//   ISO_LAYOUT_NODE64 — the 8-column model pack itself: 8 corners x one 64-byte node (16 sectors per cell, 64 B/node);
//   ISO_LAYOUT_PAIR96 — EEP-pair records: 4 x 96 bytes (12 sectors per cell, 96 B/node: every node stored twice);
//   ISO_LAYOUT_NODE48 — the six always-needed columns, 48 bytes per node, no duplication: the two EEP-adjacent nodes
//       of a corner pair are 96 contiguous bytes at a 16-byte-aligned offset = 3 or 4 sectors (14 per cell on
//       average, 48 B/node: the smallest footprint, for batches whose gathers spread beyond the L2).
#define ISO_LAYOUT_NODE64 0
#define ISO_LAYOUT_PAIR96 1
#define ISO_LAYOUT_NODE48 2
#define ISO_N48_PAD_NODES 4   // zero nodes behind the 48-byte-node array (node r + 1 and the 16-byte overhang of an odd pair)

// ------------------------------------------------------------------------------------------------
// host-side objects behind the opaque handles
// ------------------------------------------------------------------------------------------------
struct IsoAxisDev {
    int n;
    int arith;          // 1: a[i] == fma(i, step, a0) bit-exactly for every i (closed-form lookup, no table)
    int off;            // offset of this axis in the concatenated double2 node table
    int pad_;
    double a0, alast;   // first / last node (inclusive bounds, interp.py:106-114)
    double step, inv_step;
};

struct IsoGridDev {     // by-value kernel argument
    const double *g;    // [n_nodes + ISO_PAD_NODES][ncols]
    const double *gp;   // model packs only: EEP-pair records [n_nodes + ISO_PAD_NODES][ISO_PP_STRIDE] (or NULL)
    const double *g48;  // model packs only: 48-byte nodes [n_nodes + ISO_N48_PAD_NODES][ISO_PP_NCOLS] (or NULL)
    const double2 *nodes;   // concatenated axis tables
    long long n_nodes;
    int ndim, ncols;
    int n[ISO_MAX_DIM];
    int nodes_total;
    int pad_;
    IsoAxisDev ax[ISO_MAX_DIM];
};

struct iso_grid {
    IsoGridDev dev;
    double *d_grid = nullptr;
    double2 *d_nodes = nullptr;
    double *d_pair = nullptr;               // EEP-pair records, built on first use by a row kernel (iso_grid_pair_pack)
    double *d_n48 = nullptr;                // 48-byte nodes, built on first use (iso_grid_n48_pack)
    std::vector<double> h_axes[ISO_MAX_DIM];
    int64_t shape[ISO_MAX_DIM + 1];
    int device = 0;
};

#define ISO_CLAIM_SLOTS 3
#define ISO_CLAIM_STRIDE 16   // unsigned long long per slot: [0] next unclaimed row, [1] finished CTAs

struct iso_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;          // compute stream
    cudaStream_t copy_stream[2] = {nullptr, nullptr};
    cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
    cudaEvent_t ev_copy[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
    cudaDeviceProp prop;
    std::string last_error;
    std::recursive_mutex mu;                // entry points that touch the staging buffers / streams hold it
    int64_t launches = 0;
    int *d_small = nullptr;                 // 1024-int scratch for the column lists of the device-buffer entry points
    // row-claim counters of the dynamically scheduled lnpost kernels: per stream slot (compute stream, copy streams)
    // one (claim, done) pair, 128 bytes apart; zero between launches (the last CTA of a launch resets its pair)
    unsigned long long *d_claim = nullptr;
    // staging buffers for the host-pointer entry points (grown on demand)
    void *d_stage[2] = {nullptr, nullptr};
    int64_t d_stage_bytes[2] = {0, 0};
    void *h_stage[2] = {nullptr, nullptr};
    int64_t h_stage_bytes[2] = {0, 0};
    // NCCL (loaded lazily)
    void *nccl_comm = nullptr;
    int nccl_rank = 0, nccl_nranks = 1;
};

int iso_set_error(iso_ctx *ctx, int code, const char *fmt, ...);
int iso_check_cuda(iso_ctx *ctx, cudaError_t e, const char *what);
int iso_stage_reserve(iso_ctx *ctx, int slot, int64_t dev_bytes, int64_t host_bytes);
int iso_grid_pair_pack(iso_ctx *ctx, const iso_grid *model_pack);
int iso_grid_n48_pack(iso_ctx *ctx, const iso_grid *model_pack);

// fused lnpost + all-gather over NVLink peer mappings (iso_peer.cu / iso_lnpost.cu)
#define ISO_MAX_PEERS 8
struct IsoPeerTargets {
    double *out[ISO_MAX_PEERS];   // receive buffer of every rank for this step (peer-mapped device pointers)
    unsigned long long *flags[ISO_MAX_PEERS];   // flag array of every rank (peer-mapped)
    unsigned long long step;      // value the last CTA publishes in slot `rank` of every flag array
    unsigned *done;               // CTA arrival counter (own device memory, zero between launches)
    long long offset;             // this rank's block starts here in every receive buffer
    int n, rank;
    // in-kernel completion wait (ISO_PEER_INKERNEL_WAIT): the last CTA, after publishing, waits for every rank's flag
    const unsigned long long *own_flags;
    unsigned long long timeout_ns;
    unsigned *err;
};
#ifndef ISO_PEER_INKERNEL_WAIT
#define ISO_PEER_INKERNEL_WAIT 1
#endif
struct iso_models;
int iso_lnpost_launch_peers(iso_ctx *ctx, const iso_grid *model_pack, const iso_grid *bc_pack, const iso_models *models,
                            const int32_t *d_model_of_row, const double *d_pars, int64_t N, const IsoPeerTargets *peers);   // builds model_pack->d_pair if absent

// Chunked, double-buffered host<->device pipeline of the host-pointer entry points: alternating chunks run
// H2D -> kernel -> D2H on the two copy streams so the transfers of one chunk overlap the kernel of the other.
// Page-locked caller buffers (iso_host_alloc / cudaHostRegister) are copied directly; pageable ones go through
// the context's pinned staging buffers.
#define ISO_PIPE_MAX_ARRAYS 12
struct IsoPipeArray {
    const void *h_in;     // input array (NULL for outputs)
    void *h_out;          // output array (NULL for inputs; both NULL: optional output not requested)
    int64_t row_bytes;
};
typedef int (*iso_pipe_launch_fn)(iso_ctx *ctx, cudaStream_t stream, void *const *d_arrays, int64_t row0, int64_t n_rows,
                                  void *user);
int iso_run_pipeline(iso_ctx *ctx, int64_t n_rows, const IsoPipeArray *arrays, int n_arrays, iso_pipe_launch_fn launch,
                     void *user);

#define ISO_CUDA(ctx, call)                                            \
    do {                                                               \
        cudaError_t e__ = (call);                                      \
        if (e__ != cudaSuccess) return iso_check_cuda((ctx), e__, #call); \
    } while (0)

#define ISO_REQUIRE(ctx, cond, msg)                                      \
    do {                                                                 \
        if (!(cond)) return iso_set_error((ctx), ISO_E_INVALID, "%s", (msg)); \
    } while (0)

struct IsoDeviceGuard {
    int prev = -1;
    explicit IsoDeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~IsoDeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// ------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ double iso_nan() { return __longlong_as_double(0x7ff8000000000000LL); }
__device__ __forceinline__ double iso_neg_inf() { return __longlong_as_double(0xfff0000000000000LL); }

// Locate x on one axis: returns the cell index and the normalised distance y exactly as
// find_indices_* (interp.py:63-205) does after searchsorted (interp.py:10-35):
//   exact node hit -> (node, 0);  otherwise (L - 1, (x - a[L-1]) / (a[L] - a[L-1])).
// The caller has already established a0 <= x <= alast (inclusive, NaN excluded).  The division is a
// multiplication by the host-computed reciprocal of the cell width (<= 1.5 ulp on y).
// `nodes` may point to shared or global memory.
__device__ __forceinline__ int iso_axis_locate(const IsoAxisDev &ax, const double2 *__restrict__ nodes, double x,
                                               double &y)
{
    if (ax.arith) {
        int i = (int)((x - ax.a0) * ax.inv_step);
        i = max(0, min(i, ax.n - 1));
        double ai = fma((double)i, ax.step, ax.a0);
        while (ai > x) {            // rounding of the closed form can be off by one
            --i;
            ai = fma((double)i, ax.step, ax.a0);
        }
        while (i + 1 < ax.n) {
            double an = fma((double)(i + 1), ax.step, ax.a0);
            if (an > x) break;
            ++i;
            ai = an;
        }
        y = (i + 1 < ax.n) ? (x - ai) * ax.inv_step : 0.0;
        return i;
    }
    int lo = 0, len = ax.n;          // largest i with a[i] <= x
    while (len > 1) {
        int half = len >> 1;
        lo = (nodes[lo + half].x <= x) ? lo + half : lo;
        len -= half;
    }
    double2 nd = nodes[lo];
    y = (x - nd.x) * nd.y;           // nd.y == 0 on the last node; x == nd.x on an exact hit
    return lo;
}

__device__ __forceinline__ bool iso_in_bounds(const IsoAxisDev &ax, double x)
{
    return (x >= ax.a0) && (x <= ax.alast);      // false for NaN (interp.py:254, 106-114)
}

// Locate a point on every axis of a grid.  Returns false when any coordinate is NaN (interp.py:254) or outside
// the inclusive axis range (interp.py:106-114); idx / y are then undefined.
template <int NDIM>
__device__ __forceinline__ bool iso_locate(const IsoGridDev &g, const double2 *__restrict__ nodes,
                                           const double (&x)[NDIM], int (&idx)[NDIM], double (&y)[NDIM])
{
    bool ok = true;
#pragma unroll
    for (int d = 0; d < NDIM; d++) ok = ok && iso_in_bounds(g.ax[d], x[d]);
    if (!ok) return false;
#pragma unroll
    for (int d = 0; d < NDIM; d++) idx[d] = iso_axis_locate(g.ax[d], nodes + g.ax[d].off, x[d], y[d]);
    return true;
}

// Flat node index and weight of the 2^NDIM corners, in the reference's order: corner j, dim 0 = most
// significant bit, weight = product over dims (in dim order) of (1 - y) or y (interp.py:266-284).  The flat
// index follows the reference's unchecked stride arithmetic (an index + 1 equal to the axis length lands on the
// next row); an index beyond the whole array is redirected to the all-zero padding node n_nodes.
template <int NDIM>
__device__ __forceinline__ void iso_corners(const IsoGridDev &g, const int (&idx)[NDIM], const double (&y)[NDIM],
                                            unsigned (&node)[1 << NDIM], double (&w)[1 << NDIM])
{
#pragma unroll
    for (int j = 0; j < (1 << NDIM); j++) {
        unsigned nd = 0;
        double wt = 1.0;
#pragma unroll
        for (int d = 0; d < NDIM; d++) {
            int bit = (j >> (NDIM - 1 - d)) & 1;
            nd = nd * (unsigned)g.n[d] + (unsigned)(idx[d] + bit);
            wt *= bit ? y[d] : (1.0 - y[d]);
        }
        node[j] = min(nd, (unsigned)g.n_nodes);
        w[j] = wt;
    }
}

struct __align__(32) iso_d4 {
    double x, y, z, w;
};

// 256-bit read-only global load (LDG.E.256 on sm_100a): one 32-byte sector per instruction.
__device__ __forceinline__ iso_d4 iso_ldg256(const double *p)
{
    iso_d4 r;
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
    return r;
}

#endif  // __CUDACC__
