// isochrones_b200 — the fused lnprior + lnlike + lnpost batch kernel (template) and its launch dispatch.
//
// The kernel is instantiated per (stars, catalog, prior profile, grid kind, fused gather, star-sequential, unit cube):
// about sixty instantiations, compiled in three translation units (iso_lnpost_k1/k2/k3.cu, one per number of stars) so
// that the build runs in parallel; iso_lnpost.cu holds the host side.
#pragma once

#include "iso_lnpost_row.cuh"
#include "iso_philox.cuh"

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
    const unsigned long long *peer_own_flags;   // ISO_PEER_INKERNEL_WAIT: this rank's flag array, timeout and error word
    unsigned long long peer_timeout_ns;
    unsigned *peer_err;
    unsigned long long *claim;   // dynamically scheduled kernels: [0] next unclaimed row beyond the first pass, [1] finished CTAs
    // CUBE kernels: rows are points of the unit hypercube, mapped to parameters by BasicStarModel.mnest_prior
    // (starmodel.py:1637-1640: cube[i] = (hi - lo) * cube[i] + lo, unfused) before they are evaluated
    double cube_lo[ISO_MAX_STARS + 4], cube_w[ISO_MAX_STARS + 4];   // lo, hi - lo per parameter
    double *pars_out;            // [N, ndim] mapped parameters (NULL: not wanted; may alias pars)
    unsigned long long cube_seed;
    long long cube_row0;         // global index of row 0 (the Philox counter of a row is its global index)
    int cube_rng;                // 1: draw the cube point of every row on the device (Philox4x32-10) instead of reading pars
};

// Everything the kernel reads besides the grids and the rows travels in the kernel parameter block (constant
// bank): the grid descriptors, the buffer pointers and — outside catalog mode — the star model itself, so that
// observation values and prior constants are constant-bank operands instead of memory loads.
struct IsoLnpostParams {
    IsoRowGrids G;
    IsoLnpostArgs a;
    IsoModelDev model;   // the single model (unused in catalog mode)
};

#ifndef ISO_LNPOST_DYN_SINGLE
#define ISO_LNPOST_DYN_SINGLE 0
#endif
#ifndef ISO_LNPOST_DYN_MULTI
#define ISO_LNPOST_DYN_MULTI 1
#endif

// Row scheduling of the persistent grid:
//   DYN = 0 — static grid stride, the next row's parameters prefetched one iteration ahead (single-star kernels);
//   DYN = 1 — the first pass is the static one, after it every warp claims 32-row chunks from an atomic counter
//             (a.claim[0]) so that the grid drains together whatever the rows cost; no prefetch registers.  Batches of
//             at most one pass never touch the counters; the last CTA to finish resets them (a.claim[1] counts
//             finished CTAs), so a launch finds them zero.  Binary / triple kernels: their rows are long and the ten
//             prefetch registers are what made them spill.
// Measured on B200 with the 48-byte-node gather, ms per 1e6 rows, static / claiming (profiles/README.md, r2d): posterior-
// like 0.114 / 0.127, catalog 0.168 / 0.183, grid-wide 0.169 / 0.167, prior-like 0.133 / 0.135, binary 0.265 / 0.256,
// prior-like on the isochrone grid 0.191 / 0.181.  Claiming one chunk ahead with the rows prefetched into registers or
// with prefetch.global.L1 was slower than both everywhere (register pressure).
// SEQ: star-sequential evaluation of multi-star models whose BC pack is a single 4-band chunk (iso_lnpost_row.cuh).
// CUBE: rows are unit-cube points (MultiNest's live points; or drawn on the device when a.cube_rng is set).
template <int NSTARS, bool CATALOG, int PROFILE, bool TRACK, bool PEER = false, int SEQ = 0, bool CUBE = false,
          int LAYOUT = ISO_MODEL_LAYOUT, int DYN = (NSTARS == 1 ? ISO_LNPOST_DYN_SINGLE : ISO_LNPOST_DYN_MULTI)>
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
    auto eval_row = [&](long long row, double (&p)[NDIMP]) {
        const IsoModelDev &m = CATALOG ? a.models[a.model_of_row[row]] : P.model;
        if (CUBE) {
            if (a.cube_rng) {   // two Philox blocks give up to eight 53-bit uniforms of row (cube_row0 + row)
                const unsigned long long g = (unsigned long long)(a.cube_row0 + row);
                unsigned r0[4], r1[4], r2[4], r3[4];
                const unsigned k0 = (unsigned)a.cube_seed, k1 = (unsigned)(a.cube_seed >> 32);
                iso_philox4x32_10((unsigned)g, (unsigned)(g >> 32), 0u, 0x43554245u, k0, k1, r0);
                iso_philox4x32_10((unsigned)g, (unsigned)(g >> 32), 1u, 0x43554245u, k0, k1, r1);
                iso_philox4x32_10((unsigned)g, (unsigned)(g >> 32), 2u, 0x43554245u, k0, k1, r2);
                iso_philox4x32_10((unsigned)g, (unsigned)(g >> 32), 3u, 0x43554245u, k0, k1, r3);
                const double u[8] = {iso_u01(r0[0], r0[1]), iso_u01(r0[2], r0[3]), iso_u01(r1[0], r1[1]), iso_u01(r1[2], r1[3]),
                                     iso_u01(r2[0], r2[1]), iso_u01(r2[2], r2[3]), iso_u01(r3[0], r3[1]), iso_u01(r3[2], r3[3])};
#pragma unroll
                for (int j = 0; j < NDIMP; j++) p[j] = u[j];
            }
#pragma unroll
            for (int j = 0; j < NDIMP; j++) p[j] = __dadd_rn(__dmul_rn(a.cube_w[j], p[j]), a.cube_lo[j]);
            if (a.pars_out) {
#pragma unroll
                for (int j = 0; j < NDIMP; j++) a.pars_out[row * NDIMP + j] = p[j];
            }
        }
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
        const bool have_rows = !CUBE || !a.cube_rng;
        if (i < a.N && have_rows) {
#pragma unroll
            for (int j = 0; j < NDIMP; j++) pn[j] = a.pars[i * NDIMP + j];
        }
#endif
        for (; i < a.N; i += stride) {
            double p[NDIMP];
#if ISO_LNPOST_PREFETCH
#pragma unroll
            for (int j = 0; j < NDIMP; j++) p[j] = pn[j];
            if (i + stride < a.N && have_rows) {
#pragma unroll
                for (int j = 0; j < NDIMP; j++) pn[j] = a.pars[(i + stride) * NDIMP + j];
            }
#else
#pragma unroll
            for (int j = 0; j < NDIMP; j++) p[j] = (!CUBE || !a.cube_rng) ? a.pars[i * NDIMP + j] : 0.0;
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
                for (int j = 0; j < NDIMP; j++) p[j] = (!CUBE || !a.cube_rng) ? a.pars[row * NDIMP + j] : 0.0;
                eval_row(row, p);
            }
            if (a.N <= stride) break;   // single pass: nothing to claim, the counters stay untouched
            unsigned long long got = 0;
            if (lane == 0) got = atomicAdd(a.claim, 32ULL);
            base = stride + (long long)__shfl_sync(0xffffffffu, got, 0);
        }
    }
    if (DYN != 0 && !PEER && a.N > stride) {
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
        // The CTA barrier orders every thread's peer stores before the system-scope fences of warp 0 (causality is
        // cumulative: the pattern of a grid-wide barrier).  The last CTA to arrive raises this rank's flag on every rank
        // — lane q serves rank q: ONE fence, then relaxed stores that travel side by side (eight release stores in a
        // row would each wait for the previous one's round trip over NVLink) — and, ISO_PEER_INKERNEL_WAIT, waits right
        // here until every rank has raised its flag on ours, so the exchange step is one launch.
        __syncthreads();
        if (threadIdx.x < 32) {
            const int lane = threadIdx.x;
            unsigned last = 0;
            if (lane == 0) {
                __threadfence_system();
                last = atomicAdd(a.peer_done, 1u) == gridDim.x - 1;
                if (last) {
                    *a.peer_done = 0;
                    if (DYN != 0 && a.N > stride) {
                        a.claim[0] = 0;
                        a.claim[1] = 0;
                    }
                }
            }
            last = __shfl_sync(0xffffffffu, last, 0);
            if (last) {
                __threadfence_system();
                if (lane < a.n_peers)
                    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(a.peer_flags[lane] + a.peer_rank), "l"(a.peer_step)
                                 : "memory");
#if ISO_PEER_INKERNEL_WAIT
                if (a.peer_own_flags && lane < a.n_peers) {
                    unsigned long long t0, now, v;
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
                    unsigned spins = 0;
                    for (;;) {
                        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(a.peer_own_flags + lane) : "memory");
                        if (v >= a.peer_step) break;
                        if ((++spins & 255u) == 0) {
                            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                            if (now - t0 > a.peer_timeout_ns) {   // bounded: report the missing rank, do not hang the stream
                                atomicCAS_system(a.peer_err, 0u, (unsigned)(lane + 1) | ((unsigned)a.peer_step << 8));
                                break;
                            }
                        }
                    }
                }
                __syncwarp();
                __threadfence_system();
#endif
            }
        }
    }
}


// what a launch needs to know to pick its instantiation
struct IsoLnpostFlags {
    bool catalog, profile_default, track, peer, cube;
    int seq;   // multi-star models: number of 4-band chunks of the BC pack (star-sequential kernels), else 0
};

// launch dispatch of one translation unit (NS stars)
int iso_lnpost_dispatch_1(iso_ctx *ctx, cudaStream_t st, const IsoLnpostParams &P, size_t smem, const IsoLnpostFlags &f);
int iso_lnpost_dispatch_2(iso_ctx *ctx, cudaStream_t st, const IsoLnpostParams &P, size_t smem, const IsoLnpostFlags &f);
int iso_lnpost_dispatch_3(iso_ctx *ctx, cudaStream_t st, const IsoLnpostParams &P, size_t smem, const IsoLnpostFlags &f);

#ifdef __CUDACC__
template <int NS>
static int iso_lnpost_dispatch(iso_ctx *ctx, cudaStream_t st, const IsoLnpostParams &P, size_t smem, const IsoLnpostFlags &f)
{
    const int64_t want = (P.a.N + ISO_LNPOST_THREADS - 1) / ISO_LNPOST_THREADS;
    const int64_t cap = (int64_t)ctx->prop.multiProcessorCount * ISO_LNPOST_BLOCKS_PER_SM;
    int blocks = (int)(want < cap ? want : cap);
    if (blocks < 1) blocks = 1;
#define ISO_LAUNCH7(CAT, PROF, TRK, PEER, SEQ, CUBE)                                                                     \
    do {                                                                                                                 \
        if (smem > 48 * 1024)                                                                                            \
            ISO_CUDA(ctx, cudaFuncSetAttribute(iso_lnpost_kernel<NS, CAT, PROF, TRK, PEER, SEQ, CUBE>,                   \
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                 \
        iso_lnpost_kernel<NS, CAT, PROF, TRK, PEER, SEQ, CUBE><<<blocks, ISO_LNPOST_THREADS, smem, st>>>(P);             \
    } while (0)
    // the unit-cube kernels exist for single-model launches without the fused gather (what MultiNest-style callers and
    // the prior draws need); the star-sequential form for multi-star models whose BC pack is one 4-band chunk
#define ISO_LAUNCH5(CAT, PROF, TRK, PEER)                                                                                \
    do {                                                                                                                 \
        if (f.cube && !CAT && !PEER) {                                                                                   \
            if (NS > 1 && f.seq == 1) ISO_LAUNCH7(false, PROF, TRK, false, (NS > 1 ? 1 : 0), true);                      \
            else if (NS > 1 && f.seq == 2) ISO_LAUNCH7(false, PROF, TRK, false, (NS > 1 ? 2 : 0), true);                 \
            else if (NS > 1 && f.seq == 3) ISO_LAUNCH7(false, PROF, TRK, false, (NS > 1 ? 3 : 0), true);                 \
            else ISO_LAUNCH7(false, PROF, TRK, false, 0, true);                                                          \
        } else if (NS > 1 && f.seq == 1) ISO_LAUNCH7(CAT, PROF, TRK, PEER, (NS > 1 ? 1 : 0), false);                     \
        else if (NS > 1 && f.seq == 2) ISO_LAUNCH7(CAT, PROF, TRK, PEER, (NS > 1 ? 2 : 0), false);                       \
        else if (NS > 1 && f.seq == 3) ISO_LAUNCH7(CAT, PROF, TRK, PEER, (NS > 1 ? 3 : 0), false);                       \
        else ISO_LAUNCH7(CAT, PROF, TRK, PEER, 0, false);                                                                \
    } while (0)
#define ISO_LAUNCH3(CAT, PROF, TRK)                                                                                      \
    do {                                                                                                                 \
        if (f.peer) ISO_LAUNCH5(CAT, PROF, TRK, true);                                                                   \
        else ISO_LAUNCH5(CAT, PROF, TRK, false);                                                                         \
    } while (0)
#define ISO_LAUNCH1(TRK)                                                               \
    do {                                                                               \
        if (f.catalog) {                                                               \
            if (f.profile_default) ISO_LAUNCH3(true, ISO_PROFILE_DEFAULT, TRK);        \
            else ISO_LAUNCH3(true, ISO_PROFILE_GENERIC, TRK);                          \
        } else {                                                                       \
            if (f.profile_default) ISO_LAUNCH3(false, ISO_PROFILE_DEFAULT, TRK);       \
            else ISO_LAUNCH3(false, ISO_PROFILE_GENERIC, TRK);                         \
        }                                                                              \
    } while (0)
    if (f.cube && (f.catalog || f.peer))
        return iso_set_error(ctx, ISO_E_UNSUPPORTED, "unit-cube rows: single-model launches without the fused gather only");
    if (NS == 1 && f.track) ISO_LAUNCH1((NS == 1));   // track grids hold single stars only (starmodel.py:1396-1397)
    else ISO_LAUNCH1(false);
#undef ISO_LAUNCH1
#undef ISO_LAUNCH3
#undef ISO_LAUNCH5
#undef ISO_LAUNCH7
    ISO_CUDA(ctx, cudaGetLastError());
    return ISO_OK;
}
#endif
