// isochrones_b200 — evaluation of ONE row of the lnpost path (device function shared by the batch kernel in
// iso_lnpost.cu and the on-device ensemble sampler in iso_sampler.cu).
//
// Per row it replaces the chain
//   StarModel.lnpost            starmodel.py:538-542
//   BasicStarModel.lnprior      starmodel.py:1616-1635   (+ the prior classes, iso_prior.cuh; EEP_prior priors.py:409-429)
//   BasicStarModel.lnlike       starmodel.py:1563-1614
//   star_lnlike                 likelihood.py:16-147     (gauss_lnprob :10-13, fast_addmags utils.py:67-75)
//   interp_mag                  mags.py:8-61             (interp_value_3d / _4d interp.py:252-338)
#pragma once

#include "iso_common.cuh"
#include "iso_prior.cuh"

#ifndef ISO_MODEL_LAYOUT
#define ISO_MODEL_LAYOUT ISO_LAYOUT_NODE48   // default layout of the model cell gather (iso_common.cuh)
#endif

// device image of one star model (built from the public iso_model by iso_models_stage)
struct IsoGaussDev {
    double val;   // observed value
    double c;     // log(1 / sqrt(2 pi)) + log(unc)      (likelihood.py:13 — note the PLUS sign)
    double h;     // 1 / (unc * unc)
};

struct IsoModelDev {
    int n_stars, eep_replaces_age;
    int index_order[5];
    int obs_mask;          // bit c: BC-pack column c is an observed band
    int spec_mask;         // bit i: Teff / logg / feh observed (value not NaN, likelihood.py:127)
    int has_plax, has_nu_max, has_delta_nu;
    int eep_has_bounds;
    int pad_;
    IsoGaussDev spec[3];
    IsoGaussDev mag[ISO_MAX_BANDS];   // indexed by BC-pack column
    IsoGaussDev plax, nu_max, delta_nu;
    double eep_lo, eep_hi, eep_norm, eep_inv_norm;
    // default profile on isochrone grids (orig_prior = Chabrier-like BrokenPrior[LogNormal, PowerLaw] of the mass): log of
    // every x-independent factor of orig_prior(mass) / eep_norm per component, so that ln(pdf dm/dEEP) is a sum of logs
    double eep_lnc[2];
    iso_prior eep_orig, mass, age, feh, distance, AV;
    int profile_default;   // 1: the priors match ISO_PROFILE_DEFAULT
    int pad2_;
};

struct iso_models {
    IsoModelDev *d_models = nullptr;
    int n_models = 0;
    int n_stars = 0;
    int device = 0;
    int max_col = -1;      // highest BC-pack column any model observes
    IsoModelDev h_first;   // host copy of model 0 (passed by value in the kernel parameter block)
    bool needs_seismo = false;
    bool profile_default = false;   // every model matches ISO_PROFILE_DEFAULT
    bool track = false;             // evolution-track grid (all models share the grid kind)
};

// PROFILE selects how the priors are evaluated:
//   ISO_PROFILE_DEFAULT — the prior classes of a default BasicStarModel (starmodel.py:1441-1448): Chabrier mass
//       (BrokenPrior[LogNormal, PowerLaw]), FehPrior, FlatLog age, PowerLaw distance, Flat AV, with any bounds /
//       constants; the kinds are compile-time, so the code is small, switch-free and shares log(distance);
//   ISO_PROFILE_GENERIC — any supported prior object per parameter (set_prior), through out-of-line dispatchers.
// TRACK: evolution-track grid (mass, eep, feh, d, AV) vs isochrone grid (eep_0.., age, feh, d, AV).
enum { ISO_PROFILE_GENERIC = 0, ISO_PROFILE_DEFAULT = 1 };

// the grids as a kernel sees them: descriptors (kernel parameter block) + axis tables staged in shared memory
struct IsoRowGrids {
    IsoGridDev mg, bg;
    int smem_axis_off[2][ISO_MAX_DIM];   // offset (in double2) of each axis table in shared memory; -1: closed form
    int smem_nodes;                      // double2 entries of axis tables in shared memory
};

struct IsoRowResult {
    double lnpost, lnprior, lnlike;
};

// host helpers (iso_lnpost.cu)
int iso_row_grids_fill(iso_ctx *ctx, const iso_grid *mp, const iso_grid *bp, IsoRowGrids *out, size_t *smem_bytes);
int iso_check_lnpost_handles(iso_ctx *ctx, const iso_grid *mp, const iso_grid *bp, const iso_models *models);

#ifdef __CUDACC__

__device__ __forceinline__ double iso_gauss(const IsoGaussDev &g, double model_val)
{
    double resid = g.val - model_val;
    return g.c - 0.5 * resid * resid * g.h;
}

__device__ __forceinline__ double iso_sel5(const double (&a)[5], int i)
{
    return i == 0 ? a[0] : i == 1 ? a[1] : i == 2 ? a[2] : i == 3 ? a[3] : a[4];
}

template <int NDIM>
__device__ __forceinline__ bool iso_locate_smem(const IsoGridDev &g, const double2 *smem, const int (&soff)[ISO_MAX_DIM],
                                                const double (&x)[NDIM], int (&idx)[NDIM], double (&y)[NDIM])
{
    bool ok = true;
#pragma unroll
    for (int d = 0; d < NDIM; d++) ok = ok && iso_in_bounds(g.ax[d], x[d]);
    if (!ok) return false;
#pragma unroll
    for (int d = 0; d < NDIM; d++) idx[d] = iso_axis_locate(g.ax[d], smem + soff[d], x[d], y[d]);
    return true;
}

// ---- TMA bulk copy (cp.async.bulk, 1-D) + mbarrier: SASS UBLKCP / SYNCS ------------------------------------------
__device__ __forceinline__ unsigned iso_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void iso_mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(iso_smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void iso_mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(iso_smem_u32(bar)), "r"(bytes) : "memory");
}

// global -> shared bulk copy; bytes a multiple of 16, both addresses 16-byte aligned; completion is signalled on `bar`
__device__ __forceinline__ void iso_bulk_g2s(void *dst_smem, const void *src_gmem, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     iso_smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(iso_smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void iso_mbar_wait(unsigned long long *bar, unsigned parity)
{
    asm volatile(
        "{\n"
        " .reg .pred p;\n"
        " ISO_WAIT_%=:\n"
        " mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        " @p bra ISO_DONE_%=;\n"
        " bra ISO_WAIT_%=;\n"
        " ISO_DONE_%=:\n"
        "}" ::"r"(iso_smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// Stage the tables of the non-closed-form axes in shared memory: one elected thread arms an mbarrier with the total
// byte count and issues one TMA bulk copy per axis table (contiguous double2 runs, 16-byte aligned on both sides);
// every thread then waits on the barrier's phase.  The loops are fully unrolled: the parameter block must only be
// indexed with compile-time constants or it is copied to local memory.  All threads of the CTA must call this.
__device__ __forceinline__ void iso_stage_axis_tables(const IsoRowGrids &G, double2 *s_nodes)
{
    __shared__ __align__(8) unsigned long long tables_bar;
    if (threadIdx.x == 0) iso_mbar_init(&tables_bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned total = 0;
#pragma unroll
        for (int d = 0; d < 3; d++)
            if (G.smem_axis_off[0][d] >= 0) total += (unsigned)G.mg.ax[d].n * (unsigned)sizeof(double2);
#pragma unroll
        for (int d = 0; d < 4; d++)
            if (G.smem_axis_off[1][d] >= 0) total += (unsigned)G.bg.ax[d].n * (unsigned)sizeof(double2);
        iso_mbar_expect_tx(&tables_bar, total);
#pragma unroll
        for (int d = 0; d < 3; d++) {
            const int so = G.smem_axis_off[0][d];
            if (so >= 0)
                iso_bulk_g2s(s_nodes + so, G.mg.nodes + G.mg.ax[d].off, (unsigned)G.mg.ax[d].n * (unsigned)sizeof(double2),
                             &tables_bar);
        }
#pragma unroll
        for (int d = 0; d < 4; d++) {
            const int so = G.smem_axis_off[1][d];
            if (so >= 0)
                iso_bulk_g2s(s_nodes + so, G.bg.nodes + G.bg.ax[d].off, (unsigned)G.bg.ax[d].n * (unsigned)sizeof(double2),
                             &tables_bar);
        }
    }
    iso_mbar_wait(&tables_bar, 0);
}

// Bolometric corrections of one chunk of four packed bands at a located BC cell: 16 corners x one 32-byte sector
// (interp_value_4d, interp.py:296-338, for the chunk's columns; accumulation in the reference's corner order).
__device__ __forceinline__ void iso_bc_chunk(const IsoGridDev &bg, const int (&idx4)[4], const double (&y4)[4], int ch,
                                             double (&bcv)[4])
{
    unsigned node[16];
    double w[16];
    iso_corners<4>(bg, idx4, y4, node, w);
    double b0 = 0.0, b1 = 0.0, b2 = 0.0, b3 = 0.0;
#pragma unroll
    for (int j = 0; j < 16; j++) {
        iso_d4 q = iso_ldg256(bg.g + (size_t)node[j] * bg.ncols + 4 * ch);
        b0 = fma(q.x, w[j], b0);
        b1 = fma(q.y, w[j], b1);
        b2 = fma(q.z, w[j], b2);
        b3 = fma(q.w, w[j], b3);
    }
    bcv[0] = b0;
    bcv[1] = b1;
    bcv[2] = b2;
    bcv[3] = b3;
}

// One row.  want_prior / want_like: the caller asked for the separate lnprior / lnlike values (every term is then
// evaluated, as BasicStarModel.lnprior / lnlike would); otherwise rows whose prior is already known to be
// non-finite return lnpost = -inf early, which is exactly what StarModel.lnpost returns (starmodel.py:540-541).
// SEQ > 0 (multiple stars, BC pack of exactly SEQ 4-band chunks): every star's magnitudes are evaluated right after
// its model cell, so only flux[4 * SEQ] survives from star to star instead of each star's located BC cell (4 indices +
// 4 distances + Mbol): the binary / triple kernels then fit in 128 registers (they spilled 250-350 bytes before).
// The arithmetic and its order are those of the chunk-major form (SEQ = 0), the results bit-identical.
template <int NSTARS, int PROFILE, bool TRACK, int LAYOUT = ISO_MODEL_LAYOUT, int SEQ = 0>
__device__ __forceinline__ IsoRowResult iso_lnpost_row(const IsoRowGrids &G, const double2 *s_nodes, const IsoModelDev &m,
                                                       const double (&p)[NSTARS + 4], bool want_prior, bool want_like)
{
    constexpr bool DEF = PROFILE == ISO_PROFILE_DEFAULT;
    const IsoGridDev &mg = G.mg;
    const IsoGridDev &bg = G.bg;
    const double nan = iso_nan();
    const double neg_inf = iso_neg_inf();
    const bool want_parts = want_prior || want_like;
    const int bc_chunks = bg.ncols >> 2;
    IsoRowResult out;
    out.lnpost = neg_inf;
    out.lnprior = nan;
    out.lnlike = nan;
    {
        const double other = p[NSTARS], feh_in = p[NSTARS + 1], dist = p[NSTARS + 2], AV = p[NSTARS + 3];

        // ---- priors that need no grid access; order of the final sum follows param_names ----------------
        bool order_bad = false;   // starmodel.py:1618-1623 (N = 3 precedence as written in the reference)
        if (NSTARS == 2) order_bad = p[1] > p[0];
        if (NSTARS == 3) order_bad = !(p[0] > p[1]) && (p[1] > p[2]);
        // track grids (mass, eep, feh, ..): p[0] is the mass (prior "mass") and `other` the EEP;
        // isochrone grids (eep_0.., age, feh, ..): p[k] are the EEPs and `other` the age (prior "age")
        double lnp_other, lnp_feh, lnp_dist, lnp_AV, lnd = 0.0;
        if (DEF) {
            // one log(mass) serves both Chabrier components: walkers around the break at 1 Msun diverge inside a
            // warp, and each side would otherwise evaluate its own logarithm
            lnp_other = TRACK ? iso_broken2_lnpdf<ISO_PRIOR_LOGNORMAL, ISO_PRIOR_POWERLAW, true>(m.mass, p[0], log(p[0]))
                              : iso_leaf_lnpdf<ISO_PRIOR_FLATLOG>(m.age.self, other);
            lnp_feh = iso_leaf_lnpdf<ISO_PRIOR_FEH>(m.feh.self, feh_in);
            lnd = log(dist);   // shared by the distance prior and the distance modulus
            lnp_dist = iso_leaf_lnpdf<ISO_PRIOR_POWERLAW, true>(m.distance.self, dist, lnd);
            lnp_AV = iso_leaf_lnpdf<ISO_PRIOR_FLAT>(m.AV.self, AV);
        } else {
            lnp_other = TRACK ? iso_prior_lnpdf_dyn(&m.mass, p[0]) : iso_prior_lnpdf_dyn(&m.age, other);
            lnp_feh = iso_prior_lnpdf_dyn(&m.feh, feh_in);
            lnp_dist = iso_prior_lnpdf_dyn(&m.distance, dist);
            lnp_AV = iso_prior_lnpdf_dyn(&m.AV, AV);
        }
        const double cheap = lnp_other + lnp_feh + lnp_dist + lnp_AV;
        if (!want_parts && (order_bad || !isfinite(cheap))) {
            // lnprior is already known to be -inf or NaN: StarModel.lnpost returns -inf (starmodel.py:540-541)
            return out;
        }

        // ---- model-grid gather per star --------------------------------------------------------------
        constexpr int NKEEP = SEQ ? 1 : NSTARS;   // located BC cells kept for the chunk-major likelihood
        double lnp_eep[NSTARS], Mbol[NKEEP], y4[NKEEP][4];
        int idx4[NKEEP][4];
        bool bc_ok[NKEEP];
        double Teff = nan, logg = nan, feh_s = nan, nu_max = nan, delta_nu = nan;
        // mags.py:52: 5 log10(d / 10); the default profile already holds log(d)
        const double dist_mod = DEF ? 5.0 * fma(lnd, 0.43429448190325182765, -1.0) : 5.0 * log10(dist / 10.0);
        constexpr int NFLUX = SEQ ? 4 * SEQ : 4;
        double flux[NFLUX];                      // SEQ: summed fluxes of the band chunks
#pragma unroll
        for (int b = 0; b < NFLUX; b++) flux[b] = 0.0;
        bool eeps_finite = true;                 // SEQ: every star so far has a finite EEP prior
#pragma unroll
        for (int k = 0; k < NSTARS; k++) {
            // star k uses [pars[k], shared parameters...]  (likelihood.py:43-54)
            // model-grid coordinates pars[index_order[0..2]] (models.py:669, 696): tracks (feh, mass, eep), isochrones
            // (age, feh, eep) — fixed by the grid kind, checked on the host
            const double x[3] = {TRACK ? feh_in : other, TRACK ? p[0] : feh_in, TRACK ? other : p[k]};
            double y[3];
            int idx[3];
            double v[8] = {nan, nan, nan, nan, nan, nan, nan, nan};
            if (iso_locate_smem<3>(mg, s_nodes, G.smem_axis_off[0], x, idx, y)) {
                unsigned node[8];
                double w[8];
                iso_corners<3>(mg, idx, y, node, w);
#pragma unroll
                for (int c = 0; c < 8; c++) v[c] = 0.0;
                if constexpr (LAYOUT == ISO_LAYOUT_PAIR96) {
                    // The two EEP-adjacent corners 2s, 2s + 1 of a cell are ONE 96-byte pair record (3 sectors instead of
                    // 4): record node[2s] holds the six always-needed columns of flat nodes node[2s] and node[2s] + 1 —
                    // the same node the reference's unchecked index arithmetic reaches for corner 2s + 1, incl. the zero
                    // padding past the array.  Accumulation stays in the reference's corner order.
#pragma unroll
                    for (int s2 = 0; s2 < 4; s2++) {
                        const double *rec = mg.gp + (size_t)node[2 * s2] * ISO_PP_STRIDE;
                        const iso_d4 q0 = iso_ldg256(rec), q1 = iso_ldg256(rec + 4), q2 = iso_ldg256(rec + 8);
                        const double wa = w[2 * s2], wb = w[2 * s2 + 1];
                        v[0] = fma(q0.x, wa, v[0]);
                        v[1] = fma(q0.y, wa, v[1]);
                        v[2] = fma(q0.z, wa, v[2]);
                        v[3] = fma(q0.w, wa, v[3]);
                        v[4] = fma(q1.x, wa, v[4]);
                        v[5] = fma(q1.y, wa, v[5]);
                        v[0] = fma(q1.z, wb, v[0]);
                        v[1] = fma(q1.w, wb, v[1]);
                        v[2] = fma(q2.x, wb, v[2]);
                        v[3] = fma(q2.y, wb, v[3]);
                        v[4] = fma(q2.z, wb, v[4]);
                        v[5] = fma(q2.w, wb, v[5]);
                    }
                } else if constexpr (LAYOUT == ISO_LAYOUT_NODE48) {
                    // 48-byte nodes, no duplication: corners 2s, 2s + 1 are the 96 contiguous bytes of flat nodes
                    // node[2s], node[2s] + 1 (same flat arithmetic as above).  The run starts 32-byte aligned for an even
                    // node (3 sectors) and 16 bytes into a sector for an odd one (4 sectors: the fourth load is predicated);
                    // the 12 values are picked out of the aligned sectors with selects.
#pragma unroll
                    for (int s2 = 0; s2 < 4; s2++) {
                        const unsigned n0 = node[2 * s2];
                        const bool odd = (n0 & 1u) != 0;
                        const double *al = mg.g48 + ((size_t)n0 * ISO_PP_NCOLS - (odd ? 2 : 0));
                        const iso_d4 q0 = iso_ldg256(al), q1 = iso_ldg256(al + 4), q2 = iso_ldg256(al + 8);
                        iso_d4 q3;
                        q3.x = q3.y = q3.z = q3.w = 0.0;
                        if (odd) q3 = iso_ldg256(al + 12);
                        const double wa = w[2 * s2], wb = w[2 * s2 + 1];
                        v[0] = fma(odd ? q0.z : q0.x, wa, v[0]);
                        v[1] = fma(odd ? q0.w : q0.y, wa, v[1]);
                        v[2] = fma(odd ? q1.x : q0.z, wa, v[2]);
                        v[3] = fma(odd ? q1.y : q0.w, wa, v[3]);
                        v[4] = fma(odd ? q1.z : q1.x, wa, v[4]);
                        v[5] = fma(odd ? q1.w : q1.y, wa, v[5]);
                        v[0] = fma(odd ? q2.x : q1.z, wb, v[0]);
                        v[1] = fma(odd ? q2.y : q1.w, wb, v[1]);
                        v[2] = fma(odd ? q2.z : q2.x, wb, v[2]);
                        v[3] = fma(odd ? q2.w : q2.y, wb, v[3]);
                        v[4] = fma(odd ? q3.x : q2.z, wb, v[4]);
                        v[5] = fma(odd ? q3.y : q2.w, wb, v[5]);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const double *base = mg.g + (size_t)node[j] * ISO_MP_NCOLS;
                        iso_d4 lo = iso_ldg256(base), hi = iso_ldg256(base + 4);
                        v[0] = fma(lo.x, w[j], v[0]);
                        v[1] = fma(lo.y, w[j], v[1]);
                        v[2] = fma(lo.z, w[j], v[2]);
                        v[3] = fma(lo.w, w[j], v[3]);
                        v[4] = fma(hi.x, w[j], v[4]);
                        v[5] = fma(hi.y, w[j], v[5]);
                        v[6] = fma(hi.z, w[j], v[6]);
                        v[7] = fma(hi.w, w[j], v[7]);
                    }
                }
                if (LAYOUT != ISO_LAYOUT_NODE64 && k == 0 && m.has_nu_max) {   // asteroseismic columns (rare): from the 8-column nodes
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const double2 sv = __ldg(reinterpret_cast<const double2 *>(mg.g + (size_t)node[j] * ISO_MP_NCOLS + ISO_MP_NU_MAX));
                        v[ISO_MP_NU_MAX] = fma(sv.x, w[j], v[ISO_MP_NU_MAX]);
                        v[ISO_MP_DELTA_NU] = fma(sv.y, w[j], v[ISO_MP_DELTA_NU]);
                    }
                }
            }
            // EEP_prior.lnpdf: BoundedPrior.lnpdf :131-140 -> Prior.pdf :54-59 -> EEP_prior._pdf :423-429
            const double eep = TRACK ? other : p[k];
            if (m.eep_has_bounds && iso_outside(eep, m.eep_lo, m.eep_hi)) {
                lnp_eep[k] = neg_inf;
            } else if (DEF && TRACK) {
                // orig_prior = FlatLogPrior on log10 age: pdf = k2 10^age inside its bounds, so
                // log(pdf * deriv / norm) = age ln10 + log(k2 deriv / norm); 0 -> -inf, negative / NaN -> NaN as in
                // `np.log(pdf) if pdf else -np.inf`
                const iso_prior_leaf &op = m.eep_orig.self;
                const double age = v[ISO_MP_ORIG];
                if ((op.flags & ISO_PF_HAS_BOUNDS) && iso_outside(age, op.lo, op.hi))
                    lnp_eep[k] = neg_inf;
                else
                    lnp_eep[k] = fma(age, op.k[0], iso_log_or_neginf(op.k[2] * v[ISO_MP_DERIV] * m.eep_inv_norm));
            } else if (DEF) {
                // orig_prior = Chabrier-like broken prior of the mass, called as Prior.__call__ (bounds test, then
                // components[i](x) / norms[i] / _norm, priors.py:35-36, 54-59, 205-207).  The density is positive wherever it
                // is not cut to zero, so log(pdf * deriv / norm) = ln pdf + log(deriv): one log(mass) serves both components
                // (walkers around the 1 Msun break diverge inside a warp) and no exp / pow is evaluated.  0 -> -inf,
                // negative / NaN deriv -> NaN, as in `np.log(pdf) if pdf else -np.inf`.
                const iso_prior &op = m.eep_orig;
                const double mass = v[ISO_MP_ORIG];
                const bool upper = mass != mass || op.breakpoints[0] <= mass;   // np.digitize: NaN -> last component
                const iso_prior_leaf &c = upper ? op.comp[1] : op.comp[0];
                if (((op.self.flags & ISO_PF_HAS_BOUNDS) && iso_outside(mass, op.self.lo, op.self.hi)) ||
                    ((c.flags & ISO_PF_HAS_BOUNDS) && iso_outside(mass, c.lo, c.hi))) {
                    lnp_eep[k] = neg_inf;
                } else {
                    const double lnm = log(mass);
                    const double ly = lnm - op.comp[0].a[0], t = ly * op.comp[0].k[3];      // LogNormal: y = mass / scale
                    const double ln_pdf = upper ? fma(op.comp[1].a[0], lnm, m.eep_lnc[1]) : m.eep_lnc[0] - ly - 0.5 * (t * t);
                    lnp_eep[k] = ln_pdf + iso_log_or_neginf(v[ISO_MP_DERIV]);
                }
            } else {
                double pdf = iso_prior_call_dyn(&m.eep_orig, v[ISO_MP_ORIG]);
                lnp_eep[k] = iso_log_or_neginf(pdf * v[ISO_MP_DERIV] * m.eep_inv_norm);
            }
            if (!SEQ) Mbol[k] = v[ISO_MP_MBOL];
            if (k == 0) {   // companions' Teff / logg / feh are discarded (likelihood.py:76, 96)
                Teff = v[ISO_MP_TEFF];
                logg = v[ISO_MP_LOGG];
                feh_s = v[ISO_MP_FEH];
                nu_max = v[ISO_MP_NU_MAX];
                delta_nu = v[ISO_MP_DELTA_NU];
            }
            // BC cell of this star: interp_value_4d(Teff, logg, feh, AV)  mags.py:49-50
            const double x4[4] = {v[ISO_MP_TEFF], v[ISO_MP_LOGG], v[ISO_MP_FEH], AV};
            if (!SEQ) {
                bc_ok[k] = iso_locate_smem<4>(bg, s_nodes, G.smem_axis_off[1], x4, idx4[k], y4[k]);
            } else {
                // this star's fluxes now (fast_addmags utils.py:67-75 sums them in star order); skipped when the prior
                // is already known to be non-finite and no separate lnlike was asked for (StarModel.lnpost never
                // evaluates the likelihood then)
                eeps_finite = eeps_finite && isfinite(lnp_eep[k]);
                if constexpr (SEQ == 1) {
                    const int cm = m.obs_mask & 0xF;
                    if (cm && (want_like || (eeps_finite && !order_bad && isfinite(cheap)))) {
                        double mg4[4] = {nan, nan, nan, nan};
                        if (iso_locate_smem<4>(bg, s_nodes, G.smem_axis_off[1], x4, idx4[0], y4[0])) {
                            double bcv[4];
                            iso_bc_chunk(bg, idx4[0], y4[0], 0, bcv);
                            const double mb = v[ISO_MP_MBOL] + dist_mod;   // mags.py:59: Mbol + dist_mod - bc
#pragma unroll
                            for (int b = 0; b < 4; b++) mg4[b] = mb - bcv[b];
                        }
#pragma unroll
                        for (int b = 0; b < 4; b++)
                            if (cm & (1 << b)) flux[b] += exp10(-0.4 * mg4[b]);
                    }
                } else if (m.obs_mask && (want_like || (eeps_finite && !order_bad && isfinite(cheap)))) {
                    // several chunks: the located BC cell serves them one after the other
                    const bool located = iso_locate_smem<4>(bg, s_nodes, G.smem_axis_off[1], x4, idx4[0], y4[0]);
                    const double mb = v[ISO_MP_MBOL] + dist_mod;
#pragma unroll
                    for (int ch = 0; ch < SEQ; ch++) {
                        const int cm = (m.obs_mask >> (4 * ch)) & 0xF;
                        if (!cm) continue;
                        double mg4[4] = {nan, nan, nan, nan};
                        if (located) {
                            double bcv[4];
                            iso_bc_chunk(bg, idx4[0], y4[0], ch, bcv);
#pragma unroll
                            for (int b = 0; b < 4; b++) mg4[b] = mb - bcv[b];
                        }
#pragma unroll
                        for (int b = 0; b < 4; b++)
                            if (cm & (1 << b)) flux[4 * ch + b] += exp10(-0.4 * mg4[b]);
                    }
                }
            }
        }

        // ---- lnprior: sum in param_names order (models.py:665, 692; starmodel.py:1510-1518) ----------------
        double lnprior;
        if (order_bad) {
            lnprior = neg_inf;
        } else {
            lnprior = 0.0;
            if (TRACK) {   // (mass, eep, feh, distance, AV)
                lnprior += lnp_other;
                lnprior += lnp_eep[0];
            } else {                    // (eep_0 .. eep_{N-1}, age, feh, distance, AV)
#pragma unroll
                for (int k = 0; k < NSTARS; k++) lnprior += lnp_eep[k];
                lnprior += lnp_other;
            }
            lnprior += lnp_feh;
            lnprior += lnp_dist;
            lnprior += lnp_AV;
        }
        out.lnprior = lnprior;
        const bool prior_ok = isfinite(lnprior);
        if (!prior_ok && !want_like) {
            return out;
        }

        // ---- lnlike -----------------------------------------------------------------------------------
        double ll = 0.0;
        if (m.spec_mask & 1) ll += iso_gauss(m.spec[0], Teff);
        if (m.spec_mask & 2) ll += iso_gauss(m.spec[1], logg);
        if (m.spec_mask & 4) ll += iso_gauss(m.spec[2], feh_s);
        if (SEQ) {
            const int om = m.obs_mask & ((1 << NFLUX) - 1);
#pragma unroll
            for (int b = 0; b < NFLUX; b++)
                if (om & (1 << b)) ll += iso_gauss(m.mag[b], -2.5 * log10(flux[b]));
        } else if (m.obs_mask) {
            for (int ch = 0; ch < bc_chunks; ch++) {
                const int cm = (m.obs_mask >> (4 * ch)) & 0xF;
                if (!cm) continue;
                double tot[4];
                double fl[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
                for (int k = 0; k < NKEEP; k++) {
                    double mg4[4] = {nan, nan, nan, nan};
                    if (bc_ok[k]) {
                        double bcv[4];
                        iso_bc_chunk(bg, idx4[k], y4[k], ch, bcv);
                        const double mb = Mbol[k] + dist_mod;   // mags.py:59: Mbol + dist_mod - bc
#pragma unroll
                        for (int b = 0; b < 4; b++) mg4[b] = mb - bcv[b];
                    }
                    if (NSTARS == 1) {
#pragma unroll
                        for (int b = 0; b < 4; b++) tot[b] = mg4[b];
                    } else {   // fast_addmags utils.py:67-75 — only evaluated for observed columns
#pragma unroll
                        for (int b = 0; b < 4; b++)
                            if (cm & (1 << b)) fl[b] += exp10(-0.4 * mg4[b]);
                    }
                }
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    if (!(cm & (1 << b))) continue;
                    if (NSTARS > 1) tot[b] = -2.5 * log10(fl[b]);
                    ll += iso_gauss(m.mag[4 * ch + b], tot[b]);
                }
            }
        }
        if (m.has_plax) ll += iso_gauss(m.plax, 1000.0 / dist);   // starmodel.py:1599-1601
        if (m.has_nu_max) {                                        // starmodel.py:1604-1612
            ll += iso_gauss(m.nu_max, nu_max);
            if (m.has_delta_nu) ll += iso_gauss(m.delta_nu, delta_nu);
        }
        out.lnlike = ll;
        out.lnpost = prior_ok ? lnprior + ll : neg_inf;   // starmodel.py:538-542
    }
    return out;
}

#endif  // __CUDACC__
