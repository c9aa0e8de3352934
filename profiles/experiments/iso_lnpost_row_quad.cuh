// isochrones_b200 — ONE row of the lnpost path evaluated by FOUR lanes (a "quad"), for the latency-bound case: a single
// ensemble of a few hundred walkers, where a half-step is one row per thread and the run time is the dependent
// instruction chain of that row times the number of half-steps (iso_sampler.cu).  Splitting the row over a quad cuts the
// chain: every lane carries a quarter of the searches, priors, corner loads and likelihood terms, and the SM holds four
// times as many warps to overlap.
//
// The result is BIT-IDENTICAL to iso_lnpost_row<1, ..>: every sum is formed in the same order with the same operations —
//   * the multilinear sums are a chain of FMAs in the reference's corner order (interp.py:286-291): lane s of the quad
//     holds corners 2s, 2s+1 (model cell) or 4s .. 4s+3 (BC cell), and the running sums are handed from lane to lane with
//     shuffles, so the chain is the thread-per-row one cut into four pieces;
//   * prior / likelihood TERMS are computed on different lanes, gathered with shuffles and added on every lane in the
//     order of iso_lnpost_row.
// Control flow is warp-uniform (full-mask shuffles): nothing returns early; rows whose prior is not finite are carried
// through with benign indices and come out as -inf exactly as StarModel.lnpost returns them (starmodel.py:538-542).
// Single star, no asteroseismic terms (the host picks the thread-per-row kernel otherwise).
#pragma once

#include "iso_lnpost_row.cuh"

#ifdef __CUDACC__

__device__ __forceinline__ double iso_quad_bcast(double v, int src_lane) { return __shfl_sync(0xffffffffu, v, src_lane); }
__device__ __forceinline__ int iso_quad_bcast(int v, int src_lane) { return __shfl_sync(0xffffffffu, v, src_lane); }

// p: the row (identical on the four lanes).  Returns lnpost on every lane of the quad.
template <int PROFILE, bool TRACK>
__device__ __forceinline__ double iso_lnpost_row_quad(const IsoRowGrids &G, const double2 *s_nodes, const IsoModelDev &m,
                                                      const double (&p)[5])
{
    constexpr bool DEF = PROFILE == ISO_PROFILE_DEFAULT;
    const IsoGridDev &mg = G.mg;
    const IsoGridDev &bg = G.bg;
    const double nan = iso_nan();
    const double neg_inf = iso_neg_inf();
    const int lane = threadIdx.x & 31;
    const int ql = lane & 3;          // lane within the quad
    const int q0 = lane & ~3;         // first lane of the quad
    const double other = p[1], feh_in = p[2], dist = p[3], AV = p[4];

    // ---- stage 1: grid-free priors and the model-grid searches, one piece per lane ----------------------------------
    //   lane 0: prior of the parameter EEP does not replace (mass | age)        lane 2: distance prior, log(d), axis 0 search
    //   lane 1: [Fe/H] prior (the longest chain: three exp and a log)            lane 3: AV prior, axes 1 and 2 searches
    // model-grid coordinates: tracks (feh, mass, eep), isochrones (age, feh, eep)   (models.py:669, 696)
    const double x0 = TRACK ? feh_in : other, x1 = TRACK ? p[0] : feh_in, x2 = TRACK ? other : p[0];
    double term = 0.0, lnd = 0.0, ya = 0.0, yb = 0.0;
    int ia = 0, ib = 0;
    bool inb = true;
    if (ql == 0) {
        if (DEF)
            term = TRACK ? iso_broken2_lnpdf<ISO_PRIOR_LOGNORMAL, ISO_PRIOR_POWERLAW, true>(m.mass, p[0], log(p[0]))
                         : iso_leaf_lnpdf<ISO_PRIOR_FLATLOG>(m.age.self, other);
        else
            term = TRACK ? iso_prior_lnpdf_dyn(&m.mass, p[0]) : iso_prior_lnpdf_dyn(&m.age, other);
    } else if (ql == 1) {
        term = DEF ? iso_leaf_lnpdf<ISO_PRIOR_FEH>(m.feh.self, feh_in) : iso_prior_lnpdf_dyn(&m.feh, feh_in);
    } else if (ql == 2) {
        if (DEF) {
            lnd = log(dist);
            term = iso_leaf_lnpdf<ISO_PRIOR_POWERLAW, true>(m.distance.self, dist, lnd);
        } else {
            term = iso_prior_lnpdf_dyn(&m.distance, dist);
        }
        inb = iso_in_bounds(mg.ax[0], x0);
        if (inb) ia = iso_axis_locate(mg.ax[0], s_nodes + G.smem_axis_off[0][0], x0, ya);
    } else {
        term = DEF ? iso_leaf_lnpdf<ISO_PRIOR_FLAT>(m.AV.self, AV) : iso_prior_lnpdf_dyn(&m.AV, AV);
        inb = iso_in_bounds(mg.ax[1], x1) && iso_in_bounds(mg.ax[2], x2);
        if (inb) {
            ia = iso_axis_locate(mg.ax[1], s_nodes + G.smem_axis_off[0][1], x1, ya);
            ib = iso_axis_locate(mg.ax[2], s_nodes + G.smem_axis_off[0][2], x2, yb);
        }
    }
    const double lnp_other = iso_quad_bcast(term, q0), lnp_feh = iso_quad_bcast(term, q0 + 1);
    const double lnp_dist = iso_quad_bcast(term, q0 + 2), lnp_AV = iso_quad_bcast(term, q0 + 3);
    lnd = iso_quad_bcast(lnd, q0 + 2);
    // (both shuffles are evaluated on every lane: `&&` would skip the second one on lanes where the first is false, and
    // a full-mask shuffle that part of the warp never executes deadlocks the rest)
    const int ok2 = iso_quad_bcast((int)inb, q0 + 2), ok3 = iso_quad_bcast((int)inb, q0 + 3);
    const bool model_ok = (ok2 & ok3) != 0;
    int idx[3];
    double y[3];
    idx[0] = iso_quad_bcast(ia, q0 + 2);
    y[0] = iso_quad_bcast(ya, q0 + 2);
    idx[1] = iso_quad_bcast(ia, q0 + 3);
    y[1] = iso_quad_bcast(ya, q0 + 3);
    idx[2] = iso_quad_bcast(ib, q0 + 3);
    y[2] = iso_quad_bcast(yb, q0 + 3);
    if (!model_ok) {   // benign cell; the values are replaced by NaN below, as interp_value_3d returns them
        idx[0] = idx[1] = idx[2] = 0;
        y[0] = y[1] = y[2] = 0.0;
    }
    // ---- stage 2: the model cell — lane s loads corners 2s, 2s + 1 (one EEP-adjacent pair of 48-byte nodes) -----------
    double v[6];
    {
        const int b0 = ql >> 1, b1 = ql & 1;
        unsigned nd = ((unsigned)(idx[0] + b0) * (unsigned)mg.n[1] + (unsigned)(idx[1] + b1)) * (unsigned)mg.n[2] + (unsigned)idx[2];
        nd = min(nd, (unsigned)mg.n_nodes);
        // weights in iso_corners' order: ((1 * f0) * f1) * f2
        const double f0 = b0 ? y[0] : (1.0 - y[0]), f1 = b1 ? y[1] : (1.0 - y[1]);
        const double w01 = (1.0 * f0) * f1;
        const double wa = w01 * (1.0 - y[2]), wb = w01 * y[2];
        const bool odd = (nd & 1u) != 0;
        const double *al = mg.g48 + ((size_t)nd * ISO_PP_NCOLS - (odd ? 2 : 0));
        const iso_d4 c0 = iso_ldg256(al), c1 = iso_ldg256(al + 4), c2 = iso_ldg256(al + 8);
        iso_d4 c3;
        c3.x = c3.y = c3.z = c3.w = 0.0;
        if (odd) c3 = iso_ldg256(al + 12);
        const double A[6] = {odd ? c0.z : c0.x, odd ? c0.w : c0.y, odd ? c1.x : c0.z, odd ? c1.y : c0.w, odd ? c1.z : c1.x, odd ? c1.w : c1.y};
        const double B[6] = {odd ? c2.x : c1.z, odd ? c2.y : c1.w, odd ? c2.z : c2.x, odd ? c2.w : c2.y, odd ? c3.x : c2.z, odd ? c3.y : c2.w};
#pragma unroll
        for (int c = 0; c < 6; c++) v[c] = 0.0;
#pragma unroll
        for (int s = 0; s < 4; s++) {
#pragma unroll
            for (int c = 0; c < 6; c++) {
                const double t = fma(B[c], wb, fma(A[c], wa, v[c]));
                v[c] = iso_quad_bcast(ql == s ? t : v[c], q0 + s);
            }
        }
        if (!model_ok) {
#pragma unroll
            for (int c = 0; c < 6; c++) v[c] = nan;
        }
    }
    // ---- stage 3: EEP prior (lane 0) while every lane searches one axis of the BC grid --------------------------------
    const double eep = TRACK ? other : p[0];
    double lnp_eep = 0.0;
    if (ql == 0) {
        if (m.eep_has_bounds && iso_outside(eep, m.eep_lo, m.eep_hi)) {
            lnp_eep = neg_inf;
        } else if (DEF && TRACK) {
            const iso_prior_leaf &op = m.eep_orig.self;
            const double age = v[ISO_MP_ORIG];
            if ((op.flags & ISO_PF_HAS_BOUNDS) && iso_outside(age, op.lo, op.hi))
                lnp_eep = neg_inf;
            else
                lnp_eep = fma(age, op.k[0], iso_log_or_neginf(op.k[2] * v[ISO_MP_DERIV] * m.eep_inv_norm));
        } else if (DEF) {
            const iso_prior &op = m.eep_orig;
            const double mass = v[ISO_MP_ORIG];
            const bool upper = mass != mass || op.breakpoints[0] <= mass;
            const iso_prior_leaf &c = upper ? op.comp[1] : op.comp[0];
            if (((op.self.flags & ISO_PF_HAS_BOUNDS) && iso_outside(mass, op.self.lo, op.self.hi)) ||
                ((c.flags & ISO_PF_HAS_BOUNDS) && iso_outside(mass, c.lo, c.hi))) {
                lnp_eep = neg_inf;
            } else {
                const double lnm = log(mass);
                const double ly = lnm - op.comp[0].a[0], t = ly * op.comp[0].k[3];
                const double ln_pdf = upper ? fma(op.comp[1].a[0], lnm, m.eep_lnc[1]) : m.eep_lnc[0] - ly - 0.5 * (t * t);
                lnp_eep = ln_pdf + iso_log_or_neginf(v[ISO_MP_DERIV]);
            }
        } else {
            const double pdf = iso_prior_call_dyn(&m.eep_orig, v[ISO_MP_ORIG]);
            lnp_eep = iso_log_or_neginf(pdf * v[ISO_MP_DERIV] * m.eep_inv_norm);
        }
    }
    // BC cell: interp_value_4d(Teff, logg, feh, AV)  mags.py:49-50 — lane q locates axis q
    const double xq = ql == 0 ? v[ISO_MP_TEFF] : ql == 1 ? v[ISO_MP_LOGG] : ql == 2 ? v[ISO_MP_FEH] : AV;
    const IsoAxisDev &axq = ql == 0 ? bg.ax[0] : ql == 1 ? bg.ax[1] : ql == 2 ? bg.ax[2] : bg.ax[3];
    const int soff = ql == 0 ? G.smem_axis_off[1][0] : ql == 1 ? G.smem_axis_off[1][1] : ql == 2 ? G.smem_axis_off[1][2] : G.smem_axis_off[1][3];
    bool inq = iso_in_bounds(axq, xq);
    double yq = 0.0;
    int iq = 0;
    if (inq) iq = iso_axis_locate(axq, s_nodes + soff, xq, yq);
    lnp_eep = iso_quad_bcast(lnp_eep, q0);
    int idx4[4];
    double y4[4];
    int bc_in = 1;
#pragma unroll
    for (int d = 0; d < 4; d++) {
        idx4[d] = iso_quad_bcast(iq, q0 + d);
        y4[d] = iso_quad_bcast(yq, q0 + d);
        bc_in &= iso_quad_bcast((int)inq, q0 + d);      // `&`, not `&&`: every lane executes every shuffle
    }
    const bool bc_ok = bc_in != 0;
    if (!bc_ok) {
#pragma unroll
        for (int d = 0; d < 4; d++) {
            idx4[d] = 0;
            y4[d] = 0.0;
        }
    }
    // ---- lnprior: the sum in param_names order, as iso_lnpost_row forms it --------------------------------------------
    double lnprior = 0.0;
    if (TRACK) {
        lnprior += lnp_other;
        lnprior += lnp_eep;
    } else {
        lnprior += lnp_eep;
        lnprior += lnp_other;
    }
    lnprior += lnp_feh;
    lnprior += lnp_dist;
    lnprior += lnp_AV;
    const bool prior_ok = isfinite(lnprior);

    // ---- stage 4: likelihood.  Spectroscopic terms on lanes 0-2, parallax on lane 3; per chunk of four bands the BC cell
    // is gathered by the quad (lane s: corners 4s .. 4s + 3) and lane b forms the Gaussian term of band b ----------------
    double sterm = 0.0;
    if (ql == 0 && (m.spec_mask & 1)) sterm = iso_gauss(m.spec[0], v[ISO_MP_TEFF]);
    if (ql == 1 && (m.spec_mask & 2)) sterm = iso_gauss(m.spec[1], v[ISO_MP_LOGG]);
    if (ql == 2 && (m.spec_mask & 4)) sterm = iso_gauss(m.spec[2], v[ISO_MP_FEH]);
    if (ql == 3 && m.has_plax) sterm = iso_gauss(m.plax, 1000.0 / dist);
    const double t_teff = iso_quad_bcast(sterm, q0), t_logg = iso_quad_bcast(sterm, q0 + 1);
    const double t_feh = iso_quad_bcast(sterm, q0 + 2), t_plax = iso_quad_bcast(sterm, q0 + 3);
    double ll = 0.0;
    if (m.spec_mask & 1) ll += t_teff;
    if (m.spec_mask & 2) ll += t_logg;
    if (m.spec_mask & 4) ll += t_feh;
    if (m.obs_mask) {
        const double dist_mod = DEF ? 5.0 * fma(lnd, 0.43429448190325182765, -1.0) : 5.0 * log10(dist / 10.0);
        const int bc_chunks = bg.ncols >> 2;
        // this lane's four corners: dims 0, 1 from the lane, dims 2, 3 from the corner index
        const int b0 = ql >> 1, b1 = ql & 1;
        const double f0 = b0 ? y4[0] : (1.0 - y4[0]), f1 = b1 ? y4[1] : (1.0 - y4[1]);
        const double w01 = (1.0 * f0) * f1;
        unsigned nd4[4];
        double w4[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int b2 = j >> 1, b3 = j & 1;
            unsigned nd = (((unsigned)(idx4[0] + b0) * (unsigned)bg.n[1] + (unsigned)(idx4[1] + b1)) * (unsigned)bg.n[2] +
                           (unsigned)(idx4[2] + b2)) * (unsigned)bg.n[3] + (unsigned)(idx4[3] + b3);
            nd4[j] = min(nd, (unsigned)bg.n_nodes);
            w4[j] = (w01 * (b2 ? y4[2] : (1.0 - y4[2]))) * (b3 ? y4[3] : (1.0 - y4[3]));
        }
        for (int ch = 0; ch < bc_chunks; ch++) {
            const int cm = (m.obs_mask >> (4 * ch)) & 0xF;
            if (!cm) continue;   // uniform: the model is the same for the whole quad (and CTA)
            iso_d4 r[4];
#pragma unroll
            for (int j = 0; j < 4; j++) r[j] = iso_ldg256(bg.g + (size_t)nd4[j] * bg.ncols + 4 * ch);
            double bsum[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
            for (int s = 0; s < 4; s++) {
                double t0 = bsum[0], t1 = bsum[1], t2 = bsum[2], t3 = bsum[3];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    t0 = fma(r[j].x, w4[j], t0);
                    t1 = fma(r[j].y, w4[j], t1);
                    t2 = fma(r[j].z, w4[j], t2);
                    t3 = fma(r[j].w, w4[j], t3);
                }
                bsum[0] = iso_quad_bcast(ql == s ? t0 : bsum[0], q0 + s);
                bsum[1] = iso_quad_bcast(ql == s ? t1 : bsum[1], q0 + s);
                bsum[2] = iso_quad_bcast(ql == s ? t2 : bsum[2], q0 + s);
                bsum[3] = iso_quad_bcast(ql == s ? t3 : bsum[3], q0 + s);
            }
            // mags.py:59: Mbol + dist_mod - bc; lane b holds band b of the chunk
            const double mb = v[ISO_MP_MBOL] + dist_mod;
            const double bc_b = ql == 0 ? bsum[0] : ql == 1 ? bsum[1] : ql == 2 ? bsum[2] : bsum[3];
            const double mag_b = bc_ok ? mb - bc_b : nan;
            const IsoGaussDev &gb = ql == 0 ? m.mag[4 * ch] : ql == 1 ? m.mag[4 * ch + 1] : ql == 2 ? m.mag[4 * ch + 2] : m.mag[4 * ch + 3];
            const double bterm = iso_gauss(gb, mag_b);
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const double tb = iso_quad_bcast(bterm, q0 + b);
                if (cm & (1 << b)) ll += tb;
            }
        }
    }
    if (m.has_plax) ll += t_plax;
    return prior_ok ? lnprior + ll : neg_inf;
}

#endif  // __CUDACC__
