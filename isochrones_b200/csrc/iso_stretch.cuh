// isochrones_b200 — the stretch move of the affine-invariant ensemble sampler (Goodman & Weare 2010; emcee's default
// move, which the reference drives at starmodel.py:966), shared by the one-GPU persistent sampler (iso_sampler.cu) and
// the multi-GPU sharded ensemble (iso_ensemble.cu) so that both produce the SAME chain bit for bit:
//     z = ((a - 1) u + 1)^2 / a,   q = c_j - z (c_j - x_k),   accept iff  ln u' < (ndim - 1) ln z + lnpost(q) - lnpost(x_k)
// with c_j drawn uniformly from the complementary half.  All randomness is Philox4x32-10 keyed by the run's seed with
// counter (half-step, chain, walker): a proposal depends on nothing but the ensemble and those indices, never on which
// thread, CTA or GPU evaluates it.  The arithmetic is unfused so that a host replay (tests/helpers.py) is bit-identical.
#pragma once

#include "iso_philox.cuh"

// the random part of a proposal: the partner's index within the complementary half, the stretch z, the acceptance draw
__device__ __forceinline__ void iso_stretch_draw(unsigned long long seed, unsigned long long gstep, int half, int chain, int k,
                                                 int nhalf, double a, int &j_rel, double &z, double &u_acc)
{
    const unsigned k0 = (unsigned)(seed & 0xffffffffu), k1 = (unsigned)(seed >> 32);
    unsigned r[4], r2[4];
    const unsigned ctr0 = (unsigned)(gstep * 2 + half), ctr1 = (unsigned)((gstep * 2 + half) >> 32);
    iso_philox4x32_10(ctr0, ctr1 ^ ((unsigned)chain << 8), (unsigned)k, 0u, k0, k1, r);
    iso_philox4x32_10(ctr0, ctr1 ^ ((unsigned)chain << 8), (unsigned)k, 1u, k0, k1, r2);
    const double u = iso_u01(r[0], r[1]);
    j_rel = (int)(r[2] % (unsigned)nhalf);
    u_acc = iso_u01(r2[0], r2[1]);
    // z = ((a - 1) u + 1)^2 / a — unfused so that a host replay of the stream is bit-identical
    const double zr = __dadd_rn(__dmul_rn(a - 1.0, u), 1.0);
    z = __ddiv_rn(__dmul_rn(zr, zr), a);
}

// q = c - z (c - x), unfused
template <int NDIMP>
__device__ __forceinline__ void iso_stretch_point(const double (&c)[NDIMP], const double (&x)[NDIMP], double z, double (&q)[NDIMP])
{
#pragma unroll
    for (int d = 0; d < NDIMP; d++) q[d] = __dsub_rn(c[d], __dmul_rn(__dsub_rn(c[d], x[d]), z));
}

template <int NDIMP>
__device__ __forceinline__ void iso_stretch_propose(unsigned long long seed, unsigned long long gstep, int half, int chain, int k,
                                                    int other0, int nhalf, double a, const double *pos /* [n_walkers, NDIMP] */,
                                                    double (&q)[NDIMP], double &z, double &u_acc)
{
    int j_rel;
    iso_stretch_draw(seed, gstep, half, chain, k, nhalf, a, j_rel, z, u_acc);
    const int j = other0 + j_rel;
    double c[NDIMP], x[NDIMP];
#pragma unroll
    for (int d = 0; d < NDIMP; d++) {
        c[d] = pos[j * NDIMP + d];
        x[d] = pos[k * NDIMP + d];
    }
    iso_stretch_point<NDIMP>(c, x, z, q);
}

template <int NDIMP>
__device__ __forceinline__ bool iso_stretch_accept(double z, double u_acc, double lnpost_new, double lnpost_old)
{
    const double lnpdiff = (NDIMP - 1) * log(z) + lnpost_new - lnpost_old;
    return lnpdiff > log(u_acc);   // NaN compares false: a NaN lnpost (BC grid out of range) is a rejection
}
