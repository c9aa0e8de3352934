// isochrones_b200 — counter-based random numbers (Philox4x32-10, Salmon et al. 2011) shared by the on-device samplers
// and the on-device prior draws: a stream is a pure function of (key, counter), so runs are reproducible, independent
// of scheduling and of how rows are sharded over GPUs, and can be replayed on the host (tests/helpers.py).
#pragma once

__device__ __forceinline__ void iso_philox4x32_10(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0, unsigned k1,
                                                  unsigned (&out)[4])
{
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const unsigned hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const unsigned hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0;
    out[1] = c1;
    out[2] = c2;
    out[3] = c3;
}

// 53-bit uniform in [0, 1) from two 32-bit words
__device__ __forceinline__ double iso_u01(unsigned hi, unsigned lo)
{
    return (double)((((unsigned long long)hi << 32) | lo) >> 11) * (1.0 / 9007199254740992.0);
}

