// isochrones_b200 — instantiations of the fused lnpost kernel for 1-star models (iso_lnpost_kernel.cuh)
#include "iso_lnpost_kernel.cuh"

int iso_lnpost_dispatch_1(iso_ctx *ctx, cudaStream_t st, const IsoLnpostParams &P, size_t smem, const IsoLnpostFlags &f)
{
    return iso_lnpost_dispatch<1>(ctx, st, P, smem, f);
}
