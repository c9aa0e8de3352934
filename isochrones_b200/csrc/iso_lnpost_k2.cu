// isochrones_b200 — instantiations of the fused lnpost kernel for 2-star models (iso_lnpost_kernel.cuh)
#include "iso_lnpost_kernel.cuh"

int iso_lnpost_dispatch_2(iso_ctx *ctx, cudaStream_t st, const IsoLnpostParams &P, size_t smem, const IsoLnpostFlags &f)
{
    return iso_lnpost_dispatch<2>(ctx, st, P, smem, f);
}
