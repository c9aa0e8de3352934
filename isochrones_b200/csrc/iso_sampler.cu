// isochrones_b200 — on-device ensemble sampler.  Not built yet: the entry points exist so the
// ABI is complete and fail loudly.
#include "iso_common.cuh"

extern "C" {

int iso_sampler_create(iso_ctx *ctx, const iso_grid *, const iso_grid *, const iso_models *, int, int, const double *,
                       uint64_t, double, iso_sampler **out)
{
    if (out) *out = nullptr;
    return iso_set_error(ctx, ISO_E_UNSUPPORTED, "iso_sampler_create: the on-device sampler is not implemented yet");
}

int iso_sampler_run(iso_ctx *ctx, iso_sampler *, int, int, double *, double *)
{
    return iso_set_error(ctx, ISO_E_UNSUPPORTED, "iso_sampler_run: the on-device sampler is not implemented yet");
}

int iso_sampler_state(iso_ctx *ctx, iso_sampler *, double *, double *, int64_t *, int64_t *)
{
    return iso_set_error(ctx, ISO_E_UNSUPPORTED, "iso_sampler_state: the on-device sampler is not implemented yet");
}

int iso_sampler_destroy(iso_ctx *, iso_sampler *) { return ISO_OK; }

}  // extern "C"
