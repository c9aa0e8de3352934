// isochrones_b200 — scratch device memory that lives for one host call (a run's chain buffer, a struct copy): served by
// the device's stream-ordered memory pool, which keeps up to 1 GiB of freed blocks instead of returning them to the
// driver (iso_ctx_create sets the pool's release threshold) — a cudaMalloc / cudaFree pair per call costs 2 – 25 ms on
// these boxes, 0.6 s after the process sat idle (profiles/README.md, the single-chain sampler).  The block is usable on
// any stream when iso_scratch_alloc returns; free it after the last use has been waited for.
#pragma once

#include "iso_common.cuh"

cudaError_t iso_scratch_alloc(iso_ctx *ctx, void **d_ptr, size_t bytes);
void iso_scratch_free(iso_ctx *ctx, void *d_ptr);
