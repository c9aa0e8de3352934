// isochrones_b200 — chunked host<->device pipeline behind the host-pointer entry points of the C ABI.
//
// A batch of N rows is cut into chunks; chunk c runs H2D -> kernel -> D2H in order on copy_stream[c & 1], so the
// PCIe transfers of one chunk overlap the kernel (and the opposite-direction transfer) of the other.
#include <stdlib.h>
#include <string.h>

#include <unistd.h>

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#include "iso_common.cuh"

#define ISO_PIPE_CHUNK_ROWS (1 << 18)
#ifndef ISO_PIPE_PAGEABLE_DEFAULT
#define ISO_PIPE_PAGEABLE_DEFAULT 0
#endif
#define ISO_PIPE_REGISTER_MIN_BYTES (4 << 20)
// Calls of at most this many rows (the reference's scalar lnpost(p), one emcee half-step, ..) skip the copy engine:
// the kernel reads its rows from, and writes its results to, page-locked HOST memory directly (pinned memory is
// device-addressable under unified addressing), so the call is one launch + one stream synchronisation instead of
// H2D copy + launch + D2H copy + synchronisation.
#ifndef ISO_PIPE_ZERO_COPY_ROWS
#define ISO_PIPE_ZERO_COPY_ROWS 512
#endif

static inline int64_t align256(int64_t v) { return (v + 255) & ~(int64_t)255; }

// Host copies into / out of the pinned staging buffers (pageable caller arrays).  One core moves ~15 GB/s, PCIe takes
// 57: the copies are cut into 256 kB slices that a small PERSISTENT pool of workers (plus the calling thread) pulls
// from an atomic counter.  The pool lives for the process (creating threads per chunk costs more than the copy it
// parallelises); after a fork() the child has no workers and copies on its own thread.
class IsoCopyPool {
  public:
    explicit IsoCopyPool(unsigned n_workers) : pid_(getpid())
    {
        for (unsigned t = 0; t < n_workers; t++) workers_.emplace_back([this] { work(); });
    }
    ~IsoCopyPool()
    {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
        }
        wake_.notify_all();
        if (pid_ == getpid())
            for (auto &w : workers_) w.join();
        else
            for (auto &w : workers_) w.detach();   // forked child: the threads do not exist here
    }
    void copy(void *dst, const void *src, size_t bytes)
    {
        if (workers_.empty() || bytes < ((size_t)1 << 20) || pid_ != getpid()) {
            memcpy(dst, src, bytes);
            return;
        }
        std::lock_guard<std::mutex> one_job(job_);   // one copy at a time (contexts on several host threads)
        {
            std::lock_guard<std::mutex> lk(m_);
            dst_ = (char *)dst;
            src_ = (const char *)src;
            bytes_ = bytes;
            n_slices_ = (unsigned)((bytes + kSlice - 1) / kSlice);
            next_.store(0, std::memory_order_relaxed);
            busy_ = (unsigned)workers_.size();
            generation_++;
        }
        wake_.notify_all();
        pull();
        std::unique_lock<std::mutex> lk(m_);
        done_.wait(lk, [this] { return busy_ == 0; });
    }

  private:
    static constexpr size_t kSlice = (size_t)256 << 10;
    void pull()
    {
        for (;;) {
            const unsigned i = next_.fetch_add(1, std::memory_order_relaxed);
            if (i >= n_slices_) return;
            const size_t a = (size_t)i * kSlice;
            memcpy(dst_ + a, src_ + a, a + kSlice <= bytes_ ? kSlice : bytes_ - a);
        }
    }
    void work()
    {
        unsigned long long seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(m_);
                wake_.wait(lk, [&] { return stop_ || generation_ != seen; });
                if (stop_) return;
                seen = generation_;
            }
            pull();
            std::lock_guard<std::mutex> lk(m_);
            if (--busy_ == 0) done_.notify_one();
        }
    }
    std::vector<std::thread> workers_;
    std::mutex m_, job_;
    std::condition_variable wake_, done_;
    char *dst_ = nullptr;
    const char *src_ = nullptr;
    size_t bytes_ = 0;
    unsigned n_slices_ = 0, busy_ = 0;
    unsigned long long generation_ = 0;
    std::atomic<unsigned> next_{0};
    bool stop_ = false;
    pid_t pid_;
};

static void staged_memcpy(void *dst, const void *src, size_t bytes)
{
    static IsoCopyPool pool([] {
        const char *e = getenv("ISO_PIPE_COPY_THREADS");   // total copying threads incl. the caller
        unsigned hw = std::thread::hardware_concurrency();
        unsigned v = e ? (unsigned)atoi(e) : (hw >= 8 ? 4u : 1u);   // measured: 4 is the sweet spot on the 16-vCPU boxes
        v = v < 1 ? 1u : (v > 32 ? 32u : v);
        return v - 1;
    }());
    pool.copy(dst, src, bytes);
}

static bool is_pinned(const void *p)
{
    cudaPointerAttributes attr;
    cudaError_t e = cudaPointerGetAttributes(&attr, p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return attr.type == cudaMemoryTypeHost || attr.type == cudaMemoryTypeManaged;
}

struct PipeSlot {
    int64_t row0 = 0, n = 0;
    bool busy = false;
};

int iso_run_pipeline(iso_ctx *ctx, int64_t n_rows, const IsoPipeArray *arrays, int n_arrays, iso_pipe_launch_fn launch,
                     void *user)
{
    ISO_REQUIRE(ctx, n_arrays >= 1 && n_arrays <= ISO_PIPE_MAX_ARRAYS, "pipeline: bad array count");
    if (n_rows <= 0) return ISO_OK;
    std::lock_guard<std::recursive_mutex> lock(ctx->mu);   // one pipelined call at a time per context
    IsoDeviceGuard guard(ctx->device);

    static const int64_t chunk_rows = [] {
        const char *e = getenv("ISO_PIPE_CHUNK_ROWS");   // tuning knob; the default is what bench.py measures
        long long v = e ? atoll(e) : 0;
        return (int64_t)(v >= 1024 ? v : ISO_PIPE_CHUNK_ROWS);
    }();
    // how caller buffers that are NOT page-locked travel: 0 = copy through the context's pinned staging buffers,
    // 1 = hand the pageable pointer to cudaMemcpyAsync (the driver stages it), 2 = page-lock the caller's array for the
    // duration of the call (cudaHostRegister) and DMA it directly
    static const int pageable_mode = [] {
        const char *e = getenv("ISO_PIPE_PAGEABLE");
        return e ? atoi(e) : ISO_PIPE_PAGEABLE_DEFAULT;
    }();
    static const int64_t zero_copy_rows = [] {
        const char *e = getenv("ISO_PIPE_ZERO_COPY_ROWS");
        return (int64_t)(e ? atoll(e) : ISO_PIPE_ZERO_COPY_ROWS);
    }();
    if (n_rows <= zero_copy_rows) {
        // small-call path: caller buffers that are page-locked are handed to the kernel as they are, pageable ones go
        // through the context's pinned staging buffer (one host memcpy each way); ordered on the compute stream
        int64_t soff[ISO_PIPE_MAX_ARRAYS + 1];
        bool direct[ISO_PIPE_MAX_ARRAYS];
        soff[0] = 0;
        for (int k = 0; k < n_arrays; k++) {
            const void *h = arrays[k].h_in ? arrays[k].h_in : arrays[k].h_out;
            direct[k] = h != nullptr && is_pinned(h);
            soff[k + 1] = soff[k] + ((h != nullptr && !direct[k]) ? align256(n_rows * arrays[k].row_bytes) : 0);
        }
        int rc = iso_stage_reserve(ctx, 0, 0, soff[n_arrays]);
        if (rc != ISO_OK) return rc;
        void *d_arrays[ISO_PIPE_MAX_ARRAYS];
        for (int k = 0; k < n_arrays; k++) {
            const void *h = arrays[k].h_in ? arrays[k].h_in : arrays[k].h_out;
            d_arrays[k] = nullptr;
            if (!h) continue;
            if (direct[k]) {
                d_arrays[k] = const_cast<void *>(h);
            } else {
                d_arrays[k] = (char *)ctx->h_stage[0] + soff[k];
                if (arrays[k].h_in) memcpy(d_arrays[k], arrays[k].h_in, (size_t)(n_rows * arrays[k].row_bytes));
            }
        }
        rc = launch(ctx, ctx->stream, d_arrays, 0, n_rows, user);
        if (rc != ISO_OK) return rc;
        ISO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        for (int k = 0; k < n_arrays; k++)
            if (d_arrays[k] && arrays[k].h_out && !direct[k])
                memcpy(arrays[k].h_out, d_arrays[k], (size_t)(n_rows * arrays[k].row_bytes));
        return ISO_OK;
    }
    int64_t chunk = n_rows < chunk_rows ? n_rows : chunk_rows;
    int64_t off[ISO_PIPE_MAX_ARRAYS + 1];
    bool pinned[ISO_PIPE_MAX_ARRAYS];
    bool active[ISO_PIPE_MAX_ARRAYS];
    off[0] = 0;
    bool any_pageable = false;
    void *registered[ISO_PIPE_MAX_ARRAYS];
    for (int k = 0; k < n_arrays; k++) {
        const void *h = arrays[k].h_in ? arrays[k].h_in : arrays[k].h_out;
        active[k] = (h != nullptr);
        pinned[k] = active[k] && is_pinned(h);
        registered[k] = nullptr;
        if (active[k] && !pinned[k]) {
            const size_t bytes = (size_t)(n_rows * arrays[k].row_bytes);
            if (pageable_mode == 1) {
                pinned[k] = true;   // treated like a pinned pointer: the driver stages pageable memory itself
            } else if (pageable_mode == 2 && bytes >= ISO_PIPE_REGISTER_MIN_BYTES &&
                       cudaHostRegister(const_cast<void *>(h), bytes, cudaHostRegisterDefault) == cudaSuccess) {
                registered[k] = const_cast<void *>(h);
                pinned[k] = true;
            } else {
                cudaGetLastError();
            }
        }
        if (active[k] && !pinned[k]) any_pageable = true;
        off[k + 1] = off[k] + (active[k] ? align256(chunk * arrays[k].row_bytes) : 0);
    }
    struct Unregister {
        void **r;
        int n;
        ~Unregister()
        {
            for (int k = 0; k < n; k++)
                if (r[k]) cudaHostUnregister(r[k]);
        }
    } unregister{registered, n_arrays};
    int n_slots = n_rows > chunk ? 2 : 1;
    for (int s = 0; s < n_slots; s++) {
        int rc = iso_stage_reserve(ctx, s, off[n_arrays], any_pageable ? off[n_arrays] : 0);
        if (rc != ISO_OK) return rc;
    }
    // everything queued on the compute stream so far (grid staging, device-API calls) happens-before the pipeline
    ISO_CUDA(ctx, cudaEventRecord(ctx->ev_copy[0], ctx->stream));
    for (int s = 0; s < n_slots; s++) ISO_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream[s], ctx->ev_copy[0], 0));

    PipeSlot slots[2];
    // A slot is reused by chunk c + 2 on the SAME stream, so the device side needs no host wait; the host waits only
    // where it has work of its own on the slot's page-locked staging buffer (pageable caller arrays), and at the end.
    // With page-locked caller arrays every chunk is queued at once.
    auto finalize = [&](int s, bool last) -> int {
        if (!slots[s].busy) return ISO_OK;
        if (!any_pageable && !last) return ISO_OK;
        ISO_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream[s]));
        for (int k = 0; k < n_arrays; k++) {
            if (!active[k] || pinned[k] || !arrays[k].h_out) continue;
            staged_memcpy((char *)arrays[k].h_out + slots[s].row0 * arrays[k].row_bytes, (char *)ctx->h_stage[s] + off[k],
                          (size_t)(slots[s].n * arrays[k].row_bytes));
        }
        slots[s].busy = false;
        return ISO_OK;
    };

    // Chunk schedule: full chunks while more than one chunk of rows remains, then halves of what is left (down to
    // ISO_PIPE_TAIL_ROWS): the transfers are the bottleneck, and whatever the LAST chunk's kernel and result copy take
    // cannot overlap anything — so the last chunk is made small, but not too small: every chunk costs ~11 us of copy-engine
    // idle time (measured: 1e6 rows, 40 B in / 8 B out per row, bare H2D 0.735 ms: tail 16 k 0.865 ms, 64 k 0.848,
    // 128 k 0.838; profiles/r2x_e2e_chunk_probe.txt).
    static const int64_t tail_rows = [] {
        const char *e = getenv("ISO_PIPE_TAIL_ROWS");
        long long v = e ? atoll(e) : 0;
        return (int64_t)(v >= 1024 ? v : 131072);
    }();
    int c = 0;
    int64_t n = 0;
    for (int64_t row0 = 0; row0 < n_rows; row0 += n, c++) {
        int s = c & (n_slots - 1);
        const int64_t left = n_rows - row0;
        if (left > chunk) n = chunk;
        else if (left > 2 * tail_rows && n_rows > chunk) n = (left / 2 + 4095) & ~(int64_t)4095;
        else n = left;
        int rc = finalize(s, false);
        if (rc != ISO_OK) return rc;
        cudaStream_t st = ctx->copy_stream[s];
        void *d_arrays[ISO_PIPE_MAX_ARRAYS];
        for (int k = 0; k < n_arrays; k++) {
            d_arrays[k] = active[k] ? (void *)((char *)ctx->d_stage[s] + off[k]) : nullptr;
            if (!active[k] || !arrays[k].h_in) continue;
            const char *src = (const char *)arrays[k].h_in + row0 * arrays[k].row_bytes;
            size_t bytes = (size_t)(n * arrays[k].row_bytes);
            if (!pinned[k]) {
                staged_memcpy((char *)ctx->h_stage[s] + off[k], src, bytes);
                src = (const char *)ctx->h_stage[s] + off[k];
            }
            ISO_CUDA(ctx, cudaMemcpyAsync(d_arrays[k], src, bytes, cudaMemcpyHostToDevice, st));
        }
        rc = launch(ctx, st, d_arrays, row0, n, user);
        if (rc != ISO_OK) return rc;
        for (int k = 0; k < n_arrays; k++) {
            if (!active[k] || !arrays[k].h_out) continue;
            size_t bytes = (size_t)(n * arrays[k].row_bytes);
            char *dst = pinned[k] ? (char *)arrays[k].h_out + row0 * arrays[k].row_bytes : (char *)ctx->h_stage[s] + off[k];
            ISO_CUDA(ctx, cudaMemcpyAsync(dst, d_arrays[k], bytes, cudaMemcpyDeviceToHost, st));
        }
        slots[s].row0 = row0;
        slots[s].n = n;
        slots[s].busy = true;
    }
    for (int s = 0; s < n_slots; s++) {
        int rc = finalize(s, true);
        if (rc != ISO_OK) return rc;
    }
    return ISO_OK;
}
