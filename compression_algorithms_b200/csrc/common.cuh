// Shared device/host helpers for the sm_100a kernels of libb200comp.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#define B200_OK 0
#define B200_ERR_CUDA 1
#define B200_ERR_ARG 2
#define B200_ERR_CAPACITY 3
#define B200_ERR_DOMAIN 4   /* the reference would exit(1) on this input */
#define B200_ERR_FORMAT 5   /* corrupt stream */

extern thread_local char g_b200_err[512];

#define B200_SET_ERR(...) snprintf(g_b200_err, sizeof(g_b200_err), __VA_ARGS__)

#define CUDA_TRY(expr)                                                                   \
    do {                                                                                 \
        cudaError_t e_ = (expr);                                                         \
        if (e_ != cudaSuccess) {                                                         \
            B200_SET_ERR("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e_)); \
            return B200_ERR_CUDA;                                                        \
        }                                                                                \
    } while (0)

#define B200_TRY(expr)                 \
    do {                               \
        int rc_ = (expr);              \
        if (rc_ != B200_OK) return rc_; \
    } while (0)

// One growable device scratch arena + the stream every kernel of a context runs on.
struct b200_ctx {
    int          device;
    cudaStream_t stream;
    bool         own_stream;
    int          sm_count;
    // named scratch buffers, grown on demand and kept
    static const int kSlots = 64;   // two banks of 20: the host compress path runs consecutive chunks on two kernel streams
    int          bank = 0;         // 0 or 1: which bank B200_SLOT() names (only host_api.cu switches it)
    void*        buf[kSlots];
    size_t       cap[kSlots];
    // pinned host staging for small results
    void*        pinned;
    size_t       pinned_cap;
    uint64_t     launches;  // kernels launched through this context (bench "gpu_launches")
    uint32_t     lz_epoch[2];  // last epoch tag used in each bank's LZ77 table arena (v1 path)
    // optional timing of the dominant kernel of every codec call (bench.py roofline):
    // a pool of event pairs recorded in-stream, read back after the timed region
    bool         timing;
    static const int kTimed = 512;
    cudaEvent_t  ev_a[kTimed], ev_b[kTimed];
    int          ev_kind[kTimed];
    int          ev_created, ev_used;
    // copy streams + events of the host-buffer entry points (host_api.cu): input chunks are copied
    // in on s_in, results copied out on s_out, while the codec kernels run on `stream`
    static const int kPipe = 32;
    cudaStream_t s_in, s_out, s_aux;   // s_aux: second kernel stream, so that latency-bound chunk kernels overlap
    cudaEvent_t  ev_in[kPipe], ev_done[kPipe];
    bool         pipe_ready;
    // pinned bounce rings for PAGEABLE caller buffers (malloc'd memory, what an unmodified reference driver passes):
    // ring 0 feeds host -> device copies, ring 1 drains device -> host copies; see b200_copy_in / b200_copy_out
    static const int kStage = 3;
    static const size_t kStageBytes = 16u << 20;
    void*        stage[2][kStage];
    cudaEvent_t  stage_ev[2][kStage];
    bool         stage_ready[2];
};
// Host <-> device copies that are fast for pageable host memory too: a large pageable buffer goes through a ring of
// pinned 16 MiB bounce buffers, filled / drained by several host threads while the previous piece is on the bus
// (the driver's own pageable path is a few GB/s). Pinned or small buffers are copied directly.
//   b200_copy_in : returns when the source has been consumed and every piece is enqueued on `st`
//   b200_copy_out: returns when the data is in h_dst (synchronous)
int b200_copy_in(b200_ctx* ctx, void* d_dst, const void* h_src, uint64_t bytes, cudaStream_t st);
int b200_copy_out(b200_ctx* ctx, void* h_dst, const void* d_src, uint64_t bytes, cudaStream_t st);
bool b200_is_pageable(const void* h_ptr);
int b200_pipe_init(b200_ctx* ctx);

#define B200_SLOT(ctx, s) ((s) + 20 * (ctx)->bank)
// every public entry point runs on its context's device, whatever device the calling thread had current
// (one host thread per GPU in multi.cu; a caller that created contexts on several devices)
#define B200_ENTER(ctx) do { if (ctx) { int cur_ = -1; if (cudaGetDevice(&cur_) != cudaSuccess || cur_ != (ctx)->device) CUDA_TRY(cudaSetDevice((ctx)->device)); } } while (0)
int b200_scratch(b200_ctx* ctx, int slot, size_t bytes, void** out);
int b200_pinned(b200_ctx* ctx, size_t bytes, void** out);

// bracket the dominant kernel of an API call with events when timing is enabled
#define B200_K_LZ_PARSE 0
#define B200_K_LZ_DECODE 1
#define B200_K_HUFF_ENCODE 2
#define B200_K_HUFF_DECODE 3
#define B200_K_FSE_ENCODE 4
#define B200_K_FSE_DECODE 5
#define B200_K_DFL_ENCODE 6
#define B200_K_DFL_DECODE 7
void b200_timed_begin(b200_ctx* ctx, int kind);
void b200_timed_end(b200_ctx* ctx);
#define B200_TIMED_BEGIN(ctx, kind) do { if ((ctx)->timing) b200_timed_begin((ctx), (kind)); } while (0)
#define B200_TIMED_END(ctx) do { if ((ctx)->timing) b200_timed_end((ctx)); } while (0)

static inline unsigned ceil_div_u64(uint64_t a, uint64_t b) { return (unsigned)((a + b - 1) / b); }

#ifdef __CUDACC__
__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }

__device__ __forceinline__ uint32_t warp_incl_scan_u32(uint32_t v) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane_id() >= (unsigned)d) v += t;
    }
    return v;
}
__device__ __forceinline__ uint64_t warp_incl_scan_u64(uint64_t v) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint64_t t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane_id() >= (unsigned)d) v += t;
    }
    return v;
}
__device__ __forceinline__ uint32_t warp_sum_u32(uint32_t v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

// One step of a 1024-thread CTA-wide scan: returns the exclusive prefix of v inside
// the tile and the tile total. warp_tot: 33 u64 of shared memory. Must be called
// by all 1024 threads.
__device__ __forceinline__ uint64_t cta_scan_step(uint64_t v, uint64_t* warp_tot, uint64_t* tile_total) {
    const uint64_t incl = warp_incl_scan_u64(v);
    const unsigned w = threadIdx.x >> 5;
    __syncthreads();  // protect warp_tot from the previous step
    if (lane_id() == 31) warp_tot[w] = incl;
    __syncthreads();
    if (w == 0) {
        const uint64_t t = warp_tot[lane_id()];
        const uint64_t ti = warp_incl_scan_u64(t);
        warp_tot[lane_id()] = ti - t;
        if (lane_id() == 31) warp_tot[32] = ti;
    }
    __syncthreads();
    *tile_total = warp_tot[32];
    return warp_tot[w] + incl - v;
}

// Exclusive scan of n values by ONE 1024-thread CTA, 16 consecutive values per thread and step
// (the loads of a step are independent, so a step costs one memory round trip for 16384 values).
// load(i) -> uint64_t value i; store(i, exclusive_prefix). Returns the total to every thread.
template <typename LoadF, typename StoreF>
__device__ __forceinline__ uint64_t cta_exscan_1024(uint64_t n, uint64_t* warp_tot, LoadF load, StoreF store) {
    constexpr int ITEMS = 16;
    uint64_t carry = 0, tot;
    for (uint64_t base = 0; base < n; base += 1024ull * ITEMS) {
        const uint64_t i0 = base + (uint64_t)threadIdx.x * ITEMS;
        uint64_t v[ITEMS], sum = 0;
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) { v[k] = i0 + k < n ? load(i0 + k) : 0; sum += v[k]; }
        uint64_t run = carry + cta_scan_step(sum, warp_tot, &tot);
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) { if (i0 + k < n) store(i0 + k, run); run += v[k]; }
        carry += tot;
    }
    return carry;
}

// murmur3-style hash of one u32 key, seed 0, no length xor, masked to 2^20 slots
// (reference: algorithms/lz77/lz77.c:13-41, algorithms/deflate/lz77.c:14-42).
__host__ __device__ __forceinline__ uint32_t lz_hash(uint32_t k) {
    k *= 0xcc9e2d51u;
    k = (k << 15) | (k >> 17);
    k *= 0x1b873593u;
    uint32_t h = k;
    h = ((h << 13) | (h >> 19)) * 5u + 0xe6546b64u;
    h ^= h >> 16;
    h *= 0x85ebca6bu;
    h ^= h >> 13;
    h *= 0xc2b2ae35u;
    h ^= h >> 16;
    return h & 0xFFFFFu;
}
#endif
