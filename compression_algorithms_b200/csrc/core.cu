// Context, scratch memory and copy helpers of libb200comp.so.
#include <cstdlib>
#include <thread>
#include <vector>
#include "common.cuh"
#include "../../include/b200comp.h"

thread_local char g_b200_err[512] = "";

extern "C" const char* b200_last_error(void) { return g_b200_err; }

extern "C" int b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" int b200_ctx_create(b200_ctx** out, int device, void* cuda_stream) {
    if (!out) { B200_SET_ERR("b200_ctx_create: out is NULL"); return B200_ERR_ARG; }
    int n = 0;
    CUDA_TRY(cudaGetDeviceCount(&n));
    if (n <= 0 || device < 0 || device >= n) {
        B200_SET_ERR("b200_ctx_create: no CUDA device %d (found %d); there is no CPU fallback", device, n);
        return B200_ERR_CUDA;
    }
    CUDA_TRY(cudaSetDevice(device));
    b200_ctx* c = new b200_ctx();
    memset(c, 0, sizeof(*c));
    c->device = device;
    if (cuda_stream) { c->stream = (cudaStream_t)cuda_stream; c->own_stream = false; }
    else { CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)); c->own_stream = true; }
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    *out = c;
    return B200_OK;
}

extern "C" void b200_ctx_destroy(b200_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (int i = 0; i < b200_ctx::kSlots; ++i) if (ctx->buf[i]) cudaFree(ctx->buf[i]);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    for (int i = 0; i < ctx->ev_created; ++i) { cudaEventDestroy(ctx->ev_a[i]); cudaEventDestroy(ctx->ev_b[i]); }
    if (ctx->pipe_ready) {
        for (int i = 0; i < b200_ctx::kPipe; ++i) { cudaEventDestroy(ctx->ev_in[i]); cudaEventDestroy(ctx->ev_done[i]); }
        cudaStreamDestroy(ctx->s_in); cudaStreamDestroy(ctx->s_out); cudaStreamDestroy(ctx->s_aux);
    }
    for (int r = 0; r < 2; ++r) if (ctx->stage_ready[r]) for (int k = 0; k < b200_ctx::kStage; ++k) { cudaFreeHost(ctx->stage[r][k]); cudaEventDestroy(ctx->stage_ev[r][k]); }
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int b200_pipe_init(b200_ctx* ctx) {
    if (ctx->pipe_ready) return B200_OK;
    CUDA_TRY(cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&ctx->s_aux, cudaStreamNonBlocking));
    for (int i = 0; i < b200_ctx::kPipe; ++i) {
        CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_in[i], cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_done[i], cudaEventDisableTiming));
    }
    ctx->pipe_ready = true;
    return B200_OK;
}

extern "C" int b200_ctx_sync(b200_ctx* ctx) {
    B200_ENTER(ctx);
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return B200_OK;
}

void b200_timed_begin(b200_ctx* ctx, int kind) {
    if (ctx->ev_used >= b200_ctx::kTimed) return;
    const int i = ctx->ev_used;
    if (i >= ctx->ev_created) {
        if (cudaEventCreate(&ctx->ev_a[i]) != cudaSuccess || cudaEventCreate(&ctx->ev_b[i]) != cudaSuccess) return;
        ctx->ev_created = i + 1;
    }
    ctx->ev_kind[i] = kind;
    cudaEventRecord(ctx->ev_a[i], ctx->stream);
}
void b200_timed_end(b200_ctx* ctx) {
    if (ctx->ev_used >= ctx->ev_created) return;
    cudaEventRecord(ctx->ev_b[ctx->ev_used], ctx->stream);
    ++ctx->ev_used;
}

extern "C" int b200_ctx_set_timing(b200_ctx* ctx, int enable) {
    B200_ENTER(ctx);
    ctx->timing = enable != 0;
    ctx->ev_used = 0;   // (re)start collecting
    return B200_OK;
}

extern "C" int b200_ctx_timing_count(b200_ctx* ctx) { return ctx->ev_used; }

extern "C" int b200_ctx_timing_get(b200_ctx* ctx, int i, int* kind, float* ms) {
    B200_ENTER(ctx);
    if (i < 0 || i >= ctx->ev_used) { B200_SET_ERR("timing entry %d out of range", i); return B200_ERR_ARG; }
    CUDA_TRY(cudaEventSynchronize(ctx->ev_b[i]));
    CUDA_TRY(cudaEventElapsedTime(ms, ctx->ev_a[i], ctx->ev_b[i]));
    *kind = ctx->ev_kind[i];
    return B200_OK;
}

extern "C" uint64_t b200_ctx_launches(b200_ctx* ctx) { return ctx ? ctx->launches : 0; }

int b200_scratch(b200_ctx* ctx, int slot, size_t bytes, void** out) {
    if (slot < 0 || slot >= b200_ctx::kSlots) { B200_SET_ERR("bad scratch slot %d", slot); return B200_ERR_ARG; }
    if (ctx->cap[slot] < bytes) {
        if (ctx->buf[slot]) {
            CUDA_TRY(cudaStreamSynchronize(ctx->stream));
            CUDA_TRY(cudaFree(ctx->buf[slot]));
            ctx->buf[slot] = nullptr; ctx->cap[slot] = 0;
        }
        size_t want = bytes + (bytes >> 3) + 256;
        CUDA_TRY(cudaMalloc(&ctx->buf[slot], want));
        ctx->cap[slot] = want;
    }
    *out = ctx->buf[slot];
    return B200_OK;
}

int b200_pinned(b200_ctx* ctx, size_t bytes, void** out) {
    if (ctx->pinned_cap < bytes) {
        if (ctx->pinned) { CUDA_TRY(cudaStreamSynchronize(ctx->stream)); CUDA_TRY(cudaFreeHost(ctx->pinned)); ctx->pinned = nullptr; }
        CUDA_TRY(cudaMallocHost(&ctx->pinned, bytes + 256));
        ctx->pinned_cap = bytes + 256;
    }
    *out = ctx->pinned;
    return B200_OK;
}

extern "C" int b200_dev_alloc(void** d_ptr, uint64_t bytes) { CUDA_TRY(cudaMalloc(d_ptr, bytes ? bytes : 1)); return B200_OK; }
extern "C" int b200_dev_free(void* d_ptr) { CUDA_TRY(cudaFree(d_ptr)); return B200_OK; }
extern "C" int b200_host_alloc(void** h_ptr, uint64_t bytes) { CUDA_TRY(cudaMallocHost(h_ptr, bytes ? bytes : 1)); return B200_OK; }
extern "C" int b200_host_free(void* h_ptr) { CUDA_TRY(cudaFreeHost(h_ptr)); return B200_OK; }
extern "C" int b200_copy_h2d(b200_ctx* ctx, void* d_dst, const void* h_src, uint64_t bytes) {
    B200_ENTER(ctx);
    CUDA_TRY(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, ctx->stream)); return B200_OK;
}
extern "C" int b200_copy_d2h(b200_ctx* ctx, void* h_dst, const void* d_src, uint64_t bytes) {
    B200_ENTER(ctx);
    CUDA_TRY(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream)); return B200_OK;
}
extern "C" int b200_memset(b200_ctx* ctx, void* d_dst, int value, uint64_t bytes) {
    B200_ENTER(ctx);
    CUDA_TRY(cudaMemsetAsync(d_dst, value, bytes, ctx->stream)); return B200_OK;
}


// ---------------------------------------------------------------- pageable host buffers
bool b200_is_pageable(const void* h_ptr) {
    if (!h_ptr) return false;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, h_ptr) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeUnregistered;
}

namespace {
const uint64_t kStageMin = [] { const char* e = getenv("B200_STAGE_MIN_BYTES"); return e ? (uint64_t)atoll(e) : (4ull << 20); }();
const int kCopyThreads = [] { const char* e = getenv("B200_COPY_THREADS"); const int v = e ? atoi(e) : 0; return v >= 1 && v <= 32 ? v : 4; }();

void par_memcpy(void* dst, const void* src, size_t bytes) {
    const int nt = bytes >= (2u << 20) ? kCopyThreads : 1;
    if (nt == 1) { memcpy(dst, src, bytes); return; }
    std::vector<std::thread> th;
    const size_t per = ((bytes + nt - 1) / nt + 4095) & ~(size_t)4095;
    for (int t = 1; t < nt; ++t) {
        const size_t o = (size_t)t * per;
        if (o >= bytes) break;
        const size_t len = o + per < bytes ? per : bytes - o;
        th.emplace_back([=] { memcpy(static_cast<uint8_t*>(dst) + o, static_cast<const uint8_t*>(src) + o, len); });
    }
    memcpy(dst, src, per < bytes ? per : bytes);
    for (auto& t : th) t.join();
}

int stage_init(b200_ctx* ctx, int ring) {
    if (ctx->stage_ready[ring]) return B200_OK;
    for (int k = 0; k < b200_ctx::kStage; ++k) {
        CUDA_TRY(cudaMallocHost(&ctx->stage[ring][k], b200_ctx::kStageBytes));
        CUDA_TRY(cudaEventCreateWithFlags(&ctx->stage_ev[ring][k], cudaEventDisableTiming));
    }
    ctx->stage_ready[ring] = true;
    return B200_OK;
}
}  // namespace

int b200_copy_in(b200_ctx* ctx, void* d_dst, const void* h_src, uint64_t bytes, cudaStream_t st) {
    if (bytes == 0) return B200_OK;
    if (bytes < kStageMin || !b200_is_pageable(h_src)) { CUDA_TRY(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, st)); return B200_OK; }
    B200_TRY(stage_init(ctx, 0));
    const size_t S = b200_ctx::kStageBytes;
    uint64_t k = 0;
    for (uint64_t off = 0; off < bytes; off += S, ++k) {
        const int slot = (int)(k % b200_ctx::kStage);
        const size_t len = off + S < bytes ? S : (size_t)(bytes - off);
        if (k >= (uint64_t)b200_ctx::kStage) CUDA_TRY(cudaEventSynchronize(ctx->stage_ev[0][slot]));   // its previous piece has left the bounce buffer
        par_memcpy(ctx->stage[0][slot], static_cast<const uint8_t*>(h_src) + off, len);
        CUDA_TRY(cudaMemcpyAsync(static_cast<uint8_t*>(d_dst) + off, ctx->stage[0][slot], len, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaEventRecord(ctx->stage_ev[0][slot], st));
    }
    // the bounce buffers still hold the last pieces: the next call's first cudaEventSynchronize would not know, so wait here
    for (int slot = 0; slot < b200_ctx::kStage; ++slot) if ((uint64_t)slot < k) CUDA_TRY(cudaEventSynchronize(ctx->stage_ev[0][slot]));
    return B200_OK;
}

int b200_copy_out(b200_ctx* ctx, void* h_dst, const void* d_src, uint64_t bytes, cudaStream_t st) {
    if (bytes == 0) return B200_OK;
    if (bytes < kStageMin || !b200_is_pageable(h_dst)) {
        CUDA_TRY(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        return B200_OK;
    }
    B200_TRY(stage_init(ctx, 1));
    const size_t S = b200_ctx::kStageBytes;
    const uint64_t npieces = (bytes + S - 1) / S;
    // piece k is on the bus while piece k - 1 is copied out of its bounce buffer by the host threads
    for (uint64_t k = 0; k <= npieces; ++k) {
        if (k < npieces) {
            const int slot = (int)(k % b200_ctx::kStage);
            const uint64_t off = k * S;
            const size_t len = off + S < bytes ? S : (size_t)(bytes - off);
            CUDA_TRY(cudaMemcpyAsync(ctx->stage[1][slot], static_cast<const uint8_t*>(d_src) + off, len, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaEventRecord(ctx->stage_ev[1][slot], st));
        }
        if (k >= 1) {
            const uint64_t j = k - 1;
            const int slot = (int)(j % b200_ctx::kStage);
            const uint64_t off = j * S;
            const size_t len = off + S < bytes ? S : (size_t)(bytes - off);
            CUDA_TRY(cudaEventSynchronize(ctx->stage_ev[1][slot]));
            par_memcpy(static_cast<uint8_t*>(h_dst) + off, ctx->stage[1][slot], len);
        }
    }
    return B200_OK;
}
