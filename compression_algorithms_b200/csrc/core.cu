// Context, scratch memory and copy helpers of libb200comp.so.
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
