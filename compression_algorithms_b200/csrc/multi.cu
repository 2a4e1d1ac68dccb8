// Multi-GPU host entry points in C (BASELINE.json north_star: "host code stays in C"): ONE buffer,
// sharded by block over the GPUs of the box -- contiguous block ranges, ceil(nblocks / G) per device, the rule of
// compression_algorithms_b200/sharding.py -- one host thread per GPU, no data-path collective. The single exchange is
// an ncclAllGather of the G shard sizes (G x u64), from which every rank derives the global offset of its shard in the
// final stream (deflate.c:47-63's concatenation across devices); the payload then goes straight to that offset of the
// caller's buffer. NCCL is bound at run time (dlopen of libnccl.so.2), so the single-GPU library has no NCCL dependency.
#include <dlfcn.h>
#include <nccl.h>
#include <cstdlib>
#include <string>
#include <thread>
#include <vector>
#include "common.cuh"
#include "../../include/b200comp.h"

struct b200_multi {
    int ndev = 0;
    int dev[16];
    b200_ctx* ctx[16];
    ncclComm_t comm[16];
    uint64_t* d_xchg[16];      // [0] = own size, [1 .. ndev] = all sizes
    void* nccl_lib = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    uint64_t allgathers = 0;   // collectives issued (tests / bench evidence)
};

namespace {
inline uint64_t lz_bs(uint64_t n, uint64_t block_size) { return (block_size == 0 || block_size > n) ? n : block_size; }

void shard_of(uint64_t nblocks, int r, int g, uint64_t* first, uint64_t* count) {
    const uint64_t per = (nblocks + g - 1) / g;
    *first = std::min<uint64_t>((uint64_t)r * per, nblocks);
    *count = std::min<uint64_t>(per, nblocks - *first);
}

struct RankErr { int rc = B200_OK; std::string msg; };
#define RANK_TRY(expr) do { int rc__ = (expr); if (rc__ != B200_OK) { e.rc = rc__; e.msg = b200_last_error(); return; } } while (0)
#define RANK_CUDA(expr) do { cudaError_t ce__ = (expr); if (ce__ != cudaSuccess) { e.rc = B200_ERR_CUDA; e.msg = std::string(#expr) + " -> " + cudaGetErrorString(ce__); return; } } while (0)
}  // namespace

extern "C" int b200_multi_create(b200_multi** out, const int* devices, int ndev) {
    if (!out || ndev < 1 || ndev > 16) { B200_SET_ERR("b200_multi_create: 1..16 devices"); return B200_ERR_ARG; }
    b200_multi* m = new b200_multi();
    m->ndev = ndev;
    for (int r = 0; r < ndev; ++r) { m->dev[r] = devices ? devices[r] : r; m->ctx[r] = nullptr; m->comm[r] = nullptr; m->d_xchg[r] = nullptr; }
    for (int r = 0; r < ndev; ++r) {
        const int rc = b200_ctx_create(&m->ctx[r], m->dev[r], nullptr);
        if (rc != B200_OK) { for (int q = 0; q < r; ++q) b200_ctx_destroy(m->ctx[q]); delete m; return rc; }
        CUDA_TRY(cudaSetDevice(m->dev[r]));
        CUDA_TRY(cudaMalloc(&m->d_xchg[r], (size_t)(ndev + 1) * 8));
    }
    if (ndev > 1) {
        m->nccl_lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!m->nccl_lib) { B200_SET_ERR("b200_multi_create: libnccl.so.2 is not loadable (%s); the multi-GPU path needs NCCL for its size exchange", dlerror()); return B200_ERR_CUDA; }
        m->CommInitAll = reinterpret_cast<decltype(m->CommInitAll)>(dlsym(m->nccl_lib, "ncclCommInitAll"));
        m->CommDestroy = reinterpret_cast<decltype(m->CommDestroy)>(dlsym(m->nccl_lib, "ncclCommDestroy"));
        m->AllGather = reinterpret_cast<decltype(m->AllGather)>(dlsym(m->nccl_lib, "ncclAllGather"));
        m->GetErrorString = reinterpret_cast<decltype(m->GetErrorString)>(dlsym(m->nccl_lib, "ncclGetErrorString"));
        if (!m->CommInitAll || !m->CommDestroy || !m->AllGather || !m->GetErrorString) { B200_SET_ERR("b200_multi_create: libnccl.so.2 lacks a symbol"); return B200_ERR_CUDA; }
        const ncclResult_t nr = m->CommInitAll(m->comm, ndev, m->dev);
        if (nr != ncclSuccess) { B200_SET_ERR("ncclCommInitAll: %s", m->GetErrorString(nr)); return B200_ERR_CUDA; }
    }
    *out = m;
    return B200_OK;
}

extern "C" void b200_multi_destroy(b200_multi* m) {
    if (!m) return;
    for (int r = 0; r < m->ndev; ++r) {
        cudaSetDevice(m->dev[r]);
        if (m->comm[r] && m->CommDestroy) m->CommDestroy(m->comm[r]);
        if (m->d_xchg[r]) cudaFree(m->d_xchg[r]);
        if (m->ctx[r]) b200_ctx_destroy(m->ctx[r]);
    }
    delete m;
}

extern "C" int b200_multi_device_count(b200_multi* m) { return m ? m->ndev : 0; }
extern "C" uint64_t b200_multi_allgathers(b200_multi* m) { return m ? m->allgathers : 0; }
extern "C" uint64_t b200_multi_launches(b200_multi* m) {
    uint64_t s = 0;
    if (m) for (int r = 0; r < m->ndev; ++r) s += b200_ctx_launches(m->ctx[r]);
    return s;
}

// lz77_compress per block on a fresh table over all devices (variant 0 / 1 as in b200_lz77_encode_dev); output layout
// identical to b200_lz77_compress_host on one device: tokens back to back in h_out, h_block_sizes[nblocks],
// h_block_off[nblocks + 1].
extern "C" int b200_lz77_compress_multi_host(b200_multi* m, int variant, const uint8_t* h_in, uint64_t n, uint64_t block_size,
                                             uint8_t* h_out, uint64_t out_capacity, uint64_t* h_block_sizes, uint64_t* h_block_off,
                                             uint64_t* h_total_bytes) {
    if (!m) { B200_SET_ERR("b200_lz77_compress_multi_host: NULL handle"); return B200_ERR_ARG; }
    if (n == 0) { if (h_block_off) h_block_off[0] = 0; if (h_total_bytes) *h_total_bytes = 0; return B200_OK; }
    if (variant != 0 && variant != 1) { B200_SET_ERR("lz77: variant must be 0 or 1"); return B200_ERR_ARG; }
    const uint64_t bs = lz_bs(n, block_size);
    const uint64_t nblocks = (n + bs - 1) / bs;
    const int G = m->ndev;
    std::vector<RankErr> errs(G);
    std::vector<uint64_t> totals(G, 0);
    auto work = [&](int r) {
        RankErr& e = errs[r];
        b200_ctx* ctx = m->ctx[r];
        RANK_CUDA(cudaSetDevice(m->dev[r]));
        uint64_t b0, nb;
        shard_of(nblocks, r, G, &b0, &nb);
        const uint64_t start = std::min(b0 * bs, n), end = std::min((b0 + nb) * bs, n), len = end - start;
        const uint64_t cap = b200_lz77_max_bytes(variant, len, bs);
        uint8_t *d_in = nullptr, *d_out = nullptr; uint64_t* d_idx = nullptr;
        uint64_t total = 0;
        if (len) {
            RANK_TRY(b200_scratch(ctx, 8, len + 64, reinterpret_cast<void**>(&d_in)));
            RANK_TRY(b200_scratch(ctx, 9, cap + 64, reinterpret_cast<void**>(&d_out)));
            RANK_TRY(b200_scratch(ctx, 10, (2 * nb + 2) * 8, reinterpret_cast<void**>(&d_idx)));
            RANK_TRY(b200_copy_in(ctx, d_in, h_in + start, len, ctx->stream));
            RANK_TRY(b200_lz77_encode_dev(ctx, variant, d_in, len, bs, d_out, cap, d_idx, d_idx + nb, &total));
        }
        // the one exchange of the path: every rank learns every shard's size
        uint64_t all[17]; all[0] = total;
        if (G > 1) {
            RANK_CUDA(cudaMemcpyAsync(m->d_xchg[r], all, 8, cudaMemcpyHostToDevice, ctx->stream));
            const ncclResult_t nr = m->AllGather(m->d_xchg[r], m->d_xchg[r] + 1, 1, ncclUint64, m->comm[r], ctx->stream);
            if (nr != ncclSuccess) { e.rc = B200_ERR_CUDA; e.msg = std::string("ncclAllGather: ") + m->GetErrorString(nr); return; }
            RANK_CUDA(cudaMemcpyAsync(all + 1, m->d_xchg[r] + 1, (size_t)G * 8, cudaMemcpyDeviceToHost, ctx->stream));
            RANK_CUDA(cudaStreamSynchronize(ctx->stream));
        } else all[1] = total;
        uint64_t base = 0, sum = 0;
        for (int q = 0; q < G; ++q) { if (q < r) base += all[1 + q]; sum += all[1 + q]; }
        totals[r] = sum;
        if (sum > out_capacity) { e.rc = B200_ERR_CAPACITY; e.msg = "lz77: the output buffer is smaller than the stream"; return; }
        if (len) {
            RANK_TRY(b200_copy_out(ctx, h_out + base, d_out, total, ctx->stream));
            if (h_block_sizes) RANK_CUDA(cudaMemcpyAsync(h_block_sizes + b0, d_idx, nb * 8, cudaMemcpyDeviceToHost, ctx->stream));
            if (h_block_off) RANK_CUDA(cudaMemcpyAsync(h_block_off + b0, d_idx + nb, nb * 8, cudaMemcpyDeviceToHost, ctx->stream));
            RANK_CUDA(cudaStreamSynchronize(ctx->stream));
            if (h_block_off) for (uint64_t j = 0; j < nb; ++j) h_block_off[b0 + j] += base;   // shard-local -> global offsets
        }
    };
    std::vector<std::thread> th;
    for (int r = 1; r < G; ++r) th.emplace_back(work, r);
    work(0);
    for (auto& t : th) t.join();
    if (G > 1) ++m->allgathers;
    for (int r = 0; r < G; ++r) if (errs[r].rc != B200_OK) { B200_SET_ERR("rank %d (device %d): %s", r, m->dev[r], errs[r].msg.c_str()); return errs[r].rc; }
    if (h_block_off) h_block_off[nblocks] = totals[0];
    if (h_total_bytes) *h_total_bytes = totals[0];
    return B200_OK;
}

// the inverse: every device decodes its block range from the caller's stream; no collective (all sizes are in the index)
extern "C" int b200_lz77_decompress_multi_host(b200_multi* m, int variant, const uint8_t* h_stream, uint64_t stream_bytes,
                                               const uint64_t* h_block_off, const uint64_t* h_block_sizes, uint64_t n,
                                               uint64_t block_size, uint8_t* h_out) {
    if (!m) { B200_SET_ERR("b200_lz77_decompress_multi_host: NULL handle"); return B200_ERR_ARG; }
    if (n == 0) return B200_OK;
    if (variant != 0 && variant != 1) { B200_SET_ERR("lz77: variant must be 0 or 1"); return B200_ERR_ARG; }
    const uint64_t bs = lz_bs(n, block_size);
    const uint64_t nblocks = (n + bs - 1) / bs;
    for (uint64_t b = 0; b < nblocks; ++b) {
        const uint64_t o0 = h_block_off[b], o1 = b + 1 < nblocks ? h_block_off[b + 1] : stream_bytes;
        const uint64_t need = variant == B200_LZ_DEFLATE ? h_block_sizes[b] : h_block_sizes[b] / 8 + 1;
        if (o0 > o1 || o1 > stream_bytes || need > o1 - o0) { B200_SET_ERR("lz77: block %llu of the index does not fit the stream", (unsigned long long)b); return B200_ERR_FORMAT; }
    }
    const int G = m->ndev;
    std::vector<RankErr> errs(G);
    auto work = [&](int r) {
        RankErr& e = errs[r];
        b200_ctx* ctx = m->ctx[r];
        RANK_CUDA(cudaSetDevice(m->dev[r]));
        uint64_t b0, nb;
        shard_of(nblocks, r, G, &b0, &nb);
        if (!nb) return;
        const uint64_t start = b0 * bs, end = std::min((b0 + nb) * bs, n), len = end - start;
        const uint64_t s0 = h_block_off[b0], s1 = b0 + nb < nblocks ? h_block_off[b0 + nb] : stream_bytes;
        uint8_t *d_stream, *d_out; uint64_t* d_idx;
        RANK_TRY(b200_scratch(ctx, 9, s1 - s0 + 64, reinterpret_cast<void**>(&d_stream)));
        RANK_TRY(b200_scratch(ctx, 8, len + 64, reinterpret_cast<void**>(&d_out)));
        RANK_TRY(b200_scratch(ctx, 10, (2 * nb + 2) * 8, reinterpret_cast<void**>(&d_idx)));
        std::vector<uint64_t> loc(nb + 1);
        for (uint64_t j = 0; j < nb; ++j) loc[j] = h_block_off[b0 + j] - s0;
        loc[nb] = s1 - s0;
        RANK_TRY(b200_copy_in(ctx, d_stream, h_stream + s0, s1 - s0, ctx->stream));
        RANK_CUDA(cudaMemcpyAsync(d_idx, h_block_sizes + b0, nb * 8, cudaMemcpyHostToDevice, ctx->stream));
        RANK_CUDA(cudaMemcpyAsync(d_idx + nb, loc.data(), (nb + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
        RANK_TRY(b200_lz77_decode_dev(ctx, variant, d_stream, d_idx + nb, d_idx, len, bs, d_out));
        RANK_TRY(b200_copy_out(ctx, h_out + start, d_out, len, ctx->stream));
        RANK_CUDA(cudaStreamSynchronize(ctx->stream));
    };
    std::vector<std::thread> th;
    for (int r = 1; r < G; ++r) th.emplace_back(work, r);
    work(0);
    for (auto& t : th) t.join();
    for (int r = 0; r < G; ++r) if (errs[r].rc != B200_OK) { B200_SET_ERR("rank %d (device %d): %s", r, m->dev[r], errs[r].msg.c_str()); return errs[r].rc; }
    return B200_OK;
}
