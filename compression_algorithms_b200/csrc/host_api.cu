// Host-buffer entry points: H2D -> *_dev -> D2H on the context's stream. These are
// what the reference-named shims (csrc/shims/*.c) and bench.py's end-to-end leg call.
// Device staging buffers live in the context's scratch slots 8..13 and are reused.
#include <atomic>
#include <cstdlib>
#include <string>
#include <thread>
#include "common.cuh"
#include "../../include/b200comp.h"

namespace {
inline uint64_t lz_bs(uint64_t n, uint64_t block_size) { return (block_size == 0 || block_size > n) ? n : block_size; }
inline uint64_t lz_cap(int variant, uint64_t n, uint64_t nblocks) {
    return (variant == B200_LZ_DEFLATE ? 2 * n + 2 * nblocks : n + n / 8 + 8 * nblocks) + 64;
}

}  // namespace

extern "C" uint64_t b200_lz77_max_bytes(int variant, uint64_t n, uint64_t block_size) {
    const uint64_t bs = lz_bs(n, block_size);
    const uint64_t nblocks = n ? (n + bs - 1) / bs : 1;
    return lz_cap(variant, n, nblocks);
}

// Blocks per pipeline chunk: a multiple of the SM count (the shared-memory match finder runs one
// block per SM at a time) and at most kPipe chunks.
static uint64_t lz_chunk_blocks(const b200_ctx* ctx, uint64_t nblocks) {
    const uint64_t sms = (uint64_t)(ctx->sm_count > 0 ? ctx->sm_count : 148);
    if (nblocks < 4 * sms) return nblocks;                       // too small to be worth splitting
    static const uint64_t target = [] { const char* e = getenv("B200_LZ_CHUNKS"); const long v = e ? atol(e) : 0; return (uint64_t)(v > 0 ? v : 24); }();
    uint64_t per = (nblocks + target - 1) / target;              // aim at 24 chunks (swept on B200 with 1 GB, compress + decompress, chunks alternating between two kernel streams: 75.8 / 74.9 / 75.2 / 73.8 / 74.3 ms at 12 / 16 / 20 / 24 / 32; one stream: 78.8 / 76.0 / 76.1 / 75.9 / 76.9 at 8 / 12 / 16 / 24 / 32; B200_LZ_CHUNKS overrides, at most kPipe)
    per = (per + sms - 1) / sms * sms;
    while ((nblocks + per - 1) / per > (uint64_t)b200_ctx::kPipe) per += sms;
    return per;
}

// Compress from / to HOST buffers. The input is copied in chunk by chunk on its own stream, every
// chunk is compressed as soon as it has landed, and its tokens are copied out on a third stream
// while the next chunk is being compressed: H2D, kernels and D2H overlap.
extern "C" int b200_lz77_compress_host(b200_ctx* ctx, int variant, const uint8_t* h_in, uint64_t n, uint64_t block_size,
                                       uint8_t* h_out, uint64_t out_capacity, uint64_t* h_block_sizes,
                                       uint64_t* h_block_off, uint64_t* h_total_bytes) {
    B200_ENTER(ctx);
    if (n == 0) { if (h_block_off) h_block_off[0] = 0; if (h_total_bytes) *h_total_bytes = 0; return B200_OK; }
    if (variant != 0 && variant != 1) { B200_SET_ERR("lz77: variant must be 0 or 1"); return B200_ERR_ARG; }
    const uint64_t bs = lz_bs(n, block_size);
    const uint64_t nblocks = (n + bs - 1) / bs;
    const uint64_t cap = lz_cap(variant, n, nblocks);
    B200_TRY(b200_pipe_init(ctx));
    const uint64_t per = lz_chunk_blocks(ctx, nblocks);
    const uint64_t nchunks = (nblocks + per - 1) / per;
    uint8_t *d_in, *d_out; uint64_t* d_idx; uint64_t* pin;
    B200_TRY(b200_scratch(ctx, 8, n + 64, reinterpret_cast<void**>(&d_in)));
    B200_TRY(b200_scratch(ctx, 9, cap + 128 * (nchunks + 1), reinterpret_cast<void**>(&d_out)));
    B200_TRY(b200_scratch(ctx, 10, (2 * nblocks + 2 * nchunks + 2) * 8, reinterpret_cast<void**>(&d_idx)));
    B200_TRY(b200_pinned(ctx, 64 + nchunks * 16, reinterpret_cast<void**>(&pin)));
    uint64_t* d_sizes = d_idx;                       // [nblocks]
    uint64_t* d_boff = d_idx + nblocks;              // per chunk: blocks_c + 1 chunk-local offsets
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));    // scratch and pinned staging are free again
    // Steps 1 + 2 for one chunk: its input copy (through the pinned bounce ring when the caller's buffer is pageable)
    // and its kernels, into the chunk's own worst-case slot of d_out
    auto feed_chunk = [&](uint64_t c) -> int {
        const uint64_t b0 = c * per, nb = b0 + per < nblocks ? per : nblocks - b0;
        const uint64_t o = b0 * bs, len = o + nb * bs < n ? nb * bs : n - o;
        const uint64_t slot = lz_cap(variant, b0 * bs, b0) + 128 * c;  // worst cases are additive over chunks; 128 covers the rounding
        // consecutive chunks alternate between two kernel streams, each with its own bank of work arrays: the next chunk's
        // match finder fills the SMs that the drain of this chunk's last wave (and its scan + compaction) leave idle
        cudaStream_t keep = ctx->stream;
        if (c & 1) { ctx->stream = ctx->s_aux; ctx->bank = 1; }
        int rc2 = B200_OK;
        if (cudaStreamWaitEvent(ctx->stream, ctx->ev_in[c], 0) != cudaSuccess) rc2 = B200_ERR_CUDA;
        if (rc2 == B200_OK) rc2 = b200_lz77_encode_dev(ctx, variant, d_in + o, len, bs, d_out + slot, lz_cap(variant, len, nb), d_sizes + b0,
                                                       d_boff + b0 + c, nullptr);
        if (rc2 == B200_OK && cudaMemcpyAsync(pin + 8 + c, d_boff + b0 + c + nb, 8, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) rc2 = B200_ERR_CUDA;   // chunk total
        if (rc2 == B200_OK && cudaEventRecord(ctx->ev_done[c], ctx->stream) != cudaSuccess) rc2 = B200_ERR_CUDA;
        ctx->stream = keep; ctx->bank = 0;
        if (rc2 == B200_ERR_CUDA) B200_SET_ERR("lz77 compress: stream/event call failed: %s", cudaGetErrorString(cudaGetLastError()));
        return rc2;
    };
    const bool pageable = (b200_is_pageable(h_in) || b200_is_pageable(h_out)) && n >= (8u << 20);
    std::atomic<uint64_t> fed(0);          // chunks whose kernels are enqueued
    std::atomic<int> feed_rc(B200_OK);
    std::string feed_err;
    std::thread feeder;
    if (!pageable) {
        // 1. all input copies are queued at once; a chunk's kernels wait only for their own chunk
        for (uint64_t c = 0; c < nchunks; ++c) {
            const uint64_t o = c * per * bs, len = o + per * bs < n ? per * bs : n - o;
            CUDA_TRY(cudaMemcpyAsync(d_in + o, h_in + o, len, cudaMemcpyHostToDevice, ctx->s_in));
            CUDA_TRY(cudaEventRecord(ctx->ev_in[c], ctx->s_in));
        }
        // 2. kernels, chunk by chunk
        for (uint64_t c = 0; c < nchunks; ++c) B200_TRY(feed_chunk(c));
        fed.store(nchunks);
    } else {
        // malloc'd caller buffers (an unmodified reference driver): a helper thread feeds chunk after chunk through the
        // bounce ring and enqueues its kernels, this thread drains the finished chunks through the other ring
        feeder = std::thread([&] {
            cudaSetDevice(ctx->device);
            for (uint64_t c = 0; c < nchunks; ++c) {
                const uint64_t o = c * per * bs, len = o + per * bs < n ? per * bs : n - o;
                int rc2 = b200_copy_in(ctx, d_in + o, h_in + o, len, ctx->s_in);
                if (rc2 == B200_OK && cudaEventRecord(ctx->ev_in[c], ctx->s_in) != cudaSuccess) rc2 = B200_ERR_CUDA;
                if (rc2 == B200_OK) rc2 = feed_chunk(c);
                if (rc2 != B200_OK) { feed_err = b200_last_error(); feed_rc.store(rc2); fed.store(nchunks); return; }
                fed.store(c + 1);
            }
        });
    }
    struct Joiner { std::thread& t; ~Joiner() { if (t.joinable()) t.join(); } } joiner{feeder};
    // 3. as each chunk finishes, its tokens go home while later chunks are still being compressed
    uint64_t total = 0;
    int rc = B200_OK;
    for (uint64_t c = 0; c < nchunks; ++c) {
        const uint64_t b0 = c * per, nb = b0 + per < nblocks ? per : nblocks - b0;
        const uint64_t slot = lz_cap(variant, b0 * bs, b0) + 128 * c;
        while (fed.load() <= c) std::this_thread::yield();
        if (feed_rc.load() != B200_OK) { B200_SET_ERR("%s", feed_err.c_str()); return feed_rc.load(); }
        CUDA_TRY(cudaEventSynchronize(ctx->ev_done[c]));
        const uint64_t tc = pin[8 + c];
        if (total + tc > out_capacity) { rc = B200_ERR_CAPACITY; total += tc; continue; }
        CUDA_TRY(cudaStreamWaitEvent(ctx->s_out, ctx->ev_done[c], 0));
        if (pageable) B200_TRY(b200_copy_out(ctx, h_out + total, d_out + slot, tc, ctx->s_out));
        else CUDA_TRY(cudaMemcpyAsync(h_out + total, d_out + slot, tc, cudaMemcpyDeviceToHost, ctx->s_out));
        if (h_block_sizes) CUDA_TRY(cudaMemcpyAsync(h_block_sizes + b0, d_sizes + b0, nb * 8, cudaMemcpyDeviceToHost, ctx->s_out));
        if (h_block_off) CUDA_TRY(cudaMemcpyAsync(h_block_off + b0, d_boff + b0 + c, nb * 8, cudaMemcpyDeviceToHost, ctx->s_out));
        pin[8 + nchunks + c] = total;                // base of this chunk in the final stream
        total += tc;
    }
    CUDA_TRY(cudaStreamSynchronize(ctx->s_out));
    CUDA_TRY(cudaStreamSynchronize(ctx->s_aux));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (h_total_bytes) *h_total_bytes = total;
    if (rc != B200_OK) { B200_SET_ERR("lz77: output needs %llu bytes, buffer has %llu", (unsigned long long)total, (unsigned long long)out_capacity); return rc; }
    if (h_block_off) {                               // chunk-local offsets -> offsets in the concatenated stream
        for (uint64_t c = 1; c < nchunks; ++c) {
            const uint64_t b0 = c * per, nb = b0 + per < nblocks ? per : nblocks - b0, base = pin[8 + nchunks + c];
            for (uint64_t j = 0; j < nb; ++j) h_block_off[b0 + j] += base;
        }
        h_block_off[nblocks] = total;
    }
    return B200_OK;
}

// Decompress from / to HOST buffers with the same three-stream overlap (all sizes are known on the host).
extern "C" int b200_lz77_decompress_host(b200_ctx* ctx, int variant, const uint8_t* h_stream, uint64_t stream_bytes,
                                         const uint64_t* h_block_off, const uint64_t* h_block_sizes,
                                         uint64_t n, uint64_t block_size, uint8_t* h_out) {
    B200_ENTER(ctx);
    if (n == 0) return B200_OK;
    const uint64_t bs = lz_bs(n, block_size);
    const uint64_t nblocks = (n + bs - 1) / bs;
    if (variant != 0 && variant != 1) { B200_SET_ERR("lz77: variant must be 0 or 1"); return B200_ERR_ARG; }
    // the index comes from the caller: every block must lie inside the stream and carry the size its extent implies
    for (uint64_t b = 0; b < nblocks; ++b) {
        const uint64_t o0 = h_block_off[b], o1 = b + 1 < nblocks ? h_block_off[b + 1] : stream_bytes;
        const uint64_t need = variant == B200_LZ_DEFLATE ? h_block_sizes[b] : h_block_sizes[b] / 8 + 1;
        if (o0 > o1 || o1 > stream_bytes || need > o1 - o0) {
            B200_SET_ERR("lz77: block %llu of the index does not fit the stream (offset %llu, next %llu, size %llu, stream %llu)",
                         (unsigned long long)b, (unsigned long long)o0, (unsigned long long)o1, (unsigned long long)h_block_sizes[b], (unsigned long long)stream_bytes);
            return B200_ERR_FORMAT;
        }
    }
    B200_TRY(b200_pipe_init(ctx));
    const uint64_t per = lz_chunk_blocks(ctx, nblocks);
    const uint64_t nchunks = (nblocks + per - 1) / per;
    uint8_t *d_stream, *d_out; uint64_t* d_idx;
    B200_TRY(b200_scratch(ctx, 9, stream_bytes + 64, reinterpret_cast<void**>(&d_stream)));
    B200_TRY(b200_scratch(ctx, 8, n + 64, reinterpret_cast<void**>(&d_out)));
    B200_TRY(b200_scratch(ctx, 10, (2 * nblocks + 2) * 8, reinterpret_cast<void**>(&d_idx)));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(d_idx, h_block_sizes, nblocks * 8, cudaMemcpyHostToDevice, ctx->s_in));
    CUDA_TRY(cudaMemcpyAsync(d_idx + nblocks, h_block_off, (nblocks + 1) * 8, cudaMemcpyHostToDevice, ctx->s_in));
    const bool pageable = (b200_is_pageable(h_stream) || b200_is_pageable(h_out)) && n >= (8u << 20);
    auto chunk_in = [&](uint64_t c) -> int {     // the chunk's tokens to the device (bounce ring for a pageable source)
        const uint64_t b0 = c * per, b1 = b0 + per < nblocks ? b0 + per : nblocks;
        const uint64_t s0 = h_block_off[b0], s1 = b1 < nblocks ? h_block_off[b1] : stream_bytes;
        if (pageable) B200_TRY(b200_copy_in(ctx, d_stream + s0, h_stream + s0, s1 - s0, ctx->s_in));
        else CUDA_TRY(cudaMemcpyAsync(d_stream + s0, h_stream + s0, s1 - s0, cudaMemcpyHostToDevice, ctx->s_in));
        CUDA_TRY(cudaEventRecord(ctx->ev_in[c], ctx->s_in));
        return B200_OK;
    };
    auto chunk_kernels = [&](uint64_t c) -> int {
        const uint64_t b0 = c * per, b1 = b0 + per < nblocks ? b0 + per : nblocks;
        const uint64_t o = b0 * bs, len = b1 * bs < n ? (b1 - b0) * bs : n - o;
        // a chunk's decode is bound by the serial token chain of its blocks (about 4 ms whatever the chunk size), not
        // by throughput: consecutive chunks go to two kernel streams so that their chains overlap
        cudaStream_t keep = ctx->stream;
        const int keep_bank = ctx->bank;
        if (c & 1) { ctx->stream = ctx->s_aux; ctx->bank = 1; }   // (the token-parallel decoder has scratch arrays: one bank per stream)
        int rc = B200_OK;
        if (cudaStreamWaitEvent(ctx->stream, ctx->ev_in[c], 0) != cudaSuccess) rc = B200_ERR_CUDA;
        if (rc == B200_OK) rc = b200_lz77_decode_dev(ctx, variant, d_stream, d_idx + nblocks + b0, d_idx + b0, len, bs, d_out + o);
        if (rc == B200_OK && cudaEventRecord(ctx->ev_done[c], ctx->stream) != cudaSuccess) rc = B200_ERR_CUDA;
        ctx->stream = keep; ctx->bank = keep_bank;
        if (rc == B200_ERR_CUDA) B200_SET_ERR("lz77 decompress: stream/event call failed: %s", cudaGetErrorString(cudaGetLastError()));
        return rc;
    };
    std::atomic<uint64_t> fed(0);
    std::atomic<int> feed_rc(B200_OK);
    std::string feed_err;
    std::thread feeder;
    if (!pageable) {
        for (uint64_t c = 0; c < nchunks; ++c) B200_TRY(chunk_in(c));
    } else {
        feeder = std::thread([&] {
            cudaSetDevice(ctx->device);
            for (uint64_t c = 0; c < nchunks; ++c) {
                int rc = chunk_in(c);
                if (rc == B200_OK) rc = chunk_kernels(c);
                if (rc != B200_OK) { feed_err = b200_last_error(); feed_rc.store(rc); fed.store(nchunks); return; }
                fed.store(c + 1);
            }
        });
    }
    struct Joiner { std::thread& t; ~Joiner() { if (t.joinable()) t.join(); } } joiner{feeder};
    for (uint64_t c = 0; c < nchunks; ++c) {
        const uint64_t b0 = c * per, b1 = b0 + per < nblocks ? b0 + per : nblocks;
        const uint64_t o = b0 * bs, len = b1 * bs < n ? (b1 - b0) * bs : n - o;
        if (!pageable) B200_TRY(chunk_kernels(c));
        else {
            while (fed.load() <= c) std::this_thread::yield();
            if (feed_rc.load() != B200_OK) { B200_SET_ERR("%s", feed_err.c_str()); return feed_rc.load(); }
        }
        CUDA_TRY(cudaStreamWaitEvent(ctx->s_out, ctx->ev_done[c], 0));
        if (pageable) B200_TRY(b200_copy_out(ctx, h_out + o, d_out + o, len, ctx->s_out));
        else CUDA_TRY(cudaMemcpyAsync(h_out + o, d_out + o, len, cudaMemcpyDeviceToHost, ctx->s_out));
    }
    if (feeder.joinable()) feeder.join();
    CUDA_TRY(cudaStreamSynchronize(ctx->s_out));
    CUDA_TRY(cudaStreamSynchronize(ctx->s_aux));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return B200_OK;
}

extern "C" int b200_huffman_compress_host(b200_ctx* ctx, const uint8_t* h_in, uint64_t n, uint64_t block_size,
                                          uint32_t* h_words, uint64_t words_capacity, uint8_t* h_side, uint64_t side_bytes,
                                          uint64_t* h_total_words, uint32_t* h_worst_status) {
    B200_ENTER(ctx);
    b200_huff_layout L;
    B200_TRY(b200_huffman_layout(n, block_size, &L));
    if (h_side && side_bytes < L.bytes) { B200_SET_ERR("huffman: host side buffer too small"); return B200_ERR_CAPACITY; }
    const uint64_t cap = b200_huffman_max_words(n, block_size);
    uint8_t *d_in, *d_side; uint32_t* d_words;
    B200_TRY(b200_scratch(ctx, 8, n + 64, reinterpret_cast<void**>(&d_in)));
    B200_TRY(b200_scratch(ctx, 11, cap * 4, reinterpret_cast<void**>(&d_words)));
    B200_TRY(b200_scratch(ctx, 12, L.bytes, reinterpret_cast<void**>(&d_side)));
    B200_TRY(b200_copy_in(ctx, d_in, h_in, n, ctx->stream));
    uint64_t total = 0; uint32_t worst = 0;
    B200_TRY(b200_huffman_encode_dev(ctx, d_in, n, block_size, d_words, cap, d_side, L.bytes, &total, &worst));
    if (h_total_words) *h_total_words = total;
    if (h_worst_status) *h_worst_status = worst;
    if (worst == 0) {
        if (total > words_capacity) { B200_SET_ERR("huffman: output needs %llu words, buffer has %llu", (unsigned long long)total, (unsigned long long)words_capacity); return B200_ERR_CAPACITY; }
        B200_TRY(b200_copy_out(ctx, h_words, d_words, total * 4, ctx->stream));
    }
    if (h_side) B200_TRY(b200_copy_out(ctx, h_side, d_side, L.bytes, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (worst) { B200_SET_ERR("huffman: input has a block the reference cannot encode (status %u)", worst); return B200_ERR_DOMAIN; }
    return B200_OK;
}

extern "C" int b200_huffman_decompress_host(b200_ctx* ctx, const uint32_t* h_words, uint64_t total_words,
                                            const uint8_t* h_side, uint64_t side_bytes, uint64_t n, uint64_t block_size,
                                            uint8_t* h_out) {
    B200_ENTER(ctx);
    if (n == 0) return B200_OK;
    b200_huff_layout L;
    B200_TRY(b200_huffman_layout(n, block_size, &L));
    if (side_bytes < L.bytes) { B200_SET_ERR("huffman: host side buffer too small"); return B200_ERR_CAPACITY; }
    uint8_t *d_out, *d_side; uint32_t* d_words;
    B200_TRY(b200_scratch(ctx, 8, n + 64, reinterpret_cast<void**>(&d_out)));
    B200_TRY(b200_scratch(ctx, 11, (total_words + 4) * 4, reinterpret_cast<void**>(&d_words)));
    B200_TRY(b200_scratch(ctx, 12, L.bytes, reinterpret_cast<void**>(&d_side)));
    B200_TRY(b200_copy_in(ctx, d_words, h_words, total_words * 4, ctx->stream));
    B200_TRY(b200_copy_in(ctx, d_side, h_side, L.bytes, ctx->stream));
    B200_TRY(b200_huffman_decode_dev(ctx, d_words, total_words, d_side, L.bytes, n, block_size, d_out));
    B200_TRY(b200_copy_out(ctx, h_out, d_out, n, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return B200_OK;
}

// One-stream decode of a foreign Huffman stream from host buffers (index-free).
extern "C" int b200_huffman_decompress_serial_host(b200_ctx* ctx, const uint32_t* h_words, uint64_t nwords,
                                                   uint64_t buffer_size, const uint32_t* h_codes, const uint8_t* h_lens,
                                                   uint8_t* h_out, uint64_t out_capacity, uint64_t* h_count) {
    B200_ENTER(ctx);
    uint8_t *d_out, *d_tab; uint32_t* d_words;
    B200_TRY(b200_scratch(ctx, 8, out_capacity + 64, reinterpret_cast<void**>(&d_out)));
    B200_TRY(b200_scratch(ctx, 11, (nwords + 4) * 4, reinterpret_cast<void**>(&d_words)));
    B200_TRY(b200_scratch(ctx, 12, 2048, reinterpret_cast<void**>(&d_tab)));
    B200_TRY(b200_copy_in(ctx, d_words, h_words, nwords * 4, ctx->stream));
    B200_TRY(b200_copy_in(ctx, d_tab, h_codes, 1024, ctx->stream));
    B200_TRY(b200_copy_in(ctx, d_tab + 1024, h_lens, 256, ctx->stream));
    CUDA_TRY(cudaMemsetAsync(d_out, 0, out_capacity, ctx->stream));
    uint64_t cnt = 0;
    B200_TRY(b200_huffman_decode_serial_dev(ctx, d_words, nwords, buffer_size, reinterpret_cast<uint32_t*>(d_tab), d_tab + 1024,
                                            d_out, out_capacity, &cnt));
    B200_TRY(b200_copy_out(ctx, h_out, d_out, cnt < out_capacity ? cnt : out_capacity, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (h_count) *h_count = cnt;
    return B200_OK;
}

// build_huffman_tree + gather_codes from a host buffer: only the side buffer comes back
extern "C" int b200_huffman_tables_host(b200_ctx* ctx, const uint8_t* h_in, uint64_t n, uint64_t block_size,
                                        uint8_t* h_side, uint64_t side_bytes) {
    B200_ENTER(ctx);
    b200_huff_layout L;
    B200_TRY(b200_huffman_layout(n, block_size, &L));
    if (side_bytes < L.bytes) { B200_SET_ERR("huffman: host side buffer too small"); return B200_ERR_CAPACITY; }
    uint8_t *d_in, *d_side;
    B200_TRY(b200_scratch(ctx, 8, n + 64, reinterpret_cast<void**>(&d_in)));
    B200_TRY(b200_scratch(ctx, 12, L.bytes, reinterpret_cast<void**>(&d_side)));
    B200_TRY(b200_copy_in(ctx, d_in, h_in, n, ctx->stream));
    B200_TRY(b200_huffman_tables_dev(ctx, d_in, n, block_size, d_side, L.bytes));
    B200_TRY(b200_copy_out(ctx, h_side, d_side, L.off_block_bits, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return B200_OK;
}

// _huffman_compress from host buffers: the caller's code table, one table for the whole buffer
extern "C" int b200_huffman_compress_codes_host(b200_ctx* ctx, const uint8_t* h_in, uint64_t n,
                                                const uint32_t* h_codes, const uint8_t* h_lens,
                                                uint32_t* h_words, uint64_t words_capacity, uint8_t* h_side, uint64_t side_bytes,
                                                uint64_t* h_total_words, uint32_t* h_worst_status) {
    B200_ENTER(ctx);
    if (n == 0) { if (h_total_words) *h_total_words = 0; if (h_worst_status) *h_worst_status = 0; return B200_OK; }
    b200_huff_layout L;
    B200_TRY(b200_huffman_layout(n, 0, &L));
    if (h_side && side_bytes < L.bytes) { B200_SET_ERR("huffman: host side buffer too small"); return B200_ERR_CAPACITY; }
    // the reference allows up to 32 bits per symbol with a foreign table
    uint64_t cap = 0;
    { uint32_t mx = 0; for (int s = 0; s < 256; ++s) if (h_lens[s] > mx) mx = h_lens[s]; cap = (n * mx + 31) / 32 + 4; }
    uint8_t *d_in, *d_side; uint32_t* d_words;
    B200_TRY(b200_scratch(ctx, 8, n + 64, reinterpret_cast<void**>(&d_in)));
    B200_TRY(b200_scratch(ctx, 11, cap * 4, reinterpret_cast<void**>(&d_words)));
    B200_TRY(b200_scratch(ctx, 12, L.bytes, reinterpret_cast<void**>(&d_side)));
    B200_TRY(b200_copy_in(ctx, d_in, h_in, n, ctx->stream));
    uint64_t total = 0; uint32_t worst = 0;
    B200_TRY(b200_huffman_encode_with_codes_dev(ctx, d_in, n, h_codes, h_lens, d_words, cap, d_side, L.bytes, &total, &worst));
    if (h_total_words) *h_total_words = total;
    if (h_worst_status) *h_worst_status = worst;
    if (worst) { B200_SET_ERR("huffman: a symbol of the input has no code (status %u)", worst); return B200_ERR_DOMAIN; }
    if (total > words_capacity) { B200_SET_ERR("huffman: output needs %llu words, buffer has %llu", (unsigned long long)total, (unsigned long long)words_capacity); return B200_ERR_CAPACITY; }
    B200_TRY(b200_copy_out(ctx, h_words, d_words, total * 4, ctx->stream));
    if (h_side) B200_TRY(b200_copy_out(ctx, h_side, d_side, L.bytes, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return B200_OK;
}

// ---- FSE container: a self-describing stream so that fse_decompress needs nothing else.
//   u64 header[8] = {magic, n, block_size, seg_size, nblocks, nsegs, stream_words, 0}
//   u16 norm[nblocks][256]      normalised counts per block (main.zig:106-149)
//   u32 seg_bits[nsegs]         exact bits of every segment stream (padded to a u64)
//   u64 stream[stream_words]    the segment streams, each starting on a u64 word
namespace {
constexpr uint64_t FSE_MAGIC = 0x3130455346303042ull;   // "B00FSE01"
inline uint64_t w8(uint64_t bytes) { return (bytes + 7) / 8; }
}

extern "C" uint64_t b200_fse_container_max_words(uint64_t n, uint64_t block_size, uint64_t seg_size) {
    b200_fse_layout L;
    if (b200_fse_layout_for(n ? n : 1, block_size, seg_size, &L) != B200_OK) return 0;
    return 8 + w8(L.nblocks * 512) + w8(L.nsegs * 4) + b200_fse_max_words(n, seg_size);
}

extern "C" int b200_fse_compress_host(b200_ctx* ctx, const uint8_t* h_in, uint64_t n, uint64_t block_size, uint64_t seg_size,
                                      uint64_t* h_out, uint64_t out_capacity_words, uint64_t* h_total_words) {
    B200_ENTER(ctx);
    if (n == 0) { B200_SET_ERR("fse: empty input"); return B200_ERR_DOMAIN; }
    b200_fse_layout L;
    B200_TRY(b200_fse_layout_for(n, block_size, seg_size, &L));
    const uint64_t cap = b200_fse_max_words(n, seg_size);
    uint8_t *d_in, *d_side; uint64_t* d_words;
    B200_TRY(b200_scratch(ctx, 8, n + 64, reinterpret_cast<void**>(&d_in)));
    B200_TRY(b200_scratch(ctx, 11, cap * 8, reinterpret_cast<void**>(&d_words)));
    B200_TRY(b200_scratch(ctx, 12, L.bytes, reinterpret_cast<void**>(&d_side)));
    B200_TRY(b200_copy_in(ctx, d_in, h_in, n, ctx->stream));
    uint64_t total = 0;
    B200_TRY(b200_fse_encode_dev(ctx, d_in, n, block_size, seg_size, d_words, cap, d_side, L.bytes, &total));
    const uint64_t o_norm = 8, o_bits = o_norm + w8(L.nblocks * 512), o_stream = o_bits + w8(L.nsegs * 4);
    if (o_stream + total > out_capacity_words) { B200_SET_ERR("fse: container needs %llu words, buffer has %llu", (unsigned long long)(o_stream + total), (unsigned long long)out_capacity_words); return B200_ERR_CAPACITY; }
    const uint64_t bs_eff = L.segs_per_block * seg_size;
    const uint64_t hdr[8] = {FSE_MAGIC, n, bs_eff, seg_size, L.nblocks, L.nsegs, total, 0};
    memcpy(h_out, hdr, sizeof(hdr));
    if ((L.nsegs * 4) % 8) h_out[o_stream - 1] = 0;
    B200_TRY(b200_copy_out(ctx, h_out + o_norm, d_side + L.off_norm, L.nblocks * 512, ctx->stream));
    B200_TRY(b200_copy_out(ctx, h_out + o_bits, d_side + L.off_seg_bits, L.nsegs * 4, ctx->stream));
    B200_TRY(b200_copy_out(ctx, h_out + o_stream, d_words, total * 8, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (h_total_words) *h_total_words = o_stream + total;
    return B200_OK;
}

extern "C" int b200_fse_container_size(const uint64_t* h_container, uint64_t words, uint64_t* h_n) {
    if (!h_container || words < 8 || h_container[0] != FSE_MAGIC) { B200_SET_ERR("fse: not a B00FSE01 container"); return B200_ERR_FORMAT; }
    if (h_n) *h_n = h_container[1];
    return B200_OK;
}

extern "C" int b200_fse_decompress_host(b200_ctx* ctx, const uint64_t* h_container, uint64_t words,
                                        uint8_t* h_out, uint64_t out_capacity, uint64_t* h_n) {
    B200_ENTER(ctx);
    uint64_t n = 0;
    B200_TRY(b200_fse_container_size(h_container, words, &n));
    const uint64_t bs = h_container[2], seg = h_container[3], nblocks = h_container[4], nsegs = h_container[5], total = h_container[6];
    b200_fse_layout L;
    if (n == 0 || b200_fse_layout_for(n, bs, seg, &L) != B200_OK || L.nblocks != nblocks || L.nsegs != nsegs) {
        B200_SET_ERR("fse: inconsistent container header"); return B200_ERR_FORMAT;
    }
    const uint64_t o_norm = 8, o_bits = o_norm + w8(nblocks * 512), o_stream = o_bits + w8(nsegs * 4);
    if (o_stream + total > words) { B200_SET_ERR("fse: truncated container"); return B200_ERR_FORMAT; }
    if (n > out_capacity) { B200_SET_ERR("fse: output needs %llu bytes, buffer has %llu", (unsigned long long)n, (unsigned long long)out_capacity); return B200_ERR_CAPACITY; }
    {   // the tables and the index are read from the container: check them before any kernel trusts them
        const uint16_t* norm = reinterpret_cast<const uint16_t*>(h_container + o_norm);
        for (uint64_t b = 0; b < nblocks; ++b) {
            uint32_t sum = 0;
            for (int k = 0; k < 256; ++k) sum += norm[b * 256 + k];
            if (sum != 256) { B200_SET_ERR("fse: normalised counts of block %llu sum to %u, not 256", (unsigned long long)b, sum); return B200_ERR_FORMAT; }
        }
        const uint32_t* sb = reinterpret_cast<const uint32_t*>(h_container + o_bits);
        uint64_t words_needed = 0;
        for (uint64_t g = 0; g < nsegs; ++g) {
            // (0 bits = a segment behind the end of the input)
            if (sb[g] > 64ull * (seg + 8)) { B200_SET_ERR("fse: segment %llu claims %u bits", (unsigned long long)g, sb[g]); return B200_ERR_FORMAT; }
            words_needed += ((uint64_t)sb[g] + 63) >> 6;
        }
        if (words_needed != total) { B200_SET_ERR("fse: segment index covers %llu words, the stream has %llu", (unsigned long long)words_needed, (unsigned long long)total); return B200_ERR_FORMAT; }
    }
    uint8_t *d_out, *d_side; uint64_t* d_words;
    B200_TRY(b200_scratch(ctx, 8, n + 64, reinterpret_cast<void**>(&d_out)));
    B200_TRY(b200_scratch(ctx, 11, (total + 4) * 8, reinterpret_cast<void**>(&d_words)));
    B200_TRY(b200_scratch(ctx, 12, L.bytes, reinterpret_cast<void**>(&d_side)));
    B200_TRY(b200_copy_in(ctx, d_side + L.off_norm, h_container + o_norm, nblocks * 512, ctx->stream));
    B200_TRY(b200_copy_in(ctx, d_side + L.off_seg_bits, h_container + o_bits, nsegs * 4, ctx->stream));
    B200_TRY(b200_copy_in(ctx, d_words, h_container + o_stream, total * 8, ctx->stream));
    CUDA_TRY(cudaMemsetAsync(d_words + total, 0, 32, ctx->stream));
    B200_TRY(b200_fse_rebuild_index_dev(ctx, n, bs, seg, d_side, L.bytes));
    uint32_t bad = 0;
    B200_TRY(b200_fse_decode_dev(ctx, d_words, d_side, L.bytes, n, bs, seg, d_out, &bad));
    if (bad) { B200_SET_ERR("fse: %u corrupt segment streams", bad); return B200_ERR_FORMAT; }
    B200_TRY(b200_copy_out(ctx, h_out, d_out, n, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (h_n) *h_n = n;
    return B200_OK;
}

// histogram + normalisation of one table scope from a host buffer (buildFrequencyTable +
// normalizeFrequencyTable, main.zig:88-149): freq[256] and norm[256] of block 0 come back
extern "C" int b200_fse_normalize_host(b200_ctx* ctx, const uint8_t* h_in, uint64_t n, uint32_t* h_freq, uint16_t* h_norm) {
    B200_ENTER(ctx);
    if (n == 0) { B200_SET_ERR("fse: empty input"); return B200_ERR_DOMAIN; }
    b200_fse_layout L;
    B200_TRY(b200_fse_layout_for(n, 0, 1024, &L));
    uint8_t *d_in, *d_side;
    B200_TRY(b200_scratch(ctx, 8, n + 64, reinterpret_cast<void**>(&d_in)));
    B200_TRY(b200_scratch(ctx, 12, L.bytes, reinterpret_cast<void**>(&d_side)));
    B200_TRY(b200_copy_in(ctx, d_in, h_in, n, ctx->stream));
    B200_TRY(b200_fse_normalize_dev(ctx, d_in, n, 0, d_side, L.bytes));
    if (h_freq) B200_TRY(b200_copy_out(ctx, h_freq, d_side + L.off_freq, 1024, ctx->stream));
    if (h_norm) B200_TRY(b200_copy_out(ctx, h_norm, d_side + L.off_norm, 512, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return B200_OK;
}

// deflate with the entropy stage from / to HOST buffers (what a completed deflate.c compress() /
// decompress() pair would call): H2D -> lz77_compress per block -> entropy stage -> D2H of the words
// and the side buffer (tables + decode index; like Huffman's it is not part of the compressed words).
extern "C" int b200_deflate_compress_host(b200_ctx* ctx, const uint8_t* h_in, uint64_t n, uint64_t block_size,
                                          uint32_t* h_words, uint64_t words_capacity, uint8_t* h_side, uint64_t side_bytes,
                                          uint64_t* h_total_words, uint32_t* h_worst_status) {
    B200_ENTER(ctx);
    if (n == 0) { if (h_total_words) *h_total_words = 0; if (h_worst_status) *h_worst_status = 0; return B200_OK; }
    b200_dfl_layout L;
    B200_TRY(b200_dfl_layout_for(n, block_size, &L));
    if (!h_side || side_bytes < L.bytes) { B200_SET_ERR("deflate: host side buffer too small (%llu needed)", (unsigned long long)L.bytes); return B200_ERR_CAPACITY; }
    const uint64_t cap = b200_dfl_max_words(n, block_size);
    const uint64_t tok_cap = lz_cap(B200_LZ_DEFLATE, n, L.nblocks);
    uint8_t *d_in, *d_tok, *d_side; uint32_t* d_words; uint64_t* d_idx;
    B200_TRY(b200_scratch(ctx, 8, n + 64, reinterpret_cast<void**>(&d_in)));
    B200_TRY(b200_scratch(ctx, 9, tok_cap, reinterpret_cast<void**>(&d_tok)));
    B200_TRY(b200_scratch(ctx, 10, (2 * L.nblocks + 2) * 8, reinterpret_cast<void**>(&d_idx)));
    B200_TRY(b200_scratch(ctx, 11, cap * 4, reinterpret_cast<void**>(&d_words)));
    B200_TRY(b200_scratch(ctx, 12, L.bytes, reinterpret_cast<void**>(&d_side)));
    B200_TRY(b200_copy_in(ctx, d_in, h_in, n, ctx->stream));
    uint64_t total = 0; uint32_t worst = 0;
    B200_TRY(b200_deflate_compress_dev(ctx, d_in, n, block_size, d_tok, tok_cap, d_idx, d_idx + L.nblocks, d_words, cap,
                                       d_side, L.bytes, &total, &worst));
    if (h_total_words) *h_total_words = total;
    if (h_worst_status) *h_worst_status = worst;
    if (worst) { B200_SET_ERR("deflate: a block needs a code longer than 32 bits (status %u)", worst); return B200_ERR_DOMAIN; }
    if (total > words_capacity) { B200_SET_ERR("deflate: output needs %llu words, buffer has %llu", (unsigned long long)total, (unsigned long long)words_capacity); return B200_ERR_CAPACITY; }
    B200_TRY(b200_copy_out(ctx, h_words, d_words, total * 4, ctx->stream));
    B200_TRY(b200_copy_out(ctx, h_side, d_side, L.bytes, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return B200_OK;
}

extern "C" int b200_deflate_decompress_host(b200_ctx* ctx, const uint32_t* h_words, uint64_t total_words,
                                            const uint8_t* h_side, uint64_t side_bytes, uint64_t n, uint64_t block_size,
                                            uint8_t* h_out) {
    B200_ENTER(ctx);
    if (n == 0) return B200_OK;
    b200_dfl_layout L;
    B200_TRY(b200_dfl_layout_for(n, block_size, &L));
    if (side_bytes < L.bytes) { B200_SET_ERR("deflate: host side buffer too small"); return B200_ERR_CAPACITY; }
    const uint64_t tok_cap = lz_cap(B200_LZ_DEFLATE, n, L.nblocks);
    uint8_t *d_out, *d_tok, *d_side; uint32_t* d_words;
    B200_TRY(b200_scratch(ctx, 8, n + 64, reinterpret_cast<void**>(&d_out)));
    B200_TRY(b200_scratch(ctx, 9, tok_cap, reinterpret_cast<void**>(&d_tok)));
    B200_TRY(b200_scratch(ctx, 11, (total_words + 4) * 4, reinterpret_cast<void**>(&d_words)));
    B200_TRY(b200_scratch(ctx, 12, L.bytes, reinterpret_cast<void**>(&d_side)));
    B200_TRY(b200_copy_in(ctx, d_words, h_words, total_words * 4, ctx->stream));
    B200_TRY(b200_copy_in(ctx, d_side, h_side, L.bytes, ctx->stream));
    B200_TRY(b200_deflate_decompress_dev(ctx, d_words, total_words, d_side, L.bytes, n, block_size, d_tok, d_out));
    B200_TRY(b200_copy_out(ctx, h_out, d_out, n, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return B200_OK;
}
