// Host-buffer entry points: H2D -> *_dev -> D2H on the context's stream. These are
// what the reference-named shims (csrc/shims/*.c) and bench.py's end-to-end leg call.
// Device staging buffers live in the context's scratch slots 8..13 and are reused.
#include "common.cuh"
#include "../../include/b200comp.h"

namespace {
inline uint64_t lz_bs(uint64_t n, uint64_t block_size) { return (block_size == 0 || block_size > n) ? n : block_size; }
inline uint64_t lz_cap(int variant, uint64_t n, uint64_t nblocks) {
    return (variant == B200_LZ_DEFLATE ? 2 * n : n + n / 8 + 8 * nblocks) + 64;
}
}  // namespace

extern "C" uint64_t b200_lz77_max_bytes(int variant, uint64_t n, uint64_t block_size) {
    const uint64_t bs = lz_bs(n, block_size);
    const uint64_t nblocks = n ? (n + bs - 1) / bs : 1;
    return lz_cap(variant, n, nblocks);
}

extern "C" int b200_lz77_compress_host(b200_ctx* ctx, int variant, const uint8_t* h_in, uint64_t n, uint64_t block_size,
                                       uint8_t* h_out, uint64_t out_capacity, uint64_t* h_block_sizes,
                                       uint64_t* h_block_off, uint64_t* h_total_bytes) {
    if (n == 0) { if (h_block_off) h_block_off[0] = 0; if (h_total_bytes) *h_total_bytes = 0; return B200_OK; }
    const uint64_t bs = lz_bs(n, block_size);
    const uint64_t nblocks = (n + bs - 1) / bs;
    const uint64_t cap = lz_cap(variant, n, nblocks);
    uint8_t *d_in, *d_out; uint64_t* d_idx;
    B200_TRY(b200_scratch(ctx, 8, n + 64, reinterpret_cast<void**>(&d_in)));
    B200_TRY(b200_scratch(ctx, 9, cap, reinterpret_cast<void**>(&d_out)));
    B200_TRY(b200_scratch(ctx, 10, (2 * nblocks + 2) * 8, reinterpret_cast<void**>(&d_idx)));
    CUDA_TRY(cudaMemcpyAsync(d_in, h_in, n, cudaMemcpyHostToDevice, ctx->stream));
    uint64_t total = 0;
    B200_TRY(b200_lz77_encode_dev(ctx, variant, d_in, n, block_size, d_out, cap, d_idx, d_idx + nblocks, &total));
    if (total > out_capacity) { B200_SET_ERR("lz77: output needs %llu bytes, buffer has %llu", (unsigned long long)total, (unsigned long long)out_capacity); return B200_ERR_CAPACITY; }
    CUDA_TRY(cudaMemcpyAsync(h_out, d_out, total, cudaMemcpyDeviceToHost, ctx->stream));
    if (h_block_sizes) CUDA_TRY(cudaMemcpyAsync(h_block_sizes, d_idx, nblocks * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (h_block_off) CUDA_TRY(cudaMemcpyAsync(h_block_off, d_idx + nblocks, (nblocks + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (h_total_bytes) *h_total_bytes = total;
    return B200_OK;
}

extern "C" int b200_lz77_decompress_host(b200_ctx* ctx, int variant, const uint8_t* h_stream, uint64_t stream_bytes,
                                         const uint64_t* h_block_off, const uint64_t* h_block_sizes,
                                         uint64_t n, uint64_t block_size, uint8_t* h_out) {
    if (n == 0) return B200_OK;
    const uint64_t bs = lz_bs(n, block_size);
    const uint64_t nblocks = (n + bs - 1) / bs;
    uint8_t *d_stream, *d_out; uint64_t* d_idx;
    B200_TRY(b200_scratch(ctx, 9, stream_bytes + 64, reinterpret_cast<void**>(&d_stream)));
    B200_TRY(b200_scratch(ctx, 8, n + 64, reinterpret_cast<void**>(&d_out)));
    B200_TRY(b200_scratch(ctx, 10, (2 * nblocks + 2) * 8, reinterpret_cast<void**>(&d_idx)));
    CUDA_TRY(cudaMemcpyAsync(d_stream, h_stream, stream_bytes, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(d_idx, h_block_sizes, nblocks * 8, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(d_idx + nblocks, h_block_off, (nblocks + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    B200_TRY(b200_lz77_decode_dev(ctx, variant, d_stream, d_idx + nblocks, d_idx, n, block_size, d_out));
    CUDA_TRY(cudaMemcpyAsync(h_out, d_out, n, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return B200_OK;
}

extern "C" int b200_huffman_compress_host(b200_ctx* ctx, const uint8_t* h_in, uint64_t n, uint64_t block_size,
                                          uint32_t* h_words, uint64_t words_capacity, uint8_t* h_side, uint64_t side_bytes,
                                          uint64_t* h_total_words, uint32_t* h_worst_status) {
    b200_huff_layout L;
    B200_TRY(b200_huffman_layout(n, block_size, &L));
    if (h_side && side_bytes < L.bytes) { B200_SET_ERR("huffman: host side buffer too small"); return B200_ERR_CAPACITY; }
    const uint64_t cap = b200_huffman_max_words(n, block_size);
    uint8_t *d_in, *d_side; uint32_t* d_words;
    B200_TRY(b200_scratch(ctx, 8, n + 64, reinterpret_cast<void**>(&d_in)));
    B200_TRY(b200_scratch(ctx, 11, cap * 4, reinterpret_cast<void**>(&d_words)));
    B200_TRY(b200_scratch(ctx, 12, L.bytes, reinterpret_cast<void**>(&d_side)));
    CUDA_TRY(cudaMemcpyAsync(d_in, h_in, n, cudaMemcpyHostToDevice, ctx->stream));
    uint64_t total = 0; uint32_t worst = 0;
    B200_TRY(b200_huffman_encode_dev(ctx, d_in, n, block_size, d_words, cap, d_side, L.bytes, &total, &worst));
    if (h_total_words) *h_total_words = total;
    if (h_worst_status) *h_worst_status = worst;
    if (worst == 0) {
        if (total > words_capacity) { B200_SET_ERR("huffman: output needs %llu words, buffer has %llu", (unsigned long long)total, (unsigned long long)words_capacity); return B200_ERR_CAPACITY; }
        CUDA_TRY(cudaMemcpyAsync(h_words, d_words, total * 4, cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (h_side) CUDA_TRY(cudaMemcpyAsync(h_side, d_side, L.bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (worst) { B200_SET_ERR("huffman: input has a block the reference cannot encode (status %u)", worst); return B200_ERR_DOMAIN; }
    return B200_OK;
}

extern "C" int b200_huffman_decompress_host(b200_ctx* ctx, const uint32_t* h_words, uint64_t total_words,
                                            const uint8_t* h_side, uint64_t side_bytes, uint64_t n, uint64_t block_size,
                                            uint8_t* h_out) {
    if (n == 0) return B200_OK;
    b200_huff_layout L;
    B200_TRY(b200_huffman_layout(n, block_size, &L));
    if (side_bytes < L.bytes) { B200_SET_ERR("huffman: host side buffer too small"); return B200_ERR_CAPACITY; }
    uint8_t *d_out, *d_side; uint32_t* d_words;
    B200_TRY(b200_scratch(ctx, 8, n + 64, reinterpret_cast<void**>(&d_out)));
    B200_TRY(b200_scratch(ctx, 11, (total_words + 4) * 4, reinterpret_cast<void**>(&d_words)));
    B200_TRY(b200_scratch(ctx, 12, L.bytes, reinterpret_cast<void**>(&d_side)));
    CUDA_TRY(cudaMemcpyAsync(d_words, h_words, total_words * 4, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(d_side, h_side, L.bytes, cudaMemcpyHostToDevice, ctx->stream));
    B200_TRY(b200_huffman_decode_dev(ctx, d_words, total_words, d_side, L.bytes, n, block_size, d_out));
    CUDA_TRY(cudaMemcpyAsync(h_out, d_out, n, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return B200_OK;
}

// One-stream decode of a foreign Huffman stream from host buffers (index-free).
extern "C" int b200_huffman_decompress_serial_host(b200_ctx* ctx, const uint32_t* h_words, uint64_t nwords,
                                                   uint64_t buffer_size, const uint32_t* h_codes, const uint8_t* h_lens,
                                                   uint8_t* h_out, uint64_t out_capacity, uint64_t* h_count) {
    uint8_t *d_out, *d_tab; uint32_t* d_words;
    B200_TRY(b200_scratch(ctx, 8, out_capacity + 64, reinterpret_cast<void**>(&d_out)));
    B200_TRY(b200_scratch(ctx, 11, (nwords + 4) * 4, reinterpret_cast<void**>(&d_words)));
    B200_TRY(b200_scratch(ctx, 12, 2048, reinterpret_cast<void**>(&d_tab)));
    CUDA_TRY(cudaMemcpyAsync(d_words, h_words, nwords * 4, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(d_tab, h_codes, 1024, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(d_tab + 1024, h_lens, 256, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaMemsetAsync(d_out, 0, out_capacity, ctx->stream));
    uint64_t cnt = 0;
    B200_TRY(b200_huffman_decode_serial_dev(ctx, d_words, nwords, buffer_size, reinterpret_cast<uint32_t*>(d_tab), d_tab + 1024,
                                            d_out, out_capacity, &cnt));
    CUDA_TRY(cudaMemcpyAsync(h_out, d_out, cnt < out_capacity ? cnt : out_capacity, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (h_count) *h_count = cnt;
    return B200_OK;
}
