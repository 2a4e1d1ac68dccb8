// Self-describing streams for the Huffman codec and for entropy-coded deflate: ONE buffer
// holds everything a decoder needs -- header (magic, version, codec, sizes), the serialized code tables and
// the parallel-decode index, the payload -- so a stream written in one process is decoded in another without
// any side buffer or process state. The reference's only container model is the Zig Huffman's per-chunk tree
// dump + size header (algorithms/huffman/zig_huffman/src/main.zig:11-18,155-200,513-530); the C side
// concatenates raw (algorithms/deflate/deflate.c:56). The payload words are exactly the words of the
// non-container entry points (bit-exact with huffman_compress / the deflate token entropy stage).
//
// Layout (little endian, every section padded to 8 bytes):
//   u64 header[8] = { magic "B200CONT", version | codec << 32, n, block_size, nblocks, nchunks_stored, stream_words, 0 }
//   codec 1 (Huffman, algorithms/huffman):
//     u32 freq[nblocks][256]          histogram of every table scope: the "serialized code table" -- the decoder
//                                     replays the reference's heap on it (huffman.c:189-250) and gets the same
//                                     tree, codes and lengths (unlimited-length codes are not canonical, so
//                                     lengths alone would not do)
//     u32 chunk_bits[nchunks]         bits of every 4096-symbol chunk  (-> per-block {bits, first word} by summing)
//     u32 sub_off[nchunks][16]        bit offset of every 256-symbol sub-chunk inside its chunk (decode index)
//     u32 words[stream_words]         MSB-first u32 words, every block starting on a word
//   codec 2 (deflate = lz77_compress per block + Huffman-coded tokens, algorithms/deflate):
//     u32 freq[nblocks][288]          frequencies[286] (deflate/lz77.c:206) padded to 288
//     u64 tok_sizes[nblocks]          token bytes of every block (the LZ77 decoder's block sizes)
//     u32 chunk_bits[nchunks_stored]  only the chunk slots a block really uses: ceil(tok_size / 4096) per block
//     u32 sub_off[nchunks_stored][16] bit 31 = the sub-chunk starts with the tail unit of a match
//     u32 words[stream_words]
#include <cstdlib>
#include <vector>
#include "common.cuh"
#include "../../include/b200comp.h"

extern "C" int b200_huffman_tables_from_freq_dev(b200_ctx* ctx, uint8_t* d_side, uint64_t side_bytes, uint64_t n, uint64_t block_size);
extern "C" int b200_dfl_tables_from_freq_dev(b200_ctx* ctx, uint8_t* d_side, uint64_t side_bytes, uint64_t n, uint64_t block_size);

namespace {
constexpr uint64_t CONT_MAGIC = 0x544E4F4330303242ull;   // "B200CONT"
constexpr uint32_t CONT_VERSION = 1;
constexpr uint64_t CODEC_HUFFMAN = 1, CODEC_DEFLATE = 2;
inline uint64_t a8(uint64_t x) { return (x + 7) & ~(uint64_t)7; }
inline uint64_t lz_tok_cap(uint64_t n, uint64_t nblocks) { return 2 * n + 2 * nblocks + 64; }

struct Sections { uint64_t o_freq, o_tok, o_cbits, o_sub, o_words, total; };
Sections sections(uint64_t codec, uint64_t nblocks, uint64_t nchunks, uint64_t words) {
    Sections s;
    uint64_t o = 64;
    s.o_freq = o; o += a8(nblocks * (codec == CODEC_HUFFMAN ? 256 : 288) * 4);
    s.o_tok = o;  if (codec == CODEC_DEFLATE) o += nblocks * 8;
    s.o_cbits = o; o += a8(nchunks * 4);
    s.o_sub = o;   o += a8(nchunks * 64);
    s.o_words = o; o += a8(words * 4);
    s.total = o;
    return s;
}
}  // namespace

extern "C" uint64_t b200_huffman_container_max_bytes(uint64_t n, uint64_t block_size) {
    b200_huff_layout L;
    if (b200_huffman_layout(n ? n : 1, block_size, &L) != B200_OK) return 0;
    return sections(CODEC_HUFFMAN, L.nblocks, L.nchunks, b200_huffman_max_words(n, block_size)).total;
}

extern "C" uint64_t b200_deflate_container_max_bytes(uint64_t n, uint64_t block_size) {
    b200_dfl_layout L;
    if (b200_dfl_layout_for(n ? n : 1, block_size, &L) != B200_OK) return 0;
    return sections(CODEC_DEFLATE, L.nblocks, L.nchunks, b200_dfl_max_words(n, block_size)).total;
}

extern "C" int b200_container_info(const void* h_container, uint64_t bytes, uint32_t* codec, uint64_t* n, uint64_t* block_size) {
    const uint64_t* h = static_cast<const uint64_t*>(h_container);
    if (!h || bytes < 64 || h[0] != CONT_MAGIC) { B200_SET_ERR("container: not a B200CONT stream"); return B200_ERR_FORMAT; }
    if ((uint32_t)h[1] != CONT_VERSION) { B200_SET_ERR("container: version %u is not supported (this library reads version %u)", (uint32_t)h[1], CONT_VERSION); return B200_ERR_FORMAT; }
    const uint64_t cd = h[1] >> 32;
    if (cd != CODEC_HUFFMAN && cd != CODEC_DEFLATE) { B200_SET_ERR("container: unknown codec %llu", (unsigned long long)cd); return B200_ERR_FORMAT; }
    if (codec) *codec = (uint32_t)cd;
    if (n) *n = h[2];
    if (block_size) *block_size = h[3];
    return B200_OK;
}

// ------------------------------------------------------------------------------------------ Huffman
extern "C" int b200_huffman_compress_container_host(b200_ctx* ctx, const uint8_t* h_in, uint64_t n, uint64_t block_size,
                                                    void* h_out, uint64_t out_capacity, uint64_t* h_total_bytes) {
    B200_ENTER(ctx);
    b200_huff_layout L;
    B200_TRY(b200_huffman_layout(n, block_size, &L));
    const uint64_t cap = b200_huffman_max_words(n, block_size);
    uint8_t *d_in, *d_side; uint32_t* d_words;
    B200_TRY(b200_scratch(ctx, 8, n + 64, reinterpret_cast<void**>(&d_in)));
    B200_TRY(b200_scratch(ctx, 11, cap * 4, reinterpret_cast<void**>(&d_words)));
    B200_TRY(b200_scratch(ctx, 12, L.bytes, reinterpret_cast<void**>(&d_side)));
    B200_TRY(b200_copy_in(ctx, d_in, h_in, n, ctx->stream));
    uint64_t total = 0; uint32_t worst = 0;
    B200_TRY(b200_huffman_encode_dev(ctx, d_in, n, block_size, d_words, cap, d_side, L.bytes, &total, &worst));
    if (worst) { B200_SET_ERR("huffman: input has a block the reference cannot encode (status %u)", worst); return B200_ERR_DOMAIN; }
    const Sections s = sections(CODEC_HUFFMAN, L.nblocks, L.nchunks, total);
    if (h_total_bytes) *h_total_bytes = s.total;
    if (s.total > out_capacity) { B200_SET_ERR("huffman container needs %llu bytes, buffer has %llu", (unsigned long long)s.total, (unsigned long long)out_capacity); return B200_ERR_CAPACITY; }
    uint8_t* o = static_cast<uint8_t*>(h_out);
    const uint64_t bs_eff = (block_size == 0 || block_size >= n) ? 0 : block_size;
    const uint64_t hdr[8] = {CONT_MAGIC, CONT_VERSION | (CODEC_HUFFMAN << 32), n, bs_eff, L.nblocks, L.nchunks, total, 0};
    memcpy(o, hdr, 64);
    memset(o + s.o_sub - 8, 0, 8); memset(o + s.o_words - 8, 0, 8); memset(o + s.total - 8, 0, 8);   // section padding
    B200_TRY(b200_copy_out(ctx, o + s.o_freq, d_side + L.off_freq, L.nblocks * 1024, ctx->stream));
    B200_TRY(b200_copy_out(ctx, o + s.o_cbits, d_side + L.off_chunk_bits, L.nchunks * 4, ctx->stream));
    B200_TRY(b200_copy_out(ctx, o + s.o_sub, d_side + L.off_sub_off, L.nchunks * 64, ctx->stream));
    B200_TRY(b200_copy_out(ctx, o + s.o_words, d_words, total * 4, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return B200_OK;
}

extern "C" int b200_huffman_decompress_container_host(b200_ctx* ctx, const void* h_container, uint64_t bytes,
                                                      uint8_t* h_out, uint64_t out_capacity, uint64_t* h_n) {
    B200_ENTER(ctx);
    uint32_t codec; uint64_t n, bs;
    B200_TRY(b200_container_info(h_container, bytes, &codec, &n, &bs));
    if (codec != CODEC_HUFFMAN) { B200_SET_ERR("container: codec %u is not Huffman", codec); return B200_ERR_FORMAT; }
    const uint64_t* h = static_cast<const uint64_t*>(h_container);
    const uint8_t* c = static_cast<const uint8_t*>(h_container);
    b200_huff_layout L;
    if (n == 0 || b200_huffman_layout(n, bs, &L) != B200_OK || L.nblocks != h[4] || L.nchunks != h[5]) { B200_SET_ERR("huffman container: inconsistent header"); return B200_ERR_FORMAT; }
    const uint64_t total = h[6];
    const Sections s = sections(CODEC_HUFFMAN, L.nblocks, L.nchunks, total);
    if (s.total > bytes) { B200_SET_ERR("huffman container: truncated (%llu of %llu bytes)", (unsigned long long)bytes, (unsigned long long)s.total); return B200_ERR_FORMAT; }
    if (h_n) *h_n = n;
    if (n > out_capacity) { B200_SET_ERR("huffman container: output needs %llu bytes", (unsigned long long)n); return B200_ERR_CAPACITY; }
    // rebuild the index the decoder reads: per-block bits and first word, absolute bit offset of every chunk
    std::vector<uint8_t> side(L.bytes, 0);
    const uint32_t* cbits = reinterpret_cast<const uint32_t*>(c + s.o_cbits);
    const uint32_t* sub = reinterpret_cast<const uint32_t*>(c + s.o_sub);
    uint64_t* block_bits = reinterpret_cast<uint64_t*>(side.data() + L.off_block_bits);
    uint64_t* block_word = reinterpret_cast<uint64_t*>(side.data() + L.off_block_word);
    uint64_t* chunk_off = reinterpret_cast<uint64_t*>(side.data() + L.off_chunk_off);
    uint64_t word = 0;
    for (uint64_t b = 0; b < L.nblocks; ++b) {
        uint64_t bits = 0;
        block_word[b] = word;
        const uint64_t c0 = b * L.chunks_per_block, c1 = c0 + L.chunks_per_block < L.nchunks ? c0 + L.chunks_per_block : L.nchunks;
        for (uint64_t k = c0; k < c1; ++k) {
            if (cbits[k] > 4096u * 32u) { B200_SET_ERR("huffman container: chunk %llu claims %u bits", (unsigned long long)k, cbits[k]); return B200_ERR_FORMAT; }
            for (int j = 0; j < 16; ++j) if (sub[k * 16 + j] > cbits[k]) { B200_SET_ERR("huffman container: corrupt decode index in chunk %llu", (unsigned long long)k); return B200_ERR_FORMAT; }
            chunk_off[k] = word * 32 + bits;
            bits += cbits[k];
        }
        block_bits[b] = bits;
        word += (bits + 31) >> 5;
    }
    block_word[L.nblocks] = word;
    if (word != total) { B200_SET_ERR("huffman container: index covers %llu words, the stream has %llu", (unsigned long long)word, (unsigned long long)total); return B200_ERR_FORMAT; }
    memcpy(side.data() + L.off_freq, c + s.o_freq, L.nblocks * 1024);
    memcpy(side.data() + L.off_chunk_bits, cbits, L.nchunks * 4);
    memcpy(side.data() + L.off_sub_off, sub, L.nchunks * 64);
    uint8_t *d_out, *d_side; uint32_t* d_words;
    B200_TRY(b200_scratch(ctx, 8, n + 64, reinterpret_cast<void**>(&d_out)));
    B200_TRY(b200_scratch(ctx, 11, (total + 4) * 4, reinterpret_cast<void**>(&d_words)));
    B200_TRY(b200_scratch(ctx, 12, L.bytes, reinterpret_cast<void**>(&d_side)));
    CUDA_TRY(cudaMemcpyAsync(d_side, side.data(), L.bytes, cudaMemcpyHostToDevice, ctx->stream));
    B200_TRY(b200_copy_in(ctx, d_words, c + s.o_words, total * 4, ctx->stream));
    CUDA_TRY(cudaMemsetAsync(d_words + total, 0, 16, ctx->stream));
    B200_TRY(b200_huffman_tables_from_freq_dev(ctx, d_side, L.bytes, n, bs));
    {   // a histogram that cannot have produced this stream (fewer than two symbols, a code beyond 32 bits) is corrupt
        uint32_t* pin; B200_TRY(b200_pinned(ctx, 16 + L.nblocks * 16, reinterpret_cast<void**>(&pin)));
        CUDA_TRY(cudaMemcpyAsync(pin, d_side + L.off_meta, L.nblocks * 16, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));   // (also keeps `side` alive until the copy is done)
        for (uint64_t b = 0; b < L.nblocks; ++b) if (pin[b * 4]) { B200_SET_ERR("huffman container: the table of block %llu is not decodable (status %u)", (unsigned long long)b, pin[b * 4]); return B200_ERR_FORMAT; }
    }
    B200_TRY(b200_huffman_decode_dev(ctx, d_words, total, d_side, L.bytes, n, bs, d_out));
    B200_TRY(b200_copy_out(ctx, h_out, d_out, n, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return B200_OK;
}

// ------------------------------------------------------------------------------------------ deflate
extern "C" int b200_deflate_compress_container_host(b200_ctx* ctx, const uint8_t* h_in, uint64_t n, uint64_t block_size,
                                                    void* h_out, uint64_t out_capacity, uint64_t* h_total_bytes) {
    B200_ENTER(ctx);
    if (n == 0) { B200_SET_ERR("deflate container: empty input"); return B200_ERR_DOMAIN; }
    b200_dfl_layout L;
    B200_TRY(b200_dfl_layout_for(n, block_size, &L));
    const uint64_t cap = b200_dfl_max_words(n, block_size);
    const uint64_t tok_cap = lz_tok_cap(n, L.nblocks);
    uint8_t *d_in, *d_tok, *d_side; uint32_t* d_words; uint64_t* d_idx;
    B200_TRY(b200_scratch(ctx, 8, n + 64, reinterpret_cast<void**>(&d_in)));
    B200_TRY(b200_scratch(ctx, 9, tok_cap, reinterpret_cast<void**>(&d_tok)));
    B200_TRY(b200_scratch(ctx, 10, (2 * L.nblocks + 2) * 8, reinterpret_cast<void**>(&d_idx)));
    B200_TRY(b200_scratch(ctx, 11, cap * 4, reinterpret_cast<void**>(&d_words)));
    B200_TRY(b200_scratch(ctx, 12, L.bytes, reinterpret_cast<void**>(&d_side)));
    B200_TRY(b200_copy_in(ctx, d_in, h_in, n, ctx->stream));
    uint64_t total = 0; uint32_t worst = 0;
    B200_TRY(b200_deflate_compress_dev(ctx, d_in, n, block_size, d_tok, tok_cap, d_idx, d_idx + L.nblocks, d_words, cap, d_side, L.bytes, &total, &worst));
    if (worst) { B200_SET_ERR("deflate: a block needs a code longer than 32 bits (status %u)", worst); return B200_ERR_DOMAIN; }
    // the chunk arrays are strided by the worst case (33 slots per 64 KiB block): keep only the used slots
    std::vector<uint64_t> tok_sizes(L.nblocks);
    std::vector<uint32_t> cbits(L.nchunks), sub(L.nchunks * 16);
    CUDA_TRY(cudaMemcpyAsync(tok_sizes.data(), d_side + L.off_tok_sizes, L.nblocks * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(cbits.data(), d_side + L.off_chunk_bits, L.nchunks * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(sub.data(), d_side + L.off_sub_off, L.nchunks * 64, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    uint64_t stored = 0;
    for (uint64_t b = 0; b < L.nblocks; ++b) stored += (tok_sizes[b] + B200_DFL_CHUNK - 1) / B200_DFL_CHUNK;
    const Sections s = sections(CODEC_DEFLATE, L.nblocks, stored, total);
    if (h_total_bytes) *h_total_bytes = s.total;
    if (s.total > out_capacity) { B200_SET_ERR("deflate container needs %llu bytes, buffer has %llu", (unsigned long long)s.total, (unsigned long long)out_capacity); return B200_ERR_CAPACITY; }
    uint8_t* o = static_cast<uint8_t*>(h_out);
    const uint64_t bs_eff = (block_size == 0 || block_size >= n) ? 0 : block_size;
    const uint64_t hdr[8] = {CONT_MAGIC, CONT_VERSION | (CODEC_DEFLATE << 32), n, bs_eff, L.nblocks, stored, total, 0};
    memcpy(o, hdr, 64);
    memset(o + s.o_sub - 8, 0, 8); memset(o + s.o_words - 8, 0, 8); memset(o + s.total - 8, 0, 8);
    memcpy(o + s.o_tok, tok_sizes.data(), L.nblocks * 8);
    uint32_t* ocb = reinterpret_cast<uint32_t*>(o + s.o_cbits); uint32_t* osub = reinterpret_cast<uint32_t*>(o + s.o_sub);
    uint64_t k = 0;
    for (uint64_t b = 0; b < L.nblocks; ++b) {
        const uint64_t used = (tok_sizes[b] + B200_DFL_CHUNK - 1) / B200_DFL_CHUNK, c0 = b * L.chunks_per_block;
        memcpy(ocb + k, cbits.data() + c0, used * 4);
        memcpy(osub + k * 16, sub.data() + c0 * 16, used * 64);
        k += used;
    }
    B200_TRY(b200_copy_out(ctx, o + s.o_freq, d_side + L.off_freq, L.nblocks * 288 * 4, ctx->stream));
    B200_TRY(b200_copy_out(ctx, o + s.o_words, d_words, total * 4, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return B200_OK;
}

extern "C" int b200_deflate_decompress_container_host(b200_ctx* ctx, const void* h_container, uint64_t bytes,
                                                      uint8_t* h_out, uint64_t out_capacity, uint64_t* h_n) {
    B200_ENTER(ctx);
    uint32_t codec; uint64_t n, bs;
    B200_TRY(b200_container_info(h_container, bytes, &codec, &n, &bs));
    if (codec != CODEC_DEFLATE) { B200_SET_ERR("container: codec %u is not deflate", codec); return B200_ERR_FORMAT; }
    const uint64_t* h = static_cast<const uint64_t*>(h_container);
    const uint8_t* c = static_cast<const uint8_t*>(h_container);
    b200_dfl_layout L;
    if (n == 0 || b200_dfl_layout_for(n, bs, &L) != B200_OK || L.nblocks != h[4] || h[5] > L.nchunks) { B200_SET_ERR("deflate container: inconsistent header"); return B200_ERR_FORMAT; }
    const uint64_t stored = h[5], total = h[6];
    const Sections s = sections(CODEC_DEFLATE, L.nblocks, stored, total);
    if (s.total > bytes) { B200_SET_ERR("deflate container: truncated (%llu of %llu bytes)", (unsigned long long)bytes, (unsigned long long)s.total); return B200_ERR_FORMAT; }
    if (h_n) *h_n = n;
    if (n > out_capacity) { B200_SET_ERR("deflate container: output needs %llu bytes", (unsigned long long)n); return B200_ERR_CAPACITY; }
    std::vector<uint8_t> side(L.bytes, 0);
    const uint64_t* tok_sizes = reinterpret_cast<const uint64_t*>(c + s.o_tok);
    const uint32_t* cbits = reinterpret_cast<const uint32_t*>(c + s.o_cbits);
    const uint32_t* sub = reinterpret_cast<const uint32_t*>(c + s.o_sub);
    uint64_t* tok_off = reinterpret_cast<uint64_t*>(side.data() + L.off_tok_off);
    uint64_t* block_bits = reinterpret_cast<uint64_t*>(side.data() + L.off_block_bits);
    uint64_t* block_word = reinterpret_cast<uint64_t*>(side.data() + L.off_block_word);
    uint64_t* chunk_off = reinterpret_cast<uint64_t*>(side.data() + L.off_chunk_off);
    uint32_t* s_cbits = reinterpret_cast<uint32_t*>(side.data() + L.off_chunk_bits);
    uint32_t* s_sub = reinterpret_cast<uint32_t*>(side.data() + L.off_sub_off);
    const uint64_t bs_eff = (bs == 0 || bs > n) ? n : bs;
    uint64_t word = 0, k = 0, toff = 0;
    for (uint64_t b = 0; b < L.nblocks; ++b) {
        const uint64_t len = b * bs_eff + bs_eff <= n ? bs_eff : n - b * bs_eff;
        if (tok_sizes[b] > 2 * len + 2 || (tok_sizes[b] & 1)) { B200_SET_ERR("deflate container: block %llu claims %llu token bytes", (unsigned long long)b, (unsigned long long)tok_sizes[b]); return B200_ERR_FORMAT; }
        const uint64_t used = (tok_sizes[b] + B200_DFL_CHUNK - 1) / B200_DFL_CHUNK, c0 = b * L.chunks_per_block;
        if (k + used > stored) { B200_SET_ERR("deflate container: chunk index too short"); return B200_ERR_FORMAT; }
        tok_off[b] = toff; toff += tok_sizes[b];
        block_word[b] = word;
        uint64_t bits = 0;
        for (uint64_t j = 0; j < used; ++j) {
            if (cbits[k + j] > B200_DFL_CHUNK / 2 * 32u) { B200_SET_ERR("deflate container: chunk %llu claims %u bits", (unsigned long long)(k + j), cbits[k + j]); return B200_ERR_FORMAT; }
            for (int q = 0; q < 16; ++q) if ((sub[(k + j) * 16 + q] & 0x7FFFFFFFu) > cbits[k + j]) { B200_SET_ERR("deflate container: corrupt decode index in chunk %llu", (unsigned long long)(k + j)); return B200_ERR_FORMAT; }
            chunk_off[c0 + j] = word * 32 + bits;
            s_cbits[c0 + j] = cbits[k + j];
            memcpy(s_sub + (c0 + j) * 16, sub + (k + j) * 16, 64);
            bits += cbits[k + j];
        }
        for (uint64_t j = used; j < L.chunks_per_block; ++j) chunk_off[c0 + j] = word * 32 + bits;
        block_bits[b] = bits;
        word += (bits + 31) >> 5;
        k += used;
    }
    tok_off[L.nblocks] = toff;
    block_word[L.nblocks] = word;
    if (word != total || k != stored) { B200_SET_ERR("deflate container: index covers %llu words / %llu chunks, the stream has %llu / %llu", (unsigned long long)word, (unsigned long long)k, (unsigned long long)total, (unsigned long long)stored); return B200_ERR_FORMAT; }
    memcpy(side.data() + L.off_freq, c + s.o_freq, L.nblocks * 288 * 4);
    memcpy(side.data() + L.off_tok_sizes, tok_sizes, L.nblocks * 8);
    const uint64_t tok_cap = lz_tok_cap(n, L.nblocks);
    uint8_t *d_out, *d_tok, *d_side; uint32_t* d_words;
    B200_TRY(b200_scratch(ctx, 8, n + 64, reinterpret_cast<void**>(&d_out)));
    B200_TRY(b200_scratch(ctx, 9, tok_cap, reinterpret_cast<void**>(&d_tok)));
    B200_TRY(b200_scratch(ctx, 11, (total + 4) * 4, reinterpret_cast<void**>(&d_words)));
    B200_TRY(b200_scratch(ctx, 12, L.bytes, reinterpret_cast<void**>(&d_side)));
    CUDA_TRY(cudaMemcpyAsync(d_side, side.data(), L.bytes, cudaMemcpyHostToDevice, ctx->stream));
    B200_TRY(b200_copy_in(ctx, d_words, c + s.o_words, total * 4, ctx->stream));
    CUDA_TRY(cudaMemsetAsync(d_words + total, 0, 16, ctx->stream));
    B200_TRY(b200_dfl_tables_from_freq_dev(ctx, d_side, L.bytes, n, bs));
    {
        uint32_t* pin; B200_TRY(b200_pinned(ctx, 16 + L.nblocks * 16, reinterpret_cast<void**>(&pin)));
        CUDA_TRY(cudaMemcpyAsync(pin, d_side + L.off_meta, L.nblocks * 16, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        for (uint64_t b = 0; b < L.nblocks; ++b) if (pin[b * 4] && tok_sizes[b]) { B200_SET_ERR("deflate container: the table of block %llu is not decodable (status %u)", (unsigned long long)b, pin[b * 4]); return B200_ERR_FORMAT; }
    }
    B200_TRY(b200_deflate_decompress_dev(ctx, d_words, total, d_side, L.bytes, n, bs, d_tok, d_out));
    B200_TRY(b200_copy_out(ctx, h_out, d_out, n, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return B200_OK;
}
