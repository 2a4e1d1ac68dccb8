// Huffman: histogram, exact heap-order table build, scan-based MSB-first bit packing
// and table-lookup decoding for sm_100a.
//
// Reference being replaced (bit-exact): /root/reference/algorithms/huffman/huffman.c
//   histogram :184-187, heap :100-159, tree :189-211, codes :217-250,
//   write_bits :18-48, sizes :318-320, decoder :330-364.
//
// Data layout in HBM: input bytes; output = host-endian u32 words, MSB-first, each
// table scope ("block") starting on a word boundary; a side buffer (b200_huff_layout)
// with per-block tables and the chunk/sub-chunk bit index used for parallel decode.
#include <algorithm>
#include <vector>
#include "common.cuh"
#include "hist.cuh"
#include "huff_shared.cuh"
#include "../../include/b200comp.h"

namespace {

constexpr uint32_t CHUNK = B200_HUFF_CHUNK;  // symbols per encode CTA
constexpr uint32_t SUB = B200_HUFF_SUB;      // symbols per decode thread
constexpr uint32_t SUBS_PER_CHUNK = CHUNK / SUB;
constexpr uint32_t TILE_CHUNKS = 16;         // chunks per histogram / decode-lut tile
constexpr uint32_t LUT_BITS = 12;

inline uint64_t eff_block(uint64_t n, uint64_t bs) {
    if (bs == 0 || bs >= n) return ((n ? n : 1) + CHUNK - 1) / CHUNK * CHUNK;
    return bs;
}

// K1 histogram: byte_hist_kernel in hist.cuh (shared with FSE)

// K2 table build, K4 offsets: huff_shared.cuh (shared with the deflate token entropy stage)

// ---------------------------------------------------------------- K3 chunk bit counts
__global__ void __launch_bounds__(256) huff_chunkbits_kernel(const uint8_t* __restrict__ in, uint64_t n, uint32_t cpb,
                                                             const uint8_t* __restrict__ lens, uint32_t* __restrict__ chunk_bits) {
    __shared__ uint8_t sl[256];
    __shared__ uint32_t wsum[8];
    const uint64_t c = blockIdx.x, b = c / cpb;
    sl[threadIdx.x] = lens[b * 256 + threadIdx.x];
    __syncthreads();
    const uint64_t i = c * CHUNK + (uint64_t)threadIdx.x * 16;
    uint32_t bits = 0;
    if (i + 16 <= n) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(in + i));
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) bits += sl[w[q] & 0xFF] + sl[(w[q] >> 8) & 0xFF] + sl[(w[q] >> 16) & 0xFF] + sl[w[q] >> 24];
    } else {
        for (uint64_t j = i; j < n && j < i + 16; ++j) bits += sl[in[j]];
    }
    bits = warp_sum_u32(bits);
    if (lane_id() == 0) wsum[threadIdx.x >> 5] = bits;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += wsum[w];
        chunk_bits[c] = t;
    }
}

// ---------------------------------------------------------------- K5 encode
// One CTA per chunk of 4096 symbols; thread t owns 16 consecutive symbols. A block
// exclusive scan of the code lengths gives each thread its bit offset; codes are
// packed MSB-first into a shared-memory image of the output words (OR only on the
// words a thread shares with its neighbours), then stored coalesced.
__global__ void __launch_bounds__(256) huff_encode_kernel(const uint8_t* __restrict__ in, uint64_t n, uint32_t cpb,
                                                          const uint32_t* __restrict__ codes, const uint8_t* __restrict__ lens,
                                                          const uint32_t* __restrict__ meta, const uint32_t* __restrict__ chunk_bits,
                                                          const uint64_t* __restrict__ chunk_off, uint32_t* __restrict__ sub_off,
                                                          uint32_t* __restrict__ words, const uint64_t* __restrict__ info) {
    __shared__ uint32_t sc[256];
    __shared__ uint8_t  sl[256];
    __shared__ uint32_t stage[CHUNK + 2];
    __shared__ uint32_t wtot[8];
    if (info[1]) return;  // output does not fit
    const uint64_t c = blockIdx.x, b = c / cpb;
    if (meta[b * 4 + 0]) return;  // block the reference would not encode
    sc[threadIdx.x] = codes[b * 256 + threadIdx.x];
    sl[threadIdx.x] = lens[b * 256 + threadIdx.x];
    const uint64_t A = chunk_off[c];
    const uint32_t T = chunk_bits[c];
    const uint32_t r = (uint32_t)(A & 31);
    const uint32_t nw = (r + T + 31) >> 5;
    for (uint32_t j = threadIdx.x; j < nw; j += 256) stage[j] = 0;
    __syncthreads();

    const uint64_t i0 = c * CHUNK + (uint64_t)threadIdx.x * 16;
    uint8_t sym[16];
    uint32_t cnt = 0;
    if (i0 + 16 <= n) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(in + i0));
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            sym[4 * q + 0] = w[q] & 0xFF; sym[4 * q + 1] = (w[q] >> 8) & 0xFF;
            sym[4 * q + 2] = (w[q] >> 16) & 0xFF; sym[4 * q + 3] = w[q] >> 24;
        }
        cnt = 16;
    } else if (i0 < n) {
        cnt = (uint32_t)(n - i0);
#pragma unroll
        for (int q = 0; q < 16; ++q) sym[q] = (uint32_t)q < cnt ? in[i0 + q] : 0;
    }
    uint32_t mybits = 0;
#pragma unroll
    for (int q = 0; q < 16; ++q) if ((uint32_t)q < cnt) mybits += sl[sym[q]];
    // block exclusive scan of mybits
    const uint32_t incl = warp_incl_scan_u32(mybits);
    if (lane_id() == 31) wtot[threadIdx.x >> 5] = incl;
    __syncthreads();
    uint32_t wbase = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) if ((uint32_t)w < (threadIdx.x >> 5)) wbase += wtot[w];
    const uint32_t ex = wbase + incl - mybits;
    if ((threadIdx.x & 15) == 0) sub_off[c * SUBS_PER_CHUNK + (threadIdx.x >> 4)] = ex;

    if (mybits) {
        uint32_t pos = r + ex;
        uint32_t wi = pos >> 5;
        uint32_t fill = pos & 31;       // bits of word wi that belong to earlier threads
        uint64_t acc = 0;               // left-aligned: top `have` bits are meaningful
        uint32_t have = fill;
        bool first = fill != 0;
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            if ((uint32_t)q < cnt) {
                const uint32_t L = sl[sym[q]];
                const uint64_t code = sc[sym[q]];
                if (L) acc |= code << (64 - have - L);
                have += L;
                if (have >= 32) {
                    const uint32_t wv = (uint32_t)(acc >> 32);
                    if (first) { atomicOr(&stage[wi], wv); first = false; }
                    else stage[wi] = wv;
                    ++wi; acc <<= 32; have -= 32;
                }
            }
        }
        if (have) atomicOr(&stage[wi], (uint32_t)(acc >> 32));
    }
    __syncthreads();
    uint32_t* dst = words + (A >> 5);
    // the last word is shared only if another chunk of the same block follows
    const bool tail_shared = ((r + T) & 31) != 0 && (c + 1) % cpb != 0 && c + 1 < gridDim.x;
    for (uint32_t j = threadIdx.x; j < nw; j += 256) {
        const uint32_t v = stage[j];
        if ((j == 0 && r != 0) || (j == nw - 1 && tail_shared)) atomicOr(&dst[j], v);
        else dst[j] = v;
    }
}

// ---------------------------------------------------------------- K6 decode
// One CTA per tile of 16 chunks of one block (256 threads, one 256-symbol sub-chunk
// each). A 12-bit primary table in shared memory resolves codes up to 12 bits in one
// lookup: entry = 0x8000 | len << 8 | symbol. A longer (rare) code finds in the table the
// tree node reached after its first 12 bits and walks the reference's own tree from there,
// one bit per step (children in shared memory), so its cost is len - 12 steps and not a
// search over all long codes.
__global__ void __launch_bounds__(256) huff_decode_kernel(const uint32_t* __restrict__ words, uint64_t total_words,
                                                          uint64_t n, uint64_t bs, uint32_t cpb, uint32_t tiles_per_block,
                                                          const int16_t* __restrict__ tree, const uint32_t* __restrict__ meta,
                                                          const uint64_t* __restrict__ chunk_off, const uint32_t* __restrict__ sub_off,
                                                          uint8_t* __restrict__ out, const uint16_t* __restrict__ lut_g, int build_only) {
    __shared__ __align__(16) uint16_t lut[1u << LUT_BITS];
    __shared__ int16_t  kids[511][2];          // {left, right}; leaf = {-1, symbol}
    const uint64_t b = build_only ? blockIdx.x : blockIdx.x / tiles_per_block, k = build_only ? 0 : blockIdx.x % tiles_per_block;
    {
        const uint32_t* t32 = reinterpret_cast<const uint32_t*>(tree + b * 511 * 2);
        uint32_t* k32 = reinterpret_cast<uint32_t*>(&kids[0][0]);
        for (uint32_t i = threadIdx.x; i < 511; i += 256) k32[i] = t32[i];
    }
    const uint32_t root = meta[b * 4 + 2];
    __syncthreads();
    if (lut_g && !build_only) {        // the table of this block was built once (several CTAs share a table scope): copy it
        const uint4* src = reinterpret_cast<const uint4*>(lut_g + b * (1u << LUT_BITS));
        uint4* dst = reinterpret_cast<uint4*>(lut);
        for (uint32_t i = threadIdx.x; i < (1u << LUT_BITS) / 8; i += 256) dst[i] = __ldg(src + i);
    } else {
        for (uint32_t i = threadIdx.x; i < (1u << LUT_BITS); i += 256) {
            uint32_t v = root, d = 0;
            while (d < LUT_BITS && kids[v][0] >= 0) { v = (uint32_t)kids[v][(i >> (LUT_BITS - 1 - d)) & 1u]; ++d; }
            lut[i] = kids[v][0] < 0 ? (uint16_t)(0x8000u | (d << 8) | (uint32_t)(uint16_t)kids[v][1]) : (uint16_t)v;   // leaf | inner node after 12 bits
        }
    }
    __syncthreads();
    if (build_only) {
        uint16_t* dstg = const_cast<uint16_t*>(lut_g) + b * (1u << LUT_BITS);
        for (uint32_t i = threadIdx.x; i < (1u << LUT_BITS); i += 256) dstg[i] = lut[i];
        return;
    }

    const uint64_t chunk = b * cpb + (uint64_t)k * TILE_CHUNKS + (threadIdx.x >> 4);
    const uint64_t sub = chunk * SUBS_PER_CHUNK + (threadIdx.x & 15);
    const uint64_t sym0 = sub * SUB;
    uint64_t blk_end = (b + 1) * bs; if (blk_end > n) blk_end = n;
    if ((threadIdx.x >> 4) + k * TILE_CHUNKS >= cpb || sym0 >= blk_end) return;
    const uint32_t count = (uint32_t)(blk_end - sym0 < SUB ? blk_end - sym0 : SUB);
    const uint64_t bitpos = chunk_off[chunk] + sub_off[sub];
    uint64_t wi = bitpos >> 5;
    // window: top `avail` bits of win are the next stream bits
    uint64_t win = 0; uint32_t avail = 0;
    {
        const uint32_t sh = (uint32_t)(bitpos & 31);
        const uint64_t w0 = wi < total_words ? __ldg(&words[wi]) : 0; ++wi;
        win = w0 << (32 + sh); avail = 32 - sh;
        const uint64_t w1 = wi < total_words ? __ldg(&words[wi]) : 0; ++wi;
        win |= w1 << (32 - avail); avail += 32;
    }
    uint32_t nextw = wi < total_words ? __ldg(&words[wi]) : 0;   // one word ahead: the refill never waits for memory
    uint8_t* o = out + sym0;
    uint32_t done = 0;
    while (done < count) {
        uint32_t pack[4] = {0, 0, 0, 0};
        const uint32_t batch = count - done < 16 ? count - done : 16;
#pragma unroll
        for (uint32_t q = 0; q < 16; ++q) {
            if (q < batch) {
                if (avail <= 32) {
                    win |= (uint64_t)nextw << (32 - avail); avail += 32;
                    ++wi;
                    nextw = wi < total_words ? __ldg(&words[wi]) : 0;
                }
                const uint32_t e = lut[(uint32_t)(win >> (64 - LUT_BITS))];
                uint32_t s, L;
                if (e & 0x8000u) { s = e & 0xFF; L = (e >> 8) & 0x7F; }
                else {
                    uint32_t v = e; L = LUT_BITS;          // codes are at most 32 bits (status 2 otherwise), the window holds > 32
                    while (kids[v][0] >= 0 && L < 40) { v = (uint32_t)kids[v][(uint32_t)(win >> (63 - L)) & 1u]; ++L; }
                    s = (uint32_t)(uint16_t)kids[v][1] & 0xFF;
                }
                win <<= L; avail -= L;
                pack[q >> 2] |= s << (8 * (q & 3));
            }
        }
        if (batch == 16 && ((reinterpret_cast<uintptr_t>(o + done) & 15) == 0)) {
            *reinterpret_cast<uint4*>(o + done) = make_uint4(pack[0], pack[1], pack[2], pack[3]);
        } else {
            for (uint32_t q = 0; q < batch; ++q) o[done + q] = (uint8_t)(pack[q >> 2] >> (8 * (q & 3)));
        }
        done += batch;
    }
}

// ---------------------------------------------------------------- serial decoder
// Generic path for a stream without a side index: one thread, tree walk, the
// reference's "consumed bytes >= buffer_size" termination (huffman.c:344-361).
__global__ void huff_decode_serial_kernel(const uint32_t* __restrict__ words, uint64_t nwords, uint64_t buffer_size,
                                          const uint32_t* __restrict__ codes, const uint8_t* __restrict__ lens,
                                          uint8_t* __restrict__ out, uint64_t out_cap, uint64_t* __restrict__ count) {
    __shared__ int16_t left[512], right[512];
    __shared__ uint8_t symv[512];
    if (threadIdx.x != 0) return;
    int nn = 1; left[0] = right[0] = -1; symv[0] = 0;
    for (int s = 0; s < 256; ++s) {
        const int L = lens[s]; if (!L) continue;
        int cur = 0;
        for (int bit = L - 1; bit >= 0; --bit) {
            int16_t* nx = ((codes[s] >> bit) & 1) ? &right[cur] : &left[cur];
            if (*nx < 0) { *nx = (int16_t)nn; left[nn] = right[nn] = -1; symv[nn] = 0; ++nn; }
            cur = *nx;
        }
        symv[cur] = (uint8_t)s;
    }
    uint64_t consumed = 0, o = 0;
    if (left[0] < 0 || right[0] < 0) { *count = 0; return; }   // an empty or one-code table never consumes a bit: the walk below would spin
    do {
        int v = 0;
        while (left[v] >= 0 && right[v] >= 0) {
            const uint64_t w = consumed >> 5;
            const uint32_t word = w < nwords ? words[w] : 0;
            v = ((word >> (31 - (consumed & 31))) & 1) ? right[v] : left[v];
            ++consumed;
        }
        if (o < out_cap) out[o] = symv[v];
        ++o;
    } while ((consumed >> 3) < buffer_size);
    *count = o;
}

inline uint64_t align8(uint64_t x) { return (x + 7) & ~(uint64_t)7; }

}  // namespace

extern "C" int b200_huffman_layout(uint64_t n, uint64_t block_size, b200_huff_layout* L) {
    if (!L) { B200_SET_ERR("b200_huffman_layout: NULL"); return B200_ERR_ARG; }
    if (block_size && block_size < n && (block_size % CHUNK)) {
        B200_SET_ERR("huffman block_size %llu must be a multiple of %u", (unsigned long long)block_size, CHUNK);
        return B200_ERR_ARG;
    }
    const uint64_t bs = eff_block(n, block_size);
    L->nblocks = n ? (n + bs - 1) / bs : 1;
    L->nchunks = n ? (n + CHUNK - 1) / CHUNK : 1;
    L->chunks_per_block = bs / CHUNK;
    uint64_t o = 64;  // info[8] u64 lives at offset 0
    L->off_freq = o;       o += align8(L->nblocks * 256 * 4);
    L->off_codes = o;      o += align8(L->nblocks * 256 * 4);
    L->off_lens = o;       o += align8(L->nblocks * 256);
    L->off_tree = o;       o += align8(L->nblocks * 511 * 2 * 2);
    L->off_meta = o;       o += align8(L->nblocks * 4 * 4);
    L->off_block_bits = o; o += align8(L->nblocks * 8);
    L->off_block_word = o; o += align8((L->nblocks + 1) * 8);
    L->off_chunk_bits = o; o += align8(L->nchunks * 4);
    L->off_chunk_off = o;  o += align8((L->nchunks + 1) * 8);
    L->off_sub_off = o;    o += align8(L->nchunks * SUBS_PER_CHUNK * 4);
    o += align8((L->nchunks + 1) * 8);  // private prefix array P behind the public part
    L->bytes = o;
    return B200_OK;
}

extern "C" uint64_t b200_huffman_max_words(uint64_t n, uint64_t block_size) {
    // A Huffman code never beats a fixed 8-bit code on average, but per block the
    // reference only guarantees <= 32 bits/symbol; size for 8 bits/symbol + rounding
    // like init_bitwriter(size) (huffman.c:293) plus one word per block.
    const uint64_t bs = eff_block(n, block_size);
    const uint64_t nblocks = n ? (n + bs - 1) / bs : 1;
    return n / 4 + nblocks + 4;
}

static int huff_tables(b200_ctx* ctx, const uint8_t* d_in, uint64_t n, uint64_t block_size, uint8_t* d_side,
                       uint64_t side_bytes, b200_huff_layout* L, uint64_t* bs_out) {
    if (n == 0) { B200_SET_ERR("huffman: empty input (the reference exits: queue is empty, huffman.c:149-152)"); return B200_ERR_DOMAIN; }
    if ((reinterpret_cast<uintptr_t>(d_in) & 15) || (reinterpret_cast<uintptr_t>(d_side) & 7)) {
        B200_SET_ERR("huffman: d_in must be 16-byte and d_side 8-byte aligned"); return B200_ERR_ARG;
    }
    B200_TRY(b200_huffman_layout(n, block_size, L));
    if (side_bytes < L->bytes) { B200_SET_ERR("huffman: side buffer %llu < %llu", (unsigned long long)side_bytes, (unsigned long long)L->bytes); return B200_ERR_CAPACITY; }
    const uint64_t bs = eff_block(n, block_size);
    *bs_out = bs;
    // zero info + freq + codes + lens (contiguous at the front of the side buffer)
    CUDA_TRY(cudaMemsetAsync(d_side, 0, L->off_tree, ctx->stream));
    const uint32_t tpb = (uint32_t)((L->chunks_per_block + TILE_CHUNKS - 1) / TILE_CHUNKS);
    const uint64_t grid = L->nblocks * tpb;
    byte_hist_kernel<<<(unsigned)grid, 256, 0, ctx->stream>>>(d_in, n, bs, tpb, reinterpret_cast<uint32_t*>(d_side + L->off_freq));
    huff_build_kernel<256, 256, false><<<(unsigned)L->nblocks, 32, 0, ctx->stream>>>(
        reinterpret_cast<const uint32_t*>(d_side + L->off_freq), reinterpret_cast<uint32_t*>(d_side + L->off_codes),
        d_side + L->off_lens, reinterpret_cast<int16_t*>(d_side + L->off_tree), reinterpret_cast<uint32_t*>(d_side + L->off_meta));
    ctx->launches += 2;
    CUDA_TRY(cudaGetLastError());
    return B200_OK;
}

extern "C" int b200_huffman_tables_dev(b200_ctx* ctx, const uint8_t* d_in, uint64_t n, uint64_t block_size,
                                       uint8_t* d_side, uint64_t side_bytes) {
    B200_ENTER(ctx);
    b200_huff_layout L; uint64_t bs;
    return huff_tables(ctx, d_in, n, block_size, d_side, side_bytes, &L, &bs);
}

// decoder side of a stored stream (container.cu): with freq[] of every block already in d_side, rebuild codes, lengths,
// trees and meta with the same heap replay the encoder ran
extern "C" int b200_huffman_tables_from_freq_dev(b200_ctx* ctx, uint8_t* d_side, uint64_t side_bytes, uint64_t n, uint64_t block_size) {
    B200_ENTER(ctx);
    b200_huff_layout L;
    B200_TRY(b200_huffman_layout(n, block_size, &L));
    if (side_bytes < L.bytes) { B200_SET_ERR("huffman: side buffer too small"); return B200_ERR_CAPACITY; }
    huff_build_kernel<256, 256, false><<<(unsigned)L.nblocks, 32, 0, ctx->stream>>>(
        reinterpret_cast<const uint32_t*>(d_side + L.off_freq), reinterpret_cast<uint32_t*>(d_side + L.off_codes),
        d_side + L.off_lens, reinterpret_cast<int16_t*>(d_side + L.off_tree), reinterpret_cast<uint32_t*>(d_side + L.off_meta));
    ctx->launches += 1;
    CUDA_TRY(cudaGetLastError());
    return B200_OK;
}

// chunk bit counts -> offsets -> bit packing, with the tables already in the side buffer
static int huff_pack(b200_ctx* ctx, const uint8_t* d_in, uint64_t n, const b200_huff_layout& L, uint32_t* d_words,
                     uint64_t words_capacity, uint8_t* d_side, uint64_t* h_total_words, uint32_t* h_worst_status) {
    const uint32_t cpb = (uint32_t)L.chunks_per_block;
    uint64_t* info = reinterpret_cast<uint64_t*>(d_side);
    uint64_t* P = reinterpret_cast<uint64_t*>(d_side + L.off_sub_off + align8(L.nchunks * SUBS_PER_CHUNK * 4));
    huff_chunkbits_kernel<<<(unsigned)L.nchunks, 256, 0, ctx->stream>>>(d_in, n, cpb, d_side + L.off_lens,
                                                                         reinterpret_cast<uint32_t*>(d_side + L.off_chunk_bits));
    huff_offsets_kernel<<<1, 1024, 0, ctx->stream>>>(reinterpret_cast<const uint32_t*>(d_side + L.off_chunk_bits), L.nchunks, cpb,
                                                     L.nblocks, P, reinterpret_cast<uint64_t*>(d_side + L.off_block_bits),
                                                     reinterpret_cast<uint64_t*>(d_side + L.off_block_word), words_capacity, info);
    huff_chunkoff_kernel<<<(unsigned)((L.nchunks + 255) / 256), 256, 0, ctx->stream>>>(
        L.nchunks, cpb, P, reinterpret_cast<const uint64_t*>(d_side + L.off_block_word),
        reinterpret_cast<uint64_t*>(d_side + L.off_chunk_off), d_words, info);
    B200_TIMED_BEGIN(ctx, B200_K_HUFF_ENCODE);
    huff_encode_kernel<<<(unsigned)L.nchunks, 256, 0, ctx->stream>>>(
        d_in, n, cpb, reinterpret_cast<const uint32_t*>(d_side + L.off_codes), d_side + L.off_lens,
        reinterpret_cast<const uint32_t*>(d_side + L.off_meta), reinterpret_cast<const uint32_t*>(d_side + L.off_chunk_bits),
        reinterpret_cast<const uint64_t*>(d_side + L.off_chunk_off), reinterpret_cast<uint32_t*>(d_side + L.off_sub_off), d_words, info);
    B200_TIMED_END(ctx);
    ctx->launches += 4;
    CUDA_TRY(cudaGetLastError());
    if (h_total_words || h_worst_status) {
        uint64_t* pin; B200_TRY(b200_pinned(ctx, 16 + L.nblocks * 16, reinterpret_cast<void**>(&pin)));
        CUDA_TRY(cudaMemcpyAsync(pin, info, 16, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(pin + 2, d_side + L.off_meta, L.nblocks * 16, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        if (h_total_words) *h_total_words = pin[0];
        uint32_t worst = 0;
        const uint32_t* m = reinterpret_cast<const uint32_t*>(pin + 2);
        for (uint64_t b = 0; b < L.nblocks; ++b) if (m[4 * b] > worst) worst = m[4 * b];
        if (h_worst_status) *h_worst_status = worst;
        if (pin[1]) { B200_SET_ERR("huffman: stream needs %llu words, capacity %llu", (unsigned long long)pin[0], (unsigned long long)words_capacity); return B200_ERR_CAPACITY; }
    }
    return B200_OK;
}

extern "C" int b200_huffman_encode_dev(b200_ctx* ctx, const uint8_t* d_in, uint64_t n, uint64_t block_size,
                                       uint32_t* d_words, uint64_t words_capacity, uint8_t* d_side, uint64_t side_bytes,
                                       uint64_t* h_total_words, uint32_t* h_worst_status) {
    B200_ENTER(ctx);
    b200_huff_layout L; uint64_t bs;
    B200_TRY(huff_tables(ctx, d_in, n, block_size, d_side, side_bytes, &L, &bs));
    return huff_pack(ctx, d_in, n, L, d_words, words_capacity, d_side, h_total_words, h_worst_status);
}

// _huffman_compress (huffman.c:267-285): bit packing with the caller's code table, one
// table for the whole buffer. A symbol without a code makes the reference exit(1)
// (:274-277); here it is reported as status 3.
__global__ void __launch_bounds__(256) huff_check_codes_kernel(const uint32_t* __restrict__ freq, const uint8_t* __restrict__ lens,
                                                               uint32_t* __restrict__ meta) {
    const bool bad = freq[threadIdx.x] != 0 && lens[threadIdx.x] == 0;
    if (__syncthreads_or(bad) && threadIdx.x == 0) meta[0] = 3;
}

extern "C" int b200_huffman_encode_with_codes_dev(b200_ctx* ctx, const uint8_t* d_in, uint64_t n,
                                                  const uint32_t* h_codes, const uint8_t* h_lens,
                                                  uint32_t* d_words, uint64_t words_capacity, uint8_t* d_side, uint64_t side_bytes,
                                                  uint64_t* h_total_words, uint32_t* h_worst_status) {
    B200_ENTER(ctx);
    if (n == 0) { if (h_total_words) *h_total_words = 0; if (h_worst_status) *h_worst_status = 0; return B200_OK; }
    if ((reinterpret_cast<uintptr_t>(d_in) & 15) || (reinterpret_cast<uintptr_t>(d_side) & 7)) {
        B200_SET_ERR("huffman: d_in must be 16-byte and d_side 8-byte aligned"); return B200_ERR_ARG;
    }
    b200_huff_layout L;
    B200_TRY(b200_huffman_layout(n, 0, &L));
    if (side_bytes < L.bytes) { B200_SET_ERR("huffman: side buffer too small"); return B200_ERR_CAPACITY; }
    CUDA_TRY(cudaMemsetAsync(d_side, 0, L.off_block_bits, ctx->stream));   // info, freq, codes, lens, tree, meta
    uint8_t* pin; B200_TRY(b200_pinned(ctx, 4096, reinterpret_cast<void**>(&pin)));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));   // the pinned staging area may still be in flight
    memcpy(pin + 256, h_codes, 1024); memcpy(pin + 1280, h_lens, 256);
    CUDA_TRY(cudaMemcpyAsync(d_side + L.off_codes, pin + 256, 1024, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(d_side + L.off_lens, pin + 1280, 256, cudaMemcpyHostToDevice, ctx->stream));
    {   // the decoder walks a tree: rebuild it from the code table (root = node 0, leaf = {-1, symbol})
        int16_t* kids = reinterpret_cast<int16_t*>(pin + 1536);
        int nn = 1; kids[0] = kids[1] = -2;   // -2 = no child yet
        for (int sym = 0; sym < 256; ++sym) {
            const int len = h_lens[sym];
            if (!len) continue;
            if (len > 32) { B200_SET_ERR("huffman: code of symbol %d is longer than 32 bits", sym); return B200_ERR_ARG; }
            int v = 0;
            for (int bit = len - 1; bit >= 0; --bit) {
                if (kids[2 * v] == -1) { B200_SET_ERR("huffman: the code table is not prefix free"); return B200_ERR_ARG; }
                int16_t* nx = &kids[2 * v + ((h_codes[sym] >> bit) & 1)];
                if (*nx < 0) {
                    if (nn >= 511) { B200_SET_ERR("huffman: the code table needs more than 511 tree nodes"); return B200_ERR_ARG; }
                    *nx = (int16_t)nn; kids[2 * nn] = kids[2 * nn + 1] = -2; ++nn;
                }
                v = *nx;
            }
            if (kids[2 * v] != -2 || kids[2 * v + 1] != -2) { B200_SET_ERR("huffman: the code table is not prefix free"); return B200_ERR_ARG; }
            kids[2 * v] = -1; kids[2 * v + 1] = (int16_t)sym;
        }
        // an inner node with one missing child: point it at itself's sibling-less side as a leaf of symbol 0 so walks terminate
        for (int v = 0; v < nn; ++v) if (kids[2 * v] != -1) for (int c = 0; c < 2; ++c) if (kids[2 * v + c] == -2) {
            if (nn >= 511) { B200_SET_ERR("huffman: the code table needs more than 511 tree nodes"); return B200_ERR_ARG; }
            kids[2 * v + c] = (int16_t)nn; kids[2 * nn] = -1; kids[2 * nn + 1] = 0; ++nn;
        }
        B200_TRY(b200_copy_in(ctx, d_side + L.off_tree, kids, (size_t)nn * 4, ctx->stream));
    }
    const uint32_t tpb = (uint32_t)((L.chunks_per_block + TILE_CHUNKS - 1) / TILE_CHUNKS);
    byte_hist_kernel<<<tpb, 256, 0, ctx->stream>>>(d_in, n, eff_block(n, 0), tpb, reinterpret_cast<uint32_t*>(d_side + L.off_freq));
    huff_check_codes_kernel<<<1, 256, 0, ctx->stream>>>(reinterpret_cast<const uint32_t*>(d_side + L.off_freq), d_side + L.off_lens,
                                                        reinterpret_cast<uint32_t*>(d_side + L.off_meta));
    ctx->launches += 2;
    uint32_t worst = 0;
    const int rc = huff_pack(ctx, d_in, n, L, d_words, words_capacity, d_side, h_total_words, &worst);
    if (h_worst_status) *h_worst_status = worst;
    return rc;
}

// ---------------------------------------------------------------- one table over several shards
// The reference builds ONE tree over the whole input (huffman.c:184-211). When the input is sharded over ranks the
// shard histograms are summed (all-reduce of 256 x u64 by the caller), every rank builds the same table from the sum
// and packs its own shard from bit 0 of its own words; the shard streams are then spliced bit-granularly at the
// exclusive prefix of the shard bit counts, which yields the words huffman_compress writes for the whole buffer.
__global__ void __launch_bounds__(256) huff_widen_freq_kernel(const uint32_t* __restrict__ f32, uint64_t* __restrict__ f64) {
    f64[threadIdx.x] = f32[threadIdx.x];
}
__global__ void __launch_bounds__(256) huff_narrow_freq_kernel(const uint64_t* __restrict__ f64, uint32_t* __restrict__ f32) {
    f32[threadIdx.x] = (uint32_t)f64[threadIdx.x];   // `uint32_t frequencies[256]` (huffman.c:184) wraps the same way
}

extern "C" int b200_huffman_histogram_dev(b200_ctx* ctx, const uint8_t* d_in, uint64_t n, uint64_t* d_freq64) {
    B200_ENTER(ctx);
    if ((reinterpret_cast<uintptr_t>(d_in) & 15) || (reinterpret_cast<uintptr_t>(d_freq64) & 7)) {
        B200_SET_ERR("huffman: d_in must be 16-byte and d_freq64 8-byte aligned"); return B200_ERR_ARG;
    }
    if (n >> 32) { B200_SET_ERR("huffman histogram: a shard must be below 4 GiB (u32 bins)"); return B200_ERR_ARG; }
    uint32_t* f32; B200_TRY(b200_scratch(ctx, 0, 1024, reinterpret_cast<void**>(&f32)));
    CUDA_TRY(cudaMemsetAsync(f32, 0, 1024, ctx->stream));
    if (n) {
        const uint64_t bs = eff_block(n, 0);
        const uint32_t tpb = (uint32_t)((bs / CHUNK + TILE_CHUNKS - 1) / TILE_CHUNKS);
        byte_hist_kernel<<<tpb, 256, 0, ctx->stream>>>(d_in, n, bs, tpb, f32);
        ctx->launches += 1;
    }
    huff_widen_freq_kernel<<<1, 256, 0, ctx->stream>>>(f32, d_freq64);
    ctx->launches += 1;
    CUDA_TRY(cudaGetLastError());
    return B200_OK;
}

extern "C" int b200_huffman_encode_with_freq_dev(b200_ctx* ctx, const uint8_t* d_in, uint64_t n, const uint64_t* d_freq64,
                                                 uint32_t* d_words, uint64_t words_capacity, uint8_t* d_side, uint64_t side_bytes,
                                                 uint64_t* h_total_words, uint64_t* h_total_bits, uint32_t* h_worst_status) {
    B200_ENTER(ctx);
    if ((reinterpret_cast<uintptr_t>(d_in) & 15) || (reinterpret_cast<uintptr_t>(d_side) & 7)) {
        B200_SET_ERR("huffman: d_in must be 16-byte and d_side 8-byte aligned"); return B200_ERR_ARG;
    }
    b200_huff_layout L;
    B200_TRY(b200_huffman_layout(n, 0, &L));
    if (side_bytes < L.bytes) { B200_SET_ERR("huffman: side buffer %llu < %llu", (unsigned long long)side_bytes, (unsigned long long)L.bytes); return B200_ERR_CAPACITY; }
    CUDA_TRY(cudaMemsetAsync(d_side, 0, L.off_tree, ctx->stream));
    huff_narrow_freq_kernel<<<1, 256, 0, ctx->stream>>>(d_freq64, reinterpret_cast<uint32_t*>(d_side + L.off_freq));
    huff_build_kernel<256, 256, false><<<1, 32, 0, ctx->stream>>>(
        reinterpret_cast<const uint32_t*>(d_side + L.off_freq), reinterpret_cast<uint32_t*>(d_side + L.off_codes),
        d_side + L.off_lens, reinterpret_cast<int16_t*>(d_side + L.off_tree), reinterpret_cast<uint32_t*>(d_side + L.off_meta));
    ctx->launches += 2;
    CUDA_TRY(cudaGetLastError());
    if (n == 0) {   // an empty shard (more ranks than data): the table is there, the stream is empty
        CUDA_TRY(cudaMemsetAsync(d_side + L.off_block_bits, 0, L.bytes - L.off_block_bits, ctx->stream));
        if (h_total_words || h_total_bits || h_worst_status) CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        if (h_total_words) *h_total_words = 0;
        if (h_total_bits) *h_total_bits = 0;
        if (h_worst_status) {
            uint32_t* pin; B200_TRY(b200_pinned(ctx, 16, reinterpret_cast<void**>(&pin)));
            CUDA_TRY(cudaMemcpyAsync(pin, d_side + L.off_meta, 16, cudaMemcpyDeviceToHost, ctx->stream));
            CUDA_TRY(cudaStreamSynchronize(ctx->stream));
            *h_worst_status = pin[0];
        }
        return B200_OK;
    }
    B200_TRY(huff_pack(ctx, d_in, n, L, d_words, words_capacity, d_side, h_total_words, h_worst_status));
    if (h_total_bits) {
        uint64_t* pin; B200_TRY(b200_pinned(ctx, 16, reinterpret_cast<void**>(&pin)));
        CUDA_TRY(cudaMemcpyAsync(pin, d_side + L.off_block_bits, 8, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        *h_total_bits = pin[0];
    }
    return B200_OK;
}

// dst word j of the piece holds source bits [32 j - s, 32 j - s + 32), s = dst_bit % 32, MSB first (bit k of a stream
// is bit 31 - k % 32 of word k / 32, write_bits huffman.c:18-48). The first and the last word of the piece are shared
// with the neighbouring shards and are OR-ed atomically into the zeroed destination; the rest are plain stores.
__global__ void __launch_bounds__(256) huff_splice_kernel(uint32_t* __restrict__ dst, uint64_t dst_bit, const uint32_t* __restrict__ src,
                                                          uint64_t src_bits, uint64_t nwords) {
    const uint32_t s = (uint32_t)(dst_bit & 31);
    const uint64_t w0 = dst_bit >> 5, src_words = (src_bits + 31) >> 5;
    for (uint64_t j = (uint64_t)blockIdx.x * 256 + threadIdx.x; j < nwords; j += (uint64_t)gridDim.x * 256) {
        const uint32_t hi = (j >= 1 && j - 1 < src_words) ? __ldg(src + j - 1) : 0u;
        const uint32_t lo = j < src_words ? __ldg(src + j) : 0u;
        uint32_t v = s ? __funnelshift_r(lo, hi, s) : lo;   // (hi:lo) >> s, low word
        // keep only source bits below src_bits: piece bit t = 32 j + q is source bit 32 j + q - s
        const int64_t first = (int64_t)(j * 32) - (int64_t)s;                 // source bit index of this word's MSB
        const int64_t over = first + 32 - (int64_t)src_bits;                  // bits of this word past the end
        if (over > 0) v = over >= 32 ? 0u : (v & (0xFFFFFFFFu << over));
        if (first < 0) v &= 0xFFFFFFFFu >> (uint32_t)(-first);
        if (j == 0 || j + 1 == nwords) { if (v) atomicOr(dst + w0 + j, v); }
        else dst[w0 + j] = v;
    }
}

extern "C" int b200_huffman_splice_dev(b200_ctx* ctx, uint32_t* d_dst, uint64_t dst_words_capacity, uint64_t dst_bit,
                                       const uint32_t* d_src, uint64_t src_bits) {
    B200_ENTER(ctx);
    if (src_bits == 0) return B200_OK;
    const uint64_t nwords = ((dst_bit & 31) + src_bits + 31) >> 5;
    if ((dst_bit >> 5) + nwords > dst_words_capacity) { B200_SET_ERR("huffman splice: destination too small"); return B200_ERR_CAPACITY; }
    const unsigned grid = (unsigned)std::min<uint64_t>((nwords + 255) / 256, (uint64_t)ctx->sm_count * 8);
    huff_splice_kernel<<<grid, 256, 0, ctx->stream>>>(d_dst, dst_bit, d_src, src_bits, nwords);
    ctx->launches += 1;
    CUDA_TRY(cudaGetLastError());
    return B200_OK;
}

extern "C" int b200_huffman_decode_dev(b200_ctx* ctx, const uint32_t* d_words, uint64_t total_words, const uint8_t* d_side,
                                       uint64_t side_bytes, uint64_t n, uint64_t block_size, uint8_t* d_out) {
    B200_ENTER(ctx);
    if (n == 0) return B200_OK;
    b200_huff_layout L;
    B200_TRY(b200_huffman_layout(n, block_size, &L));
    if (side_bytes < L.bytes) { B200_SET_ERR("huffman decode: side buffer too small"); return B200_ERR_CAPACITY; }
    const uint64_t bs = eff_block(n, block_size);
    const uint32_t cpb = (uint32_t)L.chunks_per_block;
    const uint32_t tpb = (cpb + TILE_CHUNKS - 1) / TILE_CHUNKS;
    // a table scope decoded by four or more CTAs: its lookup table is built once and copied, instead of rebuilt by every CTA
    uint16_t* lut_g = nullptr;
    if (tpb >= 4) {
        B200_TRY(b200_scratch(ctx, 7, (size_t)L.nblocks * (1u << LUT_BITS) * 2 + 64, reinterpret_cast<void**>(&lut_g)));
        huff_decode_kernel<<<(unsigned)L.nblocks, 256, 0, ctx->stream>>>(
            d_words, total_words, n, bs, cpb, tpb, reinterpret_cast<const int16_t*>(d_side + L.off_tree),
            reinterpret_cast<const uint32_t*>(d_side + L.off_meta),
            reinterpret_cast<const uint64_t*>(d_side + L.off_chunk_off), reinterpret_cast<const uint32_t*>(d_side + L.off_sub_off), d_out, lut_g, 1);
        ctx->launches += 1;
    }
    B200_TIMED_BEGIN(ctx, B200_K_HUFF_DECODE);
    huff_decode_kernel<<<(unsigned)(L.nblocks * tpb), 256, 0, ctx->stream>>>(
        d_words, total_words, n, bs, cpb, tpb, reinterpret_cast<const int16_t*>(d_side + L.off_tree),
        reinterpret_cast<const uint32_t*>(d_side + L.off_meta),
        reinterpret_cast<const uint64_t*>(d_side + L.off_chunk_off), reinterpret_cast<const uint32_t*>(d_side + L.off_sub_off), d_out, lut_g, 0);
    B200_TIMED_END(ctx);
    ctx->launches += 1;
    CUDA_TRY(cudaGetLastError());
    return B200_OK;
}

extern "C" int b200_huffman_decode_serial_dev(b200_ctx* ctx, const uint32_t* d_words, uint64_t nwords, uint64_t buffer_size,
                                              const uint32_t* d_codes, const uint8_t* d_lens, uint8_t* d_out,
                                              uint64_t out_capacity, uint64_t* h_count) {
    B200_ENTER(ctx);
    uint64_t* d_cnt;
    B200_TRY(b200_scratch(ctx, 0, 64, reinterpret_cast<void**>(&d_cnt)));
    huff_decode_serial_kernel<<<1, 32, 0, ctx->stream>>>(d_words, nwords, buffer_size, d_codes, d_lens, d_out, out_capacity, d_cnt);
    ctx->launches += 1;
    CUDA_TRY(cudaGetLastError());
    uint64_t* pin; B200_TRY(b200_pinned(ctx, 64, reinterpret_cast<void**>(&pin)));
    CUDA_TRY(cudaMemcpyAsync(pin, d_cnt, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (h_count) *h_count = pin[0];
    return B200_OK;
}

// ================================================================================================================
// Zig-Huffman-compatible chunked mode: the file format of algorithms/huffman/zig_huffman/src/main.zig.
//   per 4 MiB chunk (BUFFER_SIZE, main.zig:5):
//     tree, pre-order: value u8 + freq u32 LE per node, then the left and the right subtree; a missing child is the
//       i32 -1 (main.zig:155-176). The tree is built by std.PriorityQueue over the histogram of the WHOLE 4 MiB read
//       buffer (main.zig:109-121 counts BUFFER_SIZE bytes whatever was read: a short last chunk also counts the stale
//       bytes the previous chunk left behind it; a first short chunk counts the allocator's zero pages).
//     CompressedSize u32 = last_block | bytes << 1 (main.zig:11-18,513-520), last_block = the read was short
//     payload: codes MSB-first into BYTES (main.zig:316-338), whole bytes only -- the trailing partial byte is dropped
//       (main.zig:523 writes compression_buffer[0..idx]), so the last symbol(s) of a chunk can be lost: the format is
//       reproduced as it is, bug included.
//   a file whose size is a multiple of 4 MiB ends with one more chunk of zero bytes (the read that finds the end).
// PARITY UNPINNED: no Zig toolchain here; the heap order is std.PriorityQueue's as restated in huff_shared.cuh /
// oracle/port/zig_huffman_port.c. The GPU must match that port byte for byte.
namespace {
constexpr uint64_t ZIG_CHUNK = 1ull << 22;

__global__ void __launch_bounds__(256) zig_bswap_kernel(uint32_t* __restrict__ w, uint64_t nwords) {
    for (uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x; i < nwords; i += (uint64_t)gridDim.x * 256) w[i] = __byte_perm(w[i], 0, 0x0123);
}

// one thread per chunk walks its bytes like main.zig:413-441: a symbol starts while byte_idx < size; bits past the
// payload read as zero (the reference reads whatever its buffer holds there)
__global__ void __launch_bounds__(32) zig_decode_kernel(const uint8_t* __restrict__ payload, const uint64_t* __restrict__ pay_off,
                                                        const uint32_t* __restrict__ pay_size, const int16_t* __restrict__ trees,
                                                        uint8_t* __restrict__ out, uint32_t* __restrict__ counts) {
    __shared__ int16_t kid[511][2];
    const uint32_t k = blockIdx.x;
    for (int i = threadIdx.x; i < 511 * 2; i += 32) (&kid[0][0])[i] = trees[(uint64_t)k * 511 * 2 + i];
    __syncwarp();
    if (threadIdx.x != 0) return;
    const uint8_t* p = payload + pay_off[k];
    const uint32_t size = pay_size[k];
    uint8_t* o = out + (uint64_t)k * ZIG_CHUNK;
    uint32_t byte_idx = 0, bit_idx = 0, cnt = 0;
    uint32_t cur = size ? p[0] : 0;
    while (byte_idx < size) {
        int v = 0;
        while (kid[v][0] >= 0) {   // inner node: both children present
            const uint32_t bit = (cur >> (7 - bit_idx)) & 1u;
            v = kid[v][bit];
            if (++bit_idx == 8) { bit_idx = 0; ++byte_idx; cur = byte_idx < size ? p[byte_idx] : 0u; }
        }
        if (cnt < ZIG_CHUNK) o[cnt] = (uint8_t)kid[v][1];
        ++cnt;
        if (kid[0][0] < 0) break;   // a root without children consumes no bits: the reference would spin
    }
    counts[k] = cnt;
}
}  // namespace

extern "C" uint64_t b200_zig_huffman_max_bytes(uint64_t n) {
    const uint64_t nchunks = n / ZIG_CHUNK + 1;
    return n + n / 4 + nchunks * (511 * 5 + 512 * 4 + 4 + 64) + 64;
}

// serialise the tree of one chunk pre-order; kids = i16[511][2] (leaf = {-1, symbol}); returns bytes written
static uint64_t zig_put_tree(const int16_t* kids, const uint32_t* freq, int v, uint8_t* o, uint32_t* f_out) {
    uint64_t w = 0;
    if (kids[2 * v] < 0) {
        const uint32_t f = freq[kids[2 * v + 1]];
        o[w++] = (uint8_t)kids[2 * v + 1]; memcpy(o + w, &f, 4); w += 4;
        memset(o + w, 0xFF, 8); w += 8;   // two missing children
        *f_out = f;
        return w;
    }
    uint8_t* head = o; w = 5;
    uint32_t fl = 0, fr = 0;
    w += zig_put_tree(kids, freq, kids[2 * v], o + w, &fl);
    w += zig_put_tree(kids, freq, kids[2 * v + 1], o + w, &fr);
    const uint32_t f = fl + fr;
    head[0] = 0; memcpy(head + 1, &f, 4);
    *f_out = f;
    return w;
}

extern "C" int b200_zig_huffman_compress_host(b200_ctx* ctx, const uint8_t* h_in, uint64_t n, uint8_t* h_out, uint64_t out_capacity,
                                              uint64_t* h_total_bytes) {
    B200_ENTER(ctx);
    const uint64_t nz = n / ZIG_CHUNK + 1;                     // chunks the reference writes (an exact multiple ends with an empty one)
    const uint64_t n_aug = nz * ZIG_CHUNK;
    const uint64_t nb = n ? (n + ZIG_CHUNK - 1) / ZIG_CHUNK : 0;   // chunks that carry data
    b200_huff_layout LA, LB;
    B200_TRY(b200_huffman_layout(n_aug, ZIG_CHUNK, &LA));
    uint8_t *d_aug, *d_sideA, *d_sideB = nullptr; uint32_t* d_words = nullptr;
    B200_TRY(b200_scratch(ctx, 8, n_aug + 64, reinterpret_cast<void**>(&d_aug)));
    B200_TRY(b200_scratch(ctx, 12, LA.bytes, reinterpret_cast<void**>(&d_sideA)));
    // the read buffer as the reference sees it chunk by chunk: the data, and behind a short read what the previous
    // chunk left there (zero pages before the first chunk)
    if (n) B200_TRY(b200_copy_in(ctx, d_aug, h_in, n, ctx->stream));
    {
        const uint64_t last = nz - 1, len_last = n - last * ZIG_CHUNK;
        if (last == 0) CUDA_TRY(cudaMemsetAsync(d_aug + len_last, 0, ZIG_CHUNK - len_last, ctx->stream));
        else CUDA_TRY(cudaMemcpyAsync(d_aug + last * ZIG_CHUNK + len_last, d_aug + (last - 1) * ZIG_CHUNK + len_last, ZIG_CHUNK - len_last,
                                      cudaMemcpyDeviceToDevice, ctx->stream));
    }
    CUDA_TRY(cudaMemsetAsync(d_sideA, 0, LA.off_tree, ctx->stream));
    const uint32_t tpb = (uint32_t)((LA.chunks_per_block + TILE_CHUNKS - 1) / TILE_CHUNKS);
    byte_hist_kernel<<<(unsigned)(LA.nblocks * tpb), 256, 0, ctx->stream>>>(d_aug, n_aug, ZIG_CHUNK, tpb, reinterpret_cast<uint32_t*>(d_sideA + LA.off_freq));
    huff_build_kernel<256, 256, false, true><<<(unsigned)LA.nblocks, 32, 0, ctx->stream>>>(
        reinterpret_cast<const uint32_t*>(d_sideA + LA.off_freq), reinterpret_cast<uint32_t*>(d_sideA + LA.off_codes),
        d_sideA + LA.off_lens, reinterpret_cast<int16_t*>(d_sideA + LA.off_tree), reinterpret_cast<uint32_t*>(d_sideA + LA.off_meta));
    ctx->launches += 2;
    CUDA_TRY(cudaGetLastError());
    // trees and histograms to the host; inputs the reference cannot handle are refused before anything is packed
    std::vector<uint8_t> sideA(LA.off_block_bits);
    CUDA_TRY(cudaMemcpyAsync(sideA.data(), d_sideA, LA.off_block_bits, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    const uint32_t* meta = reinterpret_cast<const uint32_t*>(sideA.data() + LA.off_meta);
    for (uint64_t k = 0; k < nz; ++k) {
        if (meta[k * 4] == 1) { B200_SET_ERR("zig huffman: chunk %llu has a single symbol (the reference shifts a u32 by 32, main.zig:222)", (unsigned long long)k); return B200_ERR_DOMAIN; }
        if (meta[k * 4 + 3] > 25) { B200_SET_ERR("zig huffman: chunk %llu needs a %u-bit code; the reference's byte-offset shift loses bits beyond 25 (main.zig:320)", (unsigned long long)k, meta[k * 4 + 3]); return B200_ERR_DOMAIN; }
    }
    uint64_t total_words = 0;
    std::vector<uint64_t> bbits(nb + 1), bword(nb + 2);
    if (nb) {
        B200_TRY(b200_huffman_layout(n, ZIG_CHUNK, &LB));
        // a table that also counts stale bytes is not the data's own, so the 8-bit bound of a self-built table does not hold
        const uint64_t capw = n / 2 + nb + 8;   // 16 bits per symbol; beyond that the call reports B200_ERR_CAPACITY
        B200_TRY(b200_scratch(ctx, 9, LB.bytes, reinterpret_cast<void**>(&d_sideB)));
        B200_TRY(b200_scratch(ctx, 11, capw * 4, reinterpret_cast<void**>(&d_words)));
        CUDA_TRY(cudaMemsetAsync(d_sideB, 0, LB.off_tree, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(d_sideB + LB.off_freq, d_sideA + LA.off_freq, nb * 1024, cudaMemcpyDeviceToDevice, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(d_sideB + LB.off_codes, d_sideA + LA.off_codes, nb * 1024, cudaMemcpyDeviceToDevice, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(d_sideB + LB.off_lens, d_sideA + LA.off_lens, nb * 256, cudaMemcpyDeviceToDevice, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(d_sideB + LB.off_tree, d_sideA + LA.off_tree, nb * 511 * 4, cudaMemcpyDeviceToDevice, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(d_sideB + LB.off_meta, d_sideA + LA.off_meta, nb * 16, cudaMemcpyDeviceToDevice, ctx->stream));
        uint32_t worst = 0;
        B200_TRY(huff_pack(ctx, d_aug, n, LB, d_words, capw, d_sideB, &total_words, &worst));
        if (total_words) {
            zig_bswap_kernel<<<(unsigned)std::min<uint64_t>((total_words + 255) / 256, (uint64_t)ctx->sm_count * 16), 256, 0, ctx->stream>>>(d_words, total_words);
            ctx->launches += 1;
            CUDA_TRY(cudaGetLastError());
        }
        CUDA_TRY(cudaMemcpyAsync(bbits.data(), d_sideB + LB.off_block_bits, nb * 8, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(bword.data(), d_sideB + LB.off_block_word, (nb + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    }
    uint64_t o = 0;
    for (uint64_t k = 0; k < nz; ++k) {
        if (o + 511 * 5 + 512 * 4 + 4 > out_capacity) { B200_SET_ERR("zig huffman: output buffer too small"); return B200_ERR_CAPACITY; }
        uint32_t froot = 0;
        o += zig_put_tree(reinterpret_cast<const int16_t*>(sideA.data() + LA.off_tree) + k * 511 * 2,
                          reinterpret_cast<const uint32_t*>(sideA.data() + LA.off_freq) + k * 256, (int)meta[k * 4 + 2], h_out + o, &froot);
        const uint64_t len_k = k < nb ? (k + 1 < nb || n % ZIG_CHUNK == 0 ? ZIG_CHUNK : n % ZIG_CHUNK) : 0;
        const uint64_t pay = k < nb ? bbits[k] / 8 : 0;
        const uint32_t hdr = (uint32_t)(len_k < ZIG_CHUNK ? 1u : 0u) | (uint32_t)(pay << 1);
        memcpy(h_out + o, &hdr, 4); o += 4;
        if (pay) {
            if (o + pay > out_capacity) { B200_SET_ERR("zig huffman: output buffer too small"); return B200_ERR_CAPACITY; }
            B200_TRY(b200_copy_out(ctx, h_out + o, reinterpret_cast<const uint8_t*>(d_words + bword[k]), pay, ctx->stream));
            o += pay;
        }
    }
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (h_total_bytes) *h_total_bytes = o;
    return B200_OK;
}

extern "C" int b200_zig_huffman_decompress_host(b200_ctx* ctx, const uint8_t* h_in, uint64_t bytes, uint8_t* h_out, uint64_t out_capacity,
                                                uint64_t* h_n) {
    B200_ENTER(ctx);
    // parse: trees + chunk table on the host
    std::vector<int16_t> trees;       // [chunk][511][2]
    std::vector<uint64_t> pay_off; std::vector<uint32_t> pay_size;
    uint64_t i = 0;
    bool done = false;
    while (!done) {
        const size_t base = trees.size();
        trees.resize(base + 511 * 2, -1);
        int16_t* kid = trees.data() + base;
        // iterative pre-order read (main.zig:178-200): stack of (node, which child comes next)
        int nn = 0, sp = 0; int stack[1024]; int side[1024];
        auto read_node = [&](int* out_v) -> int {
            if (i + 4 > bytes) return -1;
            int32_t m; memcpy(&m, h_in + i, 4);
            if (m == -1) { i += 4; *out_v = -1; return 0; }
            if (i + 5 > bytes || nn >= 511) return -1;
            const int v = nn++;
            kid[2 * v] = -1; kid[2 * v + 1] = (int16_t)h_in[i];   // leaf until a child shows up: {-1, symbol}
            i += 5;
            *out_v = v;
            return 0;
        };
        int root;
        if (read_node(&root) || root < 0) { B200_SET_ERR("zig huffman: corrupt tree at byte %llu", (unsigned long long)i); return B200_ERR_FORMAT; }
        stack[sp] = root; side[sp] = 0; ++sp;
        while (sp) {
            const int v = stack[sp - 1], s = side[sp - 1];
            if (s == 2) { --sp; continue; }
            side[sp - 1] = s + 1;
            int c;
            if (read_node(&c)) { B200_SET_ERR("zig huffman: corrupt tree at byte %llu", (unsigned long long)i); return B200_ERR_FORMAT; }
            if (c >= 0) {
                if (s == 0) { const int16_t sym = kid[2 * v + 1]; (void)sym; kid[2 * v] = (int16_t)c; kid[2 * v + 1] = -2; }   // becomes an inner node
                else {
                    if (kid[2 * v] < 0) { B200_SET_ERR("zig huffman: a node with one child at byte %llu", (unsigned long long)i); return B200_ERR_FORMAT; }
                    kid[2 * v + 1] = (int16_t)c;
                }
                if (sp >= 1023) { B200_SET_ERR("zig huffman: tree too deep"); return B200_ERR_FORMAT; }
                stack[sp] = c; side[sp] = 0; ++sp;
            } else if (s == 1 && kid[2 * v] >= 0) { B200_SET_ERR("zig huffman: a node with one child at byte %llu", (unsigned long long)i); return B200_ERR_FORMAT; }
        }
        for (int v = 0; v < nn; ++v) if (kid[2 * v] >= 0 && kid[2 * v + 1] < 0) { B200_SET_ERR("zig huffman: a node with one child"); return B200_ERR_FORMAT; }
        if (i + 4 > bytes) { B200_SET_ERR("zig huffman: truncated before a chunk header"); return B200_ERR_FORMAT; }
        uint32_t hdr; memcpy(&hdr, h_in + i, 4); i += 4;
        done = (hdr & 1u) != 0;
        const uint32_t size = hdr >> 1;
        if (i + size > bytes) { B200_SET_ERR("zig huffman: truncated payload"); return B200_ERR_FORMAT; }
        pay_off.push_back(i); pay_size.push_back(size);
        i += size;
    }
    const uint64_t nz = pay_off.size();
    uint8_t *d_pay, *d_out, *d_tab;
    B200_TRY(b200_scratch(ctx, 11, bytes + 64, reinterpret_cast<void**>(&d_pay)));
    B200_TRY(b200_scratch(ctx, 8, nz * ZIG_CHUNK + 64, reinterpret_cast<void**>(&d_out)));
    const uint64_t tab_bytes = nz * (511 * 4 + 8 + 4 + 4) + 64;
    B200_TRY(b200_scratch(ctx, 12, tab_bytes, reinterpret_cast<void**>(&d_tab)));
    int16_t* d_trees = reinterpret_cast<int16_t*>(d_tab);
    uint64_t* d_off = reinterpret_cast<uint64_t*>(d_tab + ((nz * 511 * 4 + 7) & ~7ull));
    uint32_t* d_size = reinterpret_cast<uint32_t*>(d_off + nz);
    uint32_t* d_cnt = d_size + nz;
    B200_TRY(b200_copy_in(ctx, d_pay, h_in, bytes, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(d_trees, trees.data(), nz * 511 * 4, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(d_off, pay_off.data(), nz * 8, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(d_size, pay_size.data(), nz * 4, cudaMemcpyHostToDevice, ctx->stream));
    zig_decode_kernel<<<(unsigned)nz, 32, 0, ctx->stream>>>(d_pay, d_off, d_size, d_trees, d_out, d_cnt);
    ctx->launches += 1;
    CUDA_TRY(cudaGetLastError());
    std::vector<uint32_t> cnt(nz);
    CUDA_TRY(cudaMemcpyAsync(cnt.data(), d_cnt, nz * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    uint64_t o = 0;
    for (uint64_t k = 0; k < nz; ++k) {
        const uint64_t c = cnt[k] < ZIG_CHUNK ? cnt[k] : ZIG_CHUNK;
        if (o + c > out_capacity) { B200_SET_ERR("zig huffman: output needs more than %llu bytes", (unsigned long long)out_capacity); return B200_ERR_CAPACITY; }
        if (c) B200_TRY(b200_copy_out(ctx, h_out + o, d_out + k * ZIG_CHUNK, c, ctx->stream));
        o += c;
    }
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (h_n) *h_n = o;
    return B200_OK;
}
