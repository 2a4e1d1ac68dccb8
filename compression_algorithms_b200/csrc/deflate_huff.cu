// Entropy stage of the deflate token stream for sm_100a: the step the reference leaves as
// "TODO: Build huffman tree and encode compressed buffer" (algorithms/deflate/lz77.c:279).
//
// What the reference defines is reproduced exactly:
//   frequencies[286]  algorithms/deflate/lz77.c:206,231,273 + huffman.c:49-62 (literal byte, or
//                     256 + clz16(offset) per match; lengths are not counted)
//   bit packing       write_bits, algorithms/deflate/huffman.c:18-48 (MSB-first u32 words)
// What it only declares (deflate/huffman.h:16-32,84-92) is completed as oracle/port/deflate_huff_port.c
// specifies (parity unpinned there): codes by the heap rule of algorithms/huffman/huffman.c:100-250
// over the 286 symbols; literal = code[byte]; match = code[256+k], the 15-k offset bits below the
// leading one, the length in 5 bits (MAX_LENGTH_BITS, deflate/lz77.h:7).
//
// Input: the compacted byte tokens of b200_lz77_encode_dev(variant 1) and their block offsets.
// Tokens are 2 or 4 bytes, so every token starts on a 2-byte "unit": a unit is a literal token, the
// head of a match (flag 1, offset low byte) or its tail (offset high byte, length). With
// h(u) = [first byte of unit u is 1], unit u is a tail iff the run of units with h = 1 that ends
// just before u has odd length; a chunk knows the state of its first unit from a per-block pass, so
// chunks are classified, measured, packed and decoded independently.
//
// Layout: one table scope per LZ block; cpb = ceil((2 * block_size + 2) / 4096) chunk slots of 4096 token
// bytes per block (slots past the block's tokens stay empty); streams are MSB-first u32 words, each
// block starting on a word; the side buffer (b200_dfl_layout) holds tables and the decode index.
#include "common.cuh"
#include "huff_shared.cuh"
#include "../../include/b200comp.h"

namespace {

constexpr uint32_t NSYM = 286, STR = 288, NODES = 2 * NSYM - 1;
constexpr uint32_t CHUNK_B = B200_DFL_CHUNK;       // token bytes per chunk (one CTA)
constexpr uint32_t CHUNK_U = CHUNK_B / 2;          // units per chunk
constexpr uint32_t UPT = 8;                        // units per encode thread (256 threads)
constexpr uint32_t SUB_U = B200_DFL_SUB / 2;       // units per decode thread
constexpr uint32_t SUBS_PER_CHUNK = CHUNK_U / SUB_U;
constexpr uint32_t TILE_CHUNKS = 16;               // chunks per decode CTA
constexpr uint32_t LUT_BITS = 12;
constexpr uint32_t STAGE_VEC = CHUNK_B / 16 + 2;   // uint4 per staged chunk: the chunk, its misalignment and one unit beyond

static_assert(CHUNK_U == 256 * UPT, "one chunk per 256-thread CTA");
static_assert(SUB_U == 16 * UPT, "a sub-chunk is 16 encode threads");

__device__ __forceinline__ uint32_t clz16(uint32_t x) { return x ? (uint32_t)__clz((int)x) - 16u : 16u; }

struct ChunkView {
    const uint16_t* u;     // unit j of the chunk = u[j] (low byte = first token byte), j <= nunits
    uint32_t nunits;       // valid units of this chunk
};

// Copies the chunk's token bytes (plus one unit beyond) into shared memory with aligned 16-byte
// loads; returns the view. sbuf: STAGE_VEC uint4. All 256 threads; ends with a __syncthreads.
__device__ __forceinline__ ChunkView stage_chunk(const uint8_t* __restrict__ tok, uint64_t tok_cap, uint64_t begin, uint64_t block_end,
                                                  uint4* sbuf) {
    const uint64_t a0 = begin & ~(uint64_t)15;
    for (uint32_t i = threadIdx.x; i < STAGE_VEC; i += 256) {
        const uint64_t a = a0 + (uint64_t)i * 16;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (a < block_end + 2) {
            if (a + 16 <= tok_cap) v = __ldg(reinterpret_cast<const uint4*>(tok + a));
            else {                                 // the last, partial vector of the token buffer
                uint32_t w[4] = {0, 0, 0, 0};
                for (uint32_t q = 0; q < 16 && a + q < tok_cap; ++q) w[q >> 2] |= (uint32_t)tok[a + q] << (8 * (q & 3));
                v = make_uint4(w[0], w[1], w[2], w[3]);
            }
        }
        sbuf[i] = v;
    }
    __syncthreads();
    ChunkView cv;
    cv.u = reinterpret_cast<const uint16_t*>(reinterpret_cast<const uint8_t*>(sbuf) + (begin - a0));
    const uint64_t left = (block_end - begin) / 2;
    cv.nunits = (uint32_t)(left < CHUNK_U ? left : CHUNK_U);
    return cv;
}

// Classifies the thread's UPT units. hb: 256 bytes of shared memory. Returns the thread's 8-bit
// head-flag mask and whether its first unit is a tail. s0 = state of the chunk's first unit.
// All 256 threads; contains one __syncthreads.
__device__ __forceinline__ uint32_t classify(const ChunkView& cv, uint32_t s0, uint8_t* hb, uint32_t* tail0) {
    const uint32_t t = threadIdx.x, u0 = t * UPT;
    uint32_t m = 0;
#pragma unroll
    for (uint32_t i = 0; i < UPT; ++i) if (u0 + i < cv.nunits && (cv.u[u0 + i] & 0xFF) == 1) m |= 1u << i;
    hb[t] = (uint8_t)m;
    __syncthreads();
    // length of the run of h = 1 units that ends just before unit u0
    uint32_t run = 0; int w = (int)t - 1; bool hit = false;
    while (w >= 0) {
        const uint32_t pm = hb[w];
        if (pm == 0xFF) { run += 8; --w; }
        else { run += (uint32_t)__clz((int)~(pm << 24)); hit = true; break; }
    }
    *tail0 = hit ? (run & 1u) : s0;    // an unbroken run back to the chunk start: parity of u0 (even) on top of s0
    return m;
}

// ---------------------------------------------------------------- K1 chunk states + histogram
// One CTA per LZ block, chunks in order (the tail state is carried from chunk to chunk).
__global__ void __launch_bounds__(256) dfl_hist_kernel(const uint8_t* __restrict__ tok, uint64_t tok_cap,
                                                       const uint64_t* __restrict__ tok_off, uint32_t cpb,
                                                       uint8_t* __restrict__ chunk_state, uint32_t* __restrict__ freq) {
    __shared__ uint4 sbuf[STAGE_VEC];
    __shared__ uint8_t hb[256];
    __shared__ uint32_t h[8][STR];
    __shared__ uint32_t carry;
    const uint64_t b = blockIdx.x;
    const uint64_t begin = tok_off[b], end = tok_off[b + 1];
    for (uint32_t i = threadIdx.x; i < 8 * STR; i += 256) (&h[0][0])[i] = 0;
    if (threadIdx.x == 0) carry = 0;
    uint32_t* my = h[threadIdx.x >> 5];
    uint32_t c = 0;
    for (uint64_t p = begin; p < end; p += CHUNK_B, ++c) {
        __syncthreads();                       // previous chunk's readers are done with sbuf / hb; carry is visible
        const ChunkView cv = stage_chunk(tok, tok_cap, p, end, sbuf);
        const uint32_t s0 = carry;
        if (threadIdx.x == 0) chunk_state[b * cpb + c] = (uint8_t)s0;
        uint32_t s;
        const uint32_t m = classify(cv, s0, hb, &s);
        const uint32_t u0 = threadIdx.x * UPT;
#pragma unroll
        for (uint32_t i = 0; i < UPT; ++i) {
            if (u0 + i < cv.nunits) {
                if (s) s = 0;
                else {
                    const uint32_t v = cv.u[u0 + i];
                    if ((m >> i) & 1u) { atomicAdd(&my[256 + clz16((v >> 8) | ((cv.u[u0 + i + 1] & 0xFFu) << 8))], 1u); s = 1; }
                    else atomicAdd(&my[v >> 8], 1u);
                }
            }
        }
        __syncthreads();                       // everybody has read carry (s0) before it is replaced
        if (threadIdx.x == 255) carry = s;
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < NSYM; i += 256) {
        uint32_t sum = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) sum += h[w][i];
        freq[b * STR + i] = sum;
    }
}

// bits of the item that starts at unit j (a literal or a match head)
__device__ __forceinline__ uint32_t item_bits(const ChunkView& cv, uint32_t j, bool head, const uint8_t* sl) {
    const uint32_t v = cv.u[j];
    if (!head) return sl[v >> 8];
    const uint32_t k = clz16((v >> 8) | ((cv.u[j + 1] & 0xFFu) << 8));
    return sl[256 + k] + (k < 15 ? 15 - k : 0) + 5;
}

// ---------------------------------------------------------------- K3 chunk bit counts
__global__ void __launch_bounds__(256) dfl_chunkbits_kernel(const uint8_t* __restrict__ tok, uint64_t tok_cap,
                                                            const uint64_t* __restrict__ tok_off, uint32_t cpb,
                                                            const uint8_t* __restrict__ chunk_state, const uint8_t* __restrict__ lens,
                                                            uint32_t* __restrict__ chunk_bits) {
    __shared__ uint4 sbuf[STAGE_VEC];
    __shared__ uint8_t hb[256];
    __shared__ uint8_t sl[STR];
    __shared__ uint32_t wsum[8];
    const uint64_t c = blockIdx.x, b = c / cpb;
    const uint64_t begin = tok_off[b] + (c % cpb) * (uint64_t)CHUNK_B, end = tok_off[b + 1];
    if (begin >= end) { if (threadIdx.x == 0) chunk_bits[c] = 0; return; }
    for (uint32_t i = threadIdx.x; i < STR; i += 256) sl[i] = lens[b * STR + i];
    const ChunkView cv = stage_chunk(tok, tok_cap, begin, end, sbuf);
    uint32_t s;
    const uint32_t m = classify(cv, chunk_state[c], hb, &s);
    const uint32_t u0 = threadIdx.x * UPT;
    uint32_t bits = 0;
#pragma unroll
    for (uint32_t i = 0; i < UPT; ++i) {
        if (u0 + i < cv.nunits) {
            if (s) s = 0;
            else { const bool head = (m >> i) & 1u; bits += item_bits(cv, u0 + i, head, sl); s = head; }
        }
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
// One CTA per chunk; thread t owns 8 consecutive units. A block exclusive scan of the item bits
// gives each thread its bit offset; items are packed MSB-first into a shared-memory image of the
// output words and stored coalesced (as huff_encode_kernel does for bytes).
__global__ void __launch_bounds__(256) dfl_encode_kernel(const uint8_t* __restrict__ tok, uint64_t tok_cap,
                                                         const uint64_t* __restrict__ tok_off, uint32_t cpb,
                                                         const uint8_t* __restrict__ chunk_state, const uint32_t* __restrict__ codes,
                                                         const uint8_t* __restrict__ lens, const uint32_t* __restrict__ meta,
                                                         const uint32_t* __restrict__ chunk_bits, const uint64_t* __restrict__ chunk_off,
                                                         uint32_t* __restrict__ sub_off, uint32_t* __restrict__ words,
                                                         const uint64_t* __restrict__ info) {
    __shared__ uint4 sbuf[STAGE_VEC];
    __shared__ uint8_t hb[256];
    __shared__ uint32_t sc[STR];
    __shared__ uint8_t  sl[STR];
    __shared__ uint32_t stage[CHUNK_U + 2];     // <= 32 bits per unit: a literal code is <= 32 bits, a match (2 units) <= 51
    __shared__ uint32_t wtot[8];
    if (info[1]) return;  // output does not fit
    const uint64_t c = blockIdx.x, b = c / cpb;
    const uint64_t begin = tok_off[b] + (c % cpb) * (uint64_t)CHUNK_B, end = tok_off[b + 1];
    if (begin >= end || meta[b * 4 + 0]) return;
    for (uint32_t i = threadIdx.x; i < STR; i += 256) { sc[i] = codes[b * STR + i]; sl[i] = lens[b * STR + i]; }
    const uint64_t A = chunk_off[c];
    const uint32_t T = chunk_bits[c];
    const uint32_t r = (uint32_t)(A & 31);
    const uint32_t nw = (r + T + 31) >> 5;
    for (uint32_t j = threadIdx.x; j < nw; j += 256) stage[j] = 0;
    const ChunkView cv = stage_chunk(tok, tok_cap, begin, end, sbuf);
    uint32_t tail0;
    const uint32_t m = classify(cv, chunk_state[c], hb, &tail0);
    const uint32_t u0 = threadIdx.x * UPT;
    uint32_t mybits = 0;
    {
        uint32_t s = tail0;
#pragma unroll
        for (uint32_t i = 0; i < UPT; ++i) {
            if (u0 + i < cv.nunits) {
                if (s) s = 0;
                else { const bool head = (m >> i) & 1u; mybits += item_bits(cv, u0 + i, head, sl); s = head; }
            }
        }
    }
    const uint32_t incl = warp_incl_scan_u32(mybits);
    if (lane_id() == 31) wtot[threadIdx.x >> 5] = incl;
    __syncthreads();
    uint32_t wbase = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) if ((uint32_t)w < (threadIdx.x >> 5)) wbase += wtot[w];
    const uint32_t ex = wbase + incl - mybits;
    // decode index: where the first item headed in this sub-chunk starts; bit 31 = its first unit is a tail
    if ((threadIdx.x & 15) == 0) sub_off[c * SUBS_PER_CHUNK + (threadIdx.x >> 4)] = ex | (tail0 << 31);

    if (mybits) {
        uint32_t pos = r + ex;
        uint32_t wi = pos >> 5;
        const uint32_t fill = pos & 31;  // bits of word wi that belong to earlier threads
        uint64_t acc = 0;                // left-aligned: top `have` bits are meaningful
        uint32_t have = fill;
        bool first = fill != 0;
        auto push = [&](uint32_t v, uint32_t L) {
            if (L) acc |= (uint64_t)v << (64 - have - L);
            have += L;
            if (have >= 32) {
                const uint32_t wv = (uint32_t)(acc >> 32);
                if (first) { atomicOr(&stage[wi], wv); first = false; }
                else stage[wi] = wv;
                ++wi; acc <<= 32; have -= 32;
            }
        };
        uint32_t s = tail0;
#pragma unroll
        for (uint32_t i = 0; i < UPT; ++i) {
            if (u0 + i < cv.nunits) {
                if (s) s = 0;
                else {
                    const uint32_t v = cv.u[u0 + i];
                    if ((m >> i) & 1u) {
                        const uint32_t nx = cv.u[u0 + i + 1];
                        const uint32_t off = (v >> 8) | ((nx & 0xFFu) << 8);
                        const uint32_t k = clz16(off), xb = k < 15 ? 15 - k : 0;
                        push(sc[256 + k], sl[256 + k]);
                        push(((off & ((1u << xb) - 1u)) << 5) | ((nx >> 8) & 31u), xb + 5);
                        s = 1;
                    } else push(sc[v >> 8], sl[v >> 8]);
                }
            }
        }
        if (have) atomicOr(&stage[wi], (uint32_t)(acc >> 32));
    }
    __syncthreads();
    uint32_t* dst = words + (A >> 5);
    // the last word is shared only if the next chunk slot of the same block holds tokens
    const bool tail_shared = ((r + T) & 31) != 0 && (c + 1) % cpb != 0 && begin + CHUNK_B < end;
    for (uint32_t j = threadIdx.x; j < nw; j += 256) {
        const uint32_t v = stage[j];
        if ((j == 0 && r != 0) || (j == nw - 1 && tail_shared)) atomicOr(&dst[j], v);
        else dst[j] = v;
    }
}

// ---------------------------------------------------------------- K6 decode (words -> byte tokens)
// One CTA per tile of 16 chunks of one block, one 256-byte sub-chunk of tokens per thread. 12-bit
// primary table (entry = 0x8000 | len << 9 | symbol); longer codes continue in the tree from the node
// the table names. Tokens are written as 2-byte units.
__global__ void __launch_bounds__(256) dfl_decode_kernel(const uint32_t* __restrict__ words, uint64_t total_words,
                                                         const uint64_t* __restrict__ tok_off, uint32_t cpb, uint32_t tiles_per_block,
                                                         const int16_t* __restrict__ tree, const uint32_t* __restrict__ meta,
                                                         const uint64_t* __restrict__ chunk_off, const uint32_t* __restrict__ sub_off,
                                                         uint8_t* __restrict__ tok_out) {
    __shared__ uint16_t lut[1u << LUT_BITS];
    __shared__ int16_t  kids[NODES][2];          // {left, right}; leaf = {-1, symbol}
    const uint64_t b = blockIdx.x / tiles_per_block, k = blockIdx.x % tiles_per_block;
    const uint64_t begin = tok_off[b], end = tok_off[b + 1];
    if (begin + k * (uint64_t)(TILE_CHUNKS * CHUNK_B) >= end) return;     // tile past the block's tokens
    {
        const uint32_t* t32 = reinterpret_cast<const uint32_t*>(tree + b * NODES * 2);
        uint32_t* k32 = reinterpret_cast<uint32_t*>(&kids[0][0]);
        for (uint32_t i = threadIdx.x; i < NODES; i += 256) k32[i] = t32[i];
    }
    const uint32_t root = meta[b * 4 + 2];
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < (1u << LUT_BITS); i += 256) {
        uint32_t v = root, d = 0;
        while (d < LUT_BITS && kids[v][0] >= 0) { v = (uint32_t)kids[v][(i >> (LUT_BITS - 1 - d)) & 1u]; ++d; }
        lut[i] = kids[v][0] < 0 ? (uint16_t)(0x8000u | (d << 9) | (uint32_t)(uint16_t)kids[v][1]) : (uint16_t)v;
    }
    __syncthreads();

    const uint32_t cib = (uint32_t)k * TILE_CHUNKS + (threadIdx.x >> 4);       // chunk slot inside the block
    const uint64_t chunk = b * cpb + cib;
    const uint64_t sub = chunk * SUBS_PER_CHUNK + (threadIdx.x & 15);
    const uint64_t unit0 = (uint64_t)cib * CHUNK_U + (uint64_t)(threadIdx.x & 15) * SUB_U;   // first unit, relative to the block
    const uint64_t nunits_block = (end - begin) / 2;
    if (cib >= cpb || unit0 >= nunits_block) return;
    const uint32_t count = (uint32_t)(nunits_block - unit0 < SUB_U ? nunits_block - unit0 : SUB_U);
    const uint32_t so = sub_off[sub];
    const uint64_t bitpos = chunk_off[chunk] + (so & 0x7FFFFFFFu);
    uint64_t wi = bitpos >> 5;
    uint64_t win = 0; uint32_t avail = 0;      // top `avail` bits of win are the next stream bits
    {
        const uint32_t sh = (uint32_t)(bitpos & 31);
        const uint64_t w0 = wi < total_words ? __ldg(&words[wi]) : 0; ++wi;
        win = w0 << (32 + sh); avail = 32 - sh;
        const uint64_t w1 = wi < total_words ? __ldg(&words[wi]) : 0; ++wi;
        win |= w1 << (32 - avail); avail += 32;
    }
    uint32_t nextw = wi < total_words ? __ldg(&words[wi]) : 0;   // one word ahead
    auto refill = [&]() {
        if (avail <= 32) {
            win |= (uint64_t)nextw << (32 - avail); avail += 32;
            ++wi;
            nextw = wi < total_words ? __ldg(&words[wi]) : 0;
        }
    };
    uint16_t* o = reinterpret_cast<uint16_t*>(tok_out + begin) + unit0;
    uint32_t u = so >> 31;                      // a leading tail unit is written by the previous sub-chunk
    while (u < count) {
        refill();
        const uint32_t e = lut[(uint32_t)(win >> (64 - LUT_BITS))];
        uint32_t s, L;
        if (e & 0x8000u) { s = e & 0x1FFu; L = (e >> 9) & 0xFu; }
        else {
            uint32_t v = e; L = LUT_BITS;
            while (kids[v][0] >= 0 && L < 40) { v = (uint32_t)kids[v][(uint32_t)(win >> (63 - L)) & 1u]; ++L; }
            s = (uint32_t)(uint16_t)kids[v][1] & 0x1FFu;
        }
        win <<= L; avail -= L;
        if (s < 256) { o[u] = (uint16_t)(s << 8); ++u; }
        else {
            refill();
            const uint32_t kk = s - 256, xb = kk < 15 ? 15 - kk : 0;
            const uint32_t x = (uint32_t)(win >> (64 - (xb + 5)));
            win <<= xb + 5; avail -= xb + 5;
            const uint32_t off = kk <= 15 ? ((1u << xb) | (x >> 5)) : 0u;
            o[u] = (uint16_t)(1u | ((off & 0xFFu) << 8));
            if (unit0 + u + 1 < nunits_block) o[u + 1] = (uint16_t)((off >> 8) | ((x & 31u) << 8));
            u += 2;
        }
    }
}

inline uint64_t align8(uint64_t x) { return (x + 7) & ~(uint64_t)7; }
inline uint64_t eff_bs(uint64_t n, uint64_t bs) { return (bs == 0 || bs > n) ? (n ? n : 1) : bs; }

}  // namespace

extern "C" int b200_dfl_layout_for(uint64_t n, uint64_t block_size, b200_dfl_layout* L) {
    if (!L) { B200_SET_ERR("b200_dfl_layout_for: NULL"); return B200_ERR_ARG; }
    const uint64_t bs = eff_bs(n, block_size);
    L->nblocks = n ? (n + bs - 1) / bs : 1;
    // worst case 2 * bs + 2 token bytes: all literals, then a match that starts on the block's last byte
    // (its word and extension read the zero padding behind the block, deflate/lz77.c:219,241)
    L->chunks_per_block = (2 * bs + 2 + CHUNK_B - 1) / CHUNK_B;
    L->nchunks = L->nblocks * L->chunks_per_block;
    uint64_t o = 64;  // info[8] u64 lives at offset 0
    L->off_freq = o;        o += align8(L->nblocks * STR * 4);
    L->off_codes = o;       o += align8(L->nblocks * STR * 4);
    L->off_lens = o;        o += align8(L->nblocks * STR);
    L->off_tree = o;        o += align8(L->nblocks * NODES * 2 * 2);
    L->off_meta = o;        o += align8(L->nblocks * 4 * 4);
    L->off_tok_off = o;     o += align8((L->nblocks + 1) * 8);
    L->off_tok_sizes = o;   o += align8(L->nblocks * 8);
    L->off_block_bits = o;  o += align8(L->nblocks * 8);
    L->off_block_word = o;  o += align8((L->nblocks + 1) * 8);
    L->off_chunk_state = o; o += align8(L->nchunks);
    L->off_chunk_bits = o;  o += align8(L->nchunks * 4);
    L->off_chunk_off = o;   o += align8((L->nchunks + 1) * 8);
    L->off_sub_off = o;     o += align8(L->nchunks * SUBS_PER_CHUNK * 4);
    o += align8((L->nchunks + 1) * 8);  // private prefix array P behind the public part
    L->bytes = o;
    return B200_OK;
}

extern "C" uint64_t b200_dfl_max_words(uint64_t n, uint64_t block_size) {
    // <= 32 bits per 2-byte unit, <= 2 * len + 2 token bytes and one word of rounding per block
    const uint64_t bs = eff_bs(n, block_size);
    const uint64_t nblocks = n ? (n + bs - 1) / bs : 1;
    return n + 2 * nblocks + 4;
}

extern "C" int b200_dfl_encode_dev(b200_ctx* ctx, const uint8_t* d_tokens, uint64_t tokens_capacity,
                                   const uint64_t* d_tok_off, const uint64_t* d_tok_sizes, uint64_t n, uint64_t block_size,
                                   uint32_t* d_words, uint64_t words_capacity, uint8_t* d_side, uint64_t side_bytes,
                                   uint64_t* h_total_words, uint32_t* h_worst_status) {
    B200_ENTER(ctx);
    if (n == 0) { if (h_total_words) *h_total_words = 0; if (h_worst_status) *h_worst_status = 0; return B200_OK; }
    if ((reinterpret_cast<uintptr_t>(d_tokens) & 15) || (reinterpret_cast<uintptr_t>(d_side) & 7)) {
        B200_SET_ERR("deflate entropy stage: d_tokens must be 16-byte and d_side 8-byte aligned"); return B200_ERR_ARG;
    }
    b200_dfl_layout L;
    B200_TRY(b200_dfl_layout_for(n, block_size, &L));
    if (side_bytes < L.bytes) { B200_SET_ERR("deflate entropy stage: side buffer %llu < %llu", (unsigned long long)side_bytes, (unsigned long long)L.bytes); return B200_ERR_CAPACITY; }
    const uint32_t cpb = (uint32_t)L.chunks_per_block;
    uint64_t* info = reinterpret_cast<uint64_t*>(d_side);
    uint64_t* P = reinterpret_cast<uint64_t*>(d_side + L.off_sub_off + align8(L.nchunks * SUBS_PER_CHUNK * 4));
    uint64_t* tok_off = reinterpret_cast<uint64_t*>(d_side + L.off_tok_off);
    CUDA_TRY(cudaMemsetAsync(d_side, 0, L.off_tree, ctx->stream));   // info, freq, codes, lens
    CUDA_TRY(cudaMemcpyAsync(tok_off, d_tok_off, (L.nblocks + 1) * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(d_side + L.off_tok_sizes, d_tok_sizes, L.nblocks * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    dfl_hist_kernel<<<(unsigned)L.nblocks, 256, 0, ctx->stream>>>(d_tokens, tokens_capacity, tok_off, cpb, d_side + L.off_chunk_state,
                                                                 reinterpret_cast<uint32_t*>(d_side + L.off_freq));
    huff_build_kernel<NSYM, STR, true><<<(unsigned)L.nblocks, 32, 0, ctx->stream>>>(
        reinterpret_cast<const uint32_t*>(d_side + L.off_freq), reinterpret_cast<uint32_t*>(d_side + L.off_codes),
        d_side + L.off_lens, reinterpret_cast<int16_t*>(d_side + L.off_tree), reinterpret_cast<uint32_t*>(d_side + L.off_meta));
    dfl_chunkbits_kernel<<<(unsigned)L.nchunks, 256, 0, ctx->stream>>>(d_tokens, tokens_capacity, tok_off, cpb, d_side + L.off_chunk_state,
                                                                      d_side + L.off_lens, reinterpret_cast<uint32_t*>(d_side + L.off_chunk_bits));
    huff_offsets_kernel<<<1, 1024, 0, ctx->stream>>>(reinterpret_cast<const uint32_t*>(d_side + L.off_chunk_bits), L.nchunks, cpb,
                                                     L.nblocks, P, reinterpret_cast<uint64_t*>(d_side + L.off_block_bits),
                                                     reinterpret_cast<uint64_t*>(d_side + L.off_block_word), words_capacity, info);
    huff_chunkoff_kernel<<<(unsigned)((L.nchunks + 255) / 256), 256, 0, ctx->stream>>>(
        L.nchunks, cpb, P, reinterpret_cast<const uint64_t*>(d_side + L.off_block_word),
        reinterpret_cast<uint64_t*>(d_side + L.off_chunk_off), d_words, info);
    B200_TIMED_BEGIN(ctx, B200_K_DFL_ENCODE);
    dfl_encode_kernel<<<(unsigned)L.nchunks, 256, 0, ctx->stream>>>(
        d_tokens, tokens_capacity, tok_off, cpb, d_side + L.off_chunk_state, reinterpret_cast<const uint32_t*>(d_side + L.off_codes),
        d_side + L.off_lens, reinterpret_cast<const uint32_t*>(d_side + L.off_meta), reinterpret_cast<const uint32_t*>(d_side + L.off_chunk_bits),
        reinterpret_cast<const uint64_t*>(d_side + L.off_chunk_off), reinterpret_cast<uint32_t*>(d_side + L.off_sub_off), d_words, info);
    B200_TIMED_END(ctx);
    ctx->launches += 6;
    CUDA_TRY(cudaGetLastError());
    if (h_total_words || h_worst_status) {
        uint64_t* pin; B200_TRY(b200_pinned(ctx, 16 + L.nblocks * 16, reinterpret_cast<void**>(&pin)));
        CUDA_TRY(cudaMemcpyAsync(pin, info, 16, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(pin + 2, d_side + L.off_meta, L.nblocks * 16, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        if (h_total_words) *h_total_words = pin[0];
        uint32_t worst = 0;
        const uint32_t* m = reinterpret_cast<const uint32_t*>(pin + 2);
        // status 1 here = a block without tokens (nothing to encode), not an error
        for (uint64_t b = 0; b < L.nblocks; ++b) if (m[4 * b] > 1 && m[4 * b] > worst) worst = m[4 * b];
        if (h_worst_status) *h_worst_status = worst;
        if (pin[1]) { B200_SET_ERR("deflate entropy stage: stream needs %llu words, capacity %llu", (unsigned long long)pin[0], (unsigned long long)words_capacity); return B200_ERR_CAPACITY; }
    }
    return B200_OK;
}

// decoder side of a stored stream (container.cu): frequencies[286] of every block -> codes, lengths, trees, meta
extern "C" int b200_dfl_tables_from_freq_dev(b200_ctx* ctx, uint8_t* d_side, uint64_t side_bytes, uint64_t n, uint64_t block_size) {
    B200_ENTER(ctx);
    b200_dfl_layout L;
    B200_TRY(b200_dfl_layout_for(n, block_size, &L));
    if (side_bytes < L.bytes) { B200_SET_ERR("deflate entropy stage: side buffer too small"); return B200_ERR_CAPACITY; }
    huff_build_kernel<NSYM, STR, true><<<(unsigned)L.nblocks, 32, 0, ctx->stream>>>(
        reinterpret_cast<const uint32_t*>(d_side + L.off_freq), reinterpret_cast<uint32_t*>(d_side + L.off_codes),
        d_side + L.off_lens, reinterpret_cast<int16_t*>(d_side + L.off_tree), reinterpret_cast<uint32_t*>(d_side + L.off_meta));
    ctx->launches += 1;
    CUDA_TRY(cudaGetLastError());
    return B200_OK;
}

extern "C" int b200_dfl_decode_dev(b200_ctx* ctx, const uint32_t* d_words, uint64_t total_words, const uint8_t* d_side,
                                   uint64_t side_bytes, uint64_t n, uint64_t block_size, uint8_t* d_tokens_out) {
    B200_ENTER(ctx);
    if (n == 0) return B200_OK;
    b200_dfl_layout L;
    B200_TRY(b200_dfl_layout_for(n, block_size, &L));
    if (side_bytes < L.bytes) { B200_SET_ERR("deflate entropy stage decode: side buffer too small"); return B200_ERR_CAPACITY; }
    if (reinterpret_cast<uintptr_t>(d_tokens_out) & 1) { B200_SET_ERR("deflate entropy stage decode: d_tokens_out must be 2-byte aligned"); return B200_ERR_ARG; }
    const uint32_t cpb = (uint32_t)L.chunks_per_block;
    const uint32_t tpb = (cpb + TILE_CHUNKS - 1) / TILE_CHUNKS;
    B200_TIMED_BEGIN(ctx, B200_K_DFL_DECODE);
    dfl_decode_kernel<<<(unsigned)(L.nblocks * tpb), 256, 0, ctx->stream>>>(
        d_words, total_words, reinterpret_cast<const uint64_t*>(d_side + L.off_tok_off), cpb, tpb,
        reinterpret_cast<const int16_t*>(d_side + L.off_tree), reinterpret_cast<const uint32_t*>(d_side + L.off_meta),
        reinterpret_cast<const uint64_t*>(d_side + L.off_chunk_off), reinterpret_cast<const uint32_t*>(d_side + L.off_sub_off), d_tokens_out);
    B200_TIMED_END(ctx);
    ctx->launches += 1;
    CUDA_TRY(cudaGetLastError());
    return B200_OK;
}

// deflate with the entropy stage: lz77_compress per block (deflate/lz77.c:199-277) -> byte tokens in
// the caller's token buffer -> frequencies, code tables and the packed stream (the TODO of :279)
extern "C" int b200_deflate_compress_dev(b200_ctx* ctx, const uint8_t* d_in, uint64_t n, uint64_t block_size,
                                         uint8_t* d_tokens, uint64_t tokens_capacity, uint64_t* d_tok_sizes, uint64_t* d_tok_off,
                                         uint32_t* d_words, uint64_t words_capacity, uint8_t* d_side, uint64_t side_bytes,
                                         uint64_t* h_total_words, uint32_t* h_worst_status) {
    B200_ENTER(ctx);
    B200_TRY(b200_lz77_encode_dev(ctx, B200_LZ_DEFLATE, d_in, n, block_size, d_tokens, tokens_capacity, d_tok_sizes, d_tok_off, nullptr));
    return b200_dfl_encode_dev(ctx, d_tokens, tokens_capacity, d_tok_off, d_tok_sizes, n, block_size, d_words, words_capacity,
                               d_side, side_bytes, h_total_words, h_worst_status);
}

// the decoder the reference never wrote (deflate/deflate.c:78-79 is empty): words -> byte tokens
// (d_tokens, scratch of at least the encoder's token bytes) -> the original bytes
extern "C" int b200_deflate_decompress_dev(b200_ctx* ctx, const uint32_t* d_words, uint64_t total_words, const uint8_t* d_side,
                                           uint64_t side_bytes, uint64_t n, uint64_t block_size, uint8_t* d_tokens, uint8_t* d_out) {
    B200_ENTER(ctx);
    if (n == 0) return B200_OK;
    b200_dfl_layout L;
    B200_TRY(b200_dfl_layout_for(n, block_size, &L));
    B200_TRY(b200_dfl_decode_dev(ctx, d_words, total_words, d_side, side_bytes, n, block_size, d_tokens));
    return b200_lz77_decode_dev(ctx, B200_LZ_DEFLATE, d_tokens, reinterpret_cast<const uint64_t*>(d_side + L.off_tok_off),
                                reinterpret_cast<const uint64_t*>(d_side + L.off_tok_sizes), n, block_size, d_out);
}
