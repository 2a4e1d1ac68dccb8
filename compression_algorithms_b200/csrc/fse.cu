// FSE (tANS, TABLE_LOG = 8): histogram, f64 normalisation, table build, encode and
// decode for sm_100a.
//
// Reference: /root/reference/algorithms/fse/src/main.zig (Zig, does not compile).
//   PINNED  : buildFrequencyTable :88-96 and normalizeFrequencyTable :106-149 --
//             reproduced bit for bit (IEEE f64 divide + multiply, truncation).
//   UNPINNED: everything after (table :151-189 is defective, encoder :42-68 is
//             unfinished, no decoder). Kept from the reference: TABLE_LOG 8, 256-state
//             table of {symbol u8, next_state u16, num_bits u8} (:73-82), single state
//             starting at 0 (:52), last byte raw in the first 8 bits (:55-56), symbols
//             visited from the end of the input (:59-62), final state flushed in 8 bits
//             (:65), LSB-first u64 words (:28-39). Chosen here: spread step 163 and the
//             standard tANS transition; oracle/port/fse_port.c is the written spec.
//
// Parallel layout: a "block" shares one table; it is cut into independent
// "segments", each one single-state stream in exactly the reference's stream shape
// and starting on a u64 word. One thread encodes/decodes one segment.
#include "common.cuh"
#include "hist.cuh"
#include "../../include/b200comp.h"

namespace {

constexpr uint32_t TSIZE = 256, TLOG = 8, SEG_TILE = 128;

inline uint64_t align8(uint64_t x) { return (x + 7) & ~(uint64_t)7; }
inline uint64_t eff_block(uint64_t n, uint64_t bs, uint64_t seg) {
    if (bs == 0 || bs >= n) { const uint64_t m = n ? n : 1; return (m + seg - 1) / seg * seg; }
    return bs;
}
inline uint64_t seg_stride_words(uint64_t seg) { return seg / 8 + 2; }  // 8 bits/symbol worst case + 16 bits

// ------------------------------------------------------------ normalise + tables
// One warp per block; lane l owns symbols 8l..8l+7.
// norm_in != NULL: the counts are already normalised (decoding a container), only the table is built.
__global__ void __launch_bounds__(32) fse_tables_kernel(const uint32_t* __restrict__ freq, const uint16_t* __restrict__ norm_in,
                                                       uint16_t* __restrict__ norm_out, uint32_t* __restrict__ tt_out) {
    __shared__ uint8_t  sym_at[TSIZE];
    __shared__ uint16_t next[256];
    const uint64_t b = blockIdx.x;
    const unsigned lane = threadIdx.x;
    uint64_t f[8];
    uint64_t total = 0; uint32_t present = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        f[k] = norm_in ? (uint64_t)norm_in[b * 256 + lane * 8 + k] : (uint64_t)freq[b * 256 + lane * 8 + k];
        total += f[k]; present += f[k] != 0;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) { total += __shfl_xor_sync(0xffffffffu, total, d); present += __shfl_xor_sync(0xffffffffu, present, d); }
    if (total == 0) {
#pragma unroll
        for (int k = 0; k < 8; ++k) { norm_out[b * 256 + lane * 8 + k] = 0; tt_out[b * 256 + lane * 8 + k] = 0; }
        return;
    }
    // main.zig:118-133 -- the only floating-point step: one divide, one multiply per symbol, truncation
    const double scale = __ddiv_rn((double)(TSIZE - present), __ull2double_rn(total));
    uint32_t nf[8]; uint32_t sum = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        uint32_t v = 0;
        if (f[k]) { v = (uint32_t)__double2ull_rz(__dmul_rn(__ull2double_rn(f[k]), scale)); if (v == 0) v = 1; }
        nf[k] = v; sum += v;
    }
    const uint32_t lane_sum = sum;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
    // main.zig:135-148 -- the remainder goes, one at a time, to the first strict maximum;
    // after the first increment that entry stays the maximum, so it receives all of it
    uint32_t best = 0, besti = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) if (nf[k] > best) { best = nf[k]; besti = lane * 8 + k; }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const uint32_t ob = __shfl_xor_sync(0xffffffffu, best, d), oi = __shfl_xor_sync(0xffffffffu, besti, d);
        if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
    }
    uint32_t remaining = TSIZE - sum;
    uint32_t my_sum = lane_sum;
    if (norm_in) {   // take the stored counts as they are (a corrupt table that does not sum to 256 decodes to garbage, not out of bounds)
        my_sum = 0; remaining = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) { nf[k] = (uint32_t)f[k]; my_sum += nf[k]; }
        if (total != TSIZE) {
#pragma unroll
            for (int k = 0; k < 8; ++k) { norm_out[b * 256 + lane * 8 + k] = 0; tt_out[b * 256 + lane * 8 + k] = 0; }
            return;
        }
    }
    if (best > 0 && (besti >> 3) == lane) { nf[besti & 7] += remaining; my_sum += remaining; }
    // cumulative counts
    const uint32_t incl = warp_incl_scan_u32(my_sum);
    uint32_t cum = incl - my_sum;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const uint32_t s = lane * 8 + k;
        norm_out[b * 256 + s] = (uint16_t)nf[k];
        next[s] = (uint16_t)nf[k];
        for (uint32_t j = cum; j < cum + nf[k]; ++j) sym_at[(j * 163u) & (TSIZE - 1)] = (uint8_t)s;
        cum += nf[k];
    }
    __syncwarp();
    if (lane == 0) {
        for (uint32_t u = 0; u < TSIZE; ++u) {
            const uint32_t s = sym_at[u];
            const uint32_t x = next[s]; next[s] = (uint16_t)(x + 1);
            const uint32_t nb = TLOG - (31u - (uint32_t)__clz(x));
            const uint32_t base = (x << nb) - TSIZE;
            tt_out[b * 256 + u] = s | (base << 8) | (nb << 24);
        }
    }
}

// ------------------------------------------------------------ encode
// One thread per segment; tables of the block in shared memory.
__global__ void __launch_bounds__(SEG_TILE) fse_encode_kernel(const uint8_t* __restrict__ in, uint64_t n, uint64_t bs, uint32_t seg,
                                                              uint32_t spb, uint32_t tiles_per_block,
                                                              const uint16_t* __restrict__ norm, const uint32_t* __restrict__ tt,
                                                              uint64_t* __restrict__ scratch, uint32_t stride_words,
                                                              uint32_t* __restrict__ seg_bits) {
    __shared__ uint16_t s_norm[256], s_cum[256];
    __shared__ uint8_t  s_enc[256];
    __shared__ uint32_t s_sym[256];   // per symbol: f | nb0 << 12 | (cum - f + 256) << 16
    const uint64_t b = blockIdx.x / tiles_per_block, tile = blockIdx.x % tiles_per_block;
    for (uint32_t s = threadIdx.x; s < 256; s += blockDim.x) s_norm[s] = norm[b * 256 + s];
    __syncthreads();
    if (threadIdx.x < 32) {
        uint32_t v[8], t = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) { v[k] = s_norm[threadIdx.x * 8 + k]; t += v[k]; }
        uint32_t c = warp_incl_scan_u32(t) - t;
#pragma unroll
        for (int k = 0; k < 8; ++k) { s_cum[threadIdx.x * 8 + k] = (uint16_t)c; c += v[k]; }
    }
    __syncthreads();
    for (uint32_t u = threadIdx.x; u < TSIZE; u += blockDim.x) {
        const uint32_t e = tt[b * 256 + u];
        const uint32_t s = e & 0xFF, nb = e >> 24, base = (e >> 8) & 0xFFFF;
        const uint32_t x = (base + TSIZE) >> nb;
        if (s_norm[s]) s_enc[s_cum[s] + x - s_norm[s]] = (uint8_t)u;
    }
    for (uint32_t s = threadIdx.x; s < 256; s += blockDim.x) {
        const uint32_t f = s_norm[s];
        const uint32_t nb0 = f ? TLOG - (31u - (uint32_t)__clz(f)) : 0u;
        s_sym[s] = f | (nb0 << 12) | ((uint32_t)(s_cum[s] + 256u - f) << 16);
    }
    __syncthreads();
    const uint32_t g = (uint32_t)tile * SEG_TILE + threadIdx.x;   // segment inside the block
    if (g >= spb) return;
    const uint64_t start = b * bs + (uint64_t)g * seg;
    uint64_t blk_end = (b + 1) * bs; if (blk_end > n) blk_end = n;
    if (start >= blk_end) return;
    const uint32_t len = (uint32_t)(blk_end - start < seg ? blk_end - start : seg);
    const uint64_t gseg = b * spb + g;
    uint32_t* out = reinterpret_cast<uint32_t*>(scratch + gseg * stride_words);   // the u64 words, written as little-endian u32 halves
    const uint8_t* src = in + start;

    // bit accumulator: lo holds the next 32 bits of the stream, hi what spills over (a symbol adds <= 8 bits)
    uint32_t lo = src[len - 1], hi = 0;   // last byte raw in the first 8 bits (main.zig:55-56)
    uint32_t accbits = 8, w = 0;
    uint32_t X = TSIZE;                    // state value 0 (main.zig:52)
    auto put = [&](uint32_t bits, uint32_t nb) {
        lo |= bits << accbits;
        hi |= __funnelshift_l(bits, 0u, accbits);        // bits >> (32 - accbits), 0 when accbits == 0
        accbits += nb;
        if (accbits >= 32) { out[w++] = lo; lo = hi; hi = 0; accbits -= 32; }
    };
    auto step = [&](uint32_t s) {
        const uint32_t e = s_sym[s];
        const uint32_t f = e & 0xFFFu;
        uint32_t nb = (e >> 12) & 0xFu;
        nb -= (X >> nb) < f;
        put(X & ((1u << nb) - 1u), nb);
        X = TSIZE + s_enc[(e >> 16) + (X >> nb) - 256u];
    };
    int i = (int)len - 2;
    while (i >= 0 && (((uint32_t)i + 1u) & 15u)) { step(src[i]); --i; }   // down to a 16-byte boundary
    if (i >= 15) {
        uint4 nxt = __ldg(reinterpret_cast<const uint4*>(src + i - 15));
        while (i >= 15) {
            const uint4 c = nxt;
            if (i >= 31) nxt = __ldg(reinterpret_cast<const uint4*>(src + i - 31));   // one 16-byte read ahead
            const uint32_t cw[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
            for (int k = 15; k >= 0; --k) step((cw[k >> 2] >> (8 * (k & 3))) & 0xFFu);
            i -= 16;
        }
    }
    put(X - TSIZE, TLOG);                  // final state in TABLE_LOG bits (main.zig:65)
    const uint32_t total = w * 32 + accbits;
    if (accbits) { out[w++] = lo; }
    if (w & 1) out[w] = 0;                 // the upper half of the last u64 word
    seg_bits[gseg] = total;
}

// ------------------------------------------------------------ offsets (two levels) + gather
// Word offset of every segment = exclusive scan of ceil(bits / 64). Tiles of 2048 segments are scanned by many CTAs
// (eight consecutive segments per thread, coalesced 32-byte reads), the few tile totals by one warp-sized step of one
// CTA; the tile base is added where the offsets are consumed (gather) or by fse_add_base_kernel (decode index).
constexpr uint32_t OFF_TILE = 2048;
__global__ void __launch_bounds__(256) fse_tile_scan_kernel(const uint32_t* __restrict__ seg_bits, uint64_t nsegs,
                                                            uint64_t* __restrict__ seg_word, uint64_t* __restrict__ tile_tot) {
    __shared__ uint32_t wtot[8];
    const uint64_t i0 = (uint64_t)blockIdx.x * OFF_TILE + (uint64_t)threadIdx.x * 8;
    uint32_t v[8], sum = 0;
    if (i0 + 8 <= nsegs) {
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(seg_bits + i0)), b = __ldg(reinterpret_cast<const uint4*>(seg_bits + i0 + 4));
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = i0 + k < nsegs ? seg_bits[i0 + k] : 0u;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) { v[k] = (v[k] + 63u) >> 6; sum += v[k]; }
    const uint32_t incl = warp_incl_scan_u32(sum);
    if (lane_id() == 31) wtot[threadIdx.x >> 5] = incl;
    __syncthreads();
    uint32_t base = 0, tot = 0;
#pragma unroll
    for (uint32_t w = 0; w < 8; ++w) { const uint32_t x = wtot[w]; if (w < (threadIdx.x >> 5)) base += x; tot += x; }
    uint32_t run = base + incl - sum;
#pragma unroll
    for (int k = 0; k < 8; ++k) { if (i0 + k < nsegs) seg_word[i0 + k] = run; run += v[k]; }
    if (threadIdx.x == 0) tile_tot[blockIdx.x] = tot;
}

// exclusive scan of the tile totals (in place: tile_tot[t] becomes the base of tile t), total and capacity check
__global__ void __launch_bounds__(1024) fse_tile_base_kernel(uint64_t* __restrict__ tile_tot, uint64_t ntiles, uint64_t nsegs,
                                                             uint64_t* __restrict__ seg_word, uint64_t capacity, uint64_t* __restrict__ info) {
    __shared__ uint64_t warp_tot[33];
    const uint64_t total = cta_exscan_1024(ntiles, warp_tot,
        [&](uint64_t i) { return tile_tot[i]; }, [&](uint64_t i, uint64_t ex) { tile_tot[i] = ex; });
    if (threadIdx.x == 0) { seg_word[nsegs] = total; info[0] = total; info[1] = total > capacity ? 1 : 0; }
}

__global__ void __launch_bounds__(256) fse_add_base_kernel(uint64_t* __restrict__ seg_word, uint64_t nsegs, const uint64_t* __restrict__ tile_base) {
    const uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    if (i < nsegs) seg_word[i] += tile_base[i / OFF_TILE];
}

// one warp per segment: copies its words to their place in the stream and makes the segment's offset absolute
__global__ void __launch_bounds__(256) fse_gather_kernel(const uint64_t* __restrict__ scratch, uint32_t stride_words,
                                                         const uint32_t* __restrict__ seg_bits, uint64_t* __restrict__ seg_word,
                                                         const uint64_t* __restrict__ tile_base,
                                                         uint64_t nsegs, uint64_t* __restrict__ out, const uint64_t* __restrict__ info) {
    const uint64_t g = (uint64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (g >= nsegs) return;
    const uint64_t off = seg_word[g] + tile_base[g / OFF_TILE];
    __syncwarp();
    if (lane_id() == 0) seg_word[g] = off;
    if (info[1]) return;
    const uint32_t nw = (seg_bits[g] + 63) >> 6;
    const uint64_t* src = scratch + g * stride_words;
    uint64_t* dst = out + off;
    for (uint32_t i = lane_id(); i < nw; i += 32) dst[i] = src[i];
}

// ------------------------------------------------------------ decode
__device__ __forceinline__ uint32_t get_bits(const uint64_t* __restrict__ w, uint64_t pos, uint32_t nb) {
    if (!nb) return 0;
    const uint64_t i = pos >> 6; const uint32_t o = (uint32_t)(pos & 63);
    uint64_t v = __ldg(w + i) >> o;
    if (o + nb > 64) v |= __ldg(w + i + 1) << (64 - o);
    return (uint32_t)(v & ((1ull << nb) - 1));
}

__global__ void __launch_bounds__(SEG_TILE) fse_decode_kernel(const uint64_t* __restrict__ words, uint64_t n, uint64_t bs, uint32_t seg,
                                                              uint32_t spb, uint32_t tiles_per_block, const uint32_t* __restrict__ tt,
                                                              const uint32_t* __restrict__ seg_bits, const uint64_t* __restrict__ seg_word,
                                                              uint8_t* __restrict__ out, uint32_t* __restrict__ bad) {
    __shared__ uint32_t s_tt[TSIZE];
    const uint64_t b = blockIdx.x / tiles_per_block, tile = blockIdx.x % tiles_per_block;
    for (uint32_t u = threadIdx.x; u < TSIZE; u += blockDim.x) s_tt[u] = tt[b * 256 + u];
    __syncthreads();
    const uint32_t g = (uint32_t)tile * SEG_TILE + threadIdx.x;
    if (g >= spb) return;
    const uint64_t start = b * bs + (uint64_t)g * seg;
    uint64_t blk_end = (b + 1) * bs; if (blk_end > n) blk_end = n;
    if (start >= blk_end) return;
    const uint32_t len = (uint32_t)(blk_end - start < seg ? blk_end - start : seg);
    const uint64_t gseg = b * spb + g;
    const uint64_t* w = words + seg_word[gseg];
    const uint32_t T = seg_bits[gseg];
    uint8_t* o = out + start;
    if (T < 16) { atomicAdd(bad, 1u); return; }
    // The stream is read from its end towards bit 8, as 32-bit words (the u64 words are little endian). hi:lo is a
    // 64-bit window of the bits just below the cursor (bit 63 = the bit right below it), refilled one word at a time
    // from a word fetched ahead: a symbol costs one table read, two funnel shifts and no wait for memory. A corrupt
    // stream reads zeros below its first word and fails the final check (cursor back at bit 8 in state 0).
    const uint32_t* w32 = reinterpret_cast<const uint32_t*>(w);
    const uint32_t last32 = (T - 1) >> 5;
    const uint32_t pos0 = T - TLOG;
    const int jw = (int)(pos0 >> 5);
    const uint32_t sh = pos0 & 31u;
    auto ldw = [&](int k) -> uint32_t { return k >= 0 ? __ldg(w32 + k) : 0u; };
    const uint32_t wA = ldw(jw), wB = ldw(jw - 1);
    uint32_t u = __funnelshift_r(wA, (uint32_t)jw + 1u <= last32 ? __ldg(w32 + jw + 1) : 0u, sh) & 0xFFu;
    uint32_t hi = __funnelshift_r(wB, wA, sh), lo = __funnelshift_r(0u, wB, sh);   // the sh bits of word jw below the cursor + word jw - 1
    int nxi = jw - 2;
    uint32_t nxt = ldw(nxi);
    uint32_t avail = 32u + sh, consumed = 0;
    uint32_t i = 0;
    const bool aligned = (reinterpret_cast<uintptr_t>(o) & 15) == 0;
    while (i + 1 < len) {
        uint32_t pack[4] = {0, 0, 0, 0};
        const uint32_t batch = len - 1 - i < 16 ? len - 1 - i : 16;
#pragma unroll
        for (uint32_t q = 0; q < 16; ++q) {                   // fully unrolled: pack[] stays in registers
            if (q < batch) {
                const uint32_t e = s_tt[u & 0xFFu];
                const uint32_t nb = e >> 24;
                pack[q >> 2] |= (e & 0xFFu) << (8 * (q & 3));
                const uint32_t bits = __funnelshift_l(hi, 0u, nb);          // the nb bits below the cursor
                hi = __funnelshift_l(lo, hi, nb);
                lo <<= nb;
                avail -= nb; consumed += nb;
                u = ((e >> 8) & 0xFFFFu) + bits;
                if (avail <= 32) {
                    const uint32_t s2 = 32u - avail;
                    lo = nxt << s2;
                    hi |= __funnelshift_l(nxt, 0u, s2);
                    avail += 32;
                    --nxi;
                    nxt = ldw(nxi);
                }
            }
        }
        if (batch == 16 && aligned) *reinterpret_cast<uint4*>(o + i) = make_uint4(pack[0], pack[1], pack[2], pack[3]);
        else for (uint32_t q = 0; q < batch; ++q) o[i + q] = (uint8_t)(pack[q >> 2] >> (8 * (q & 3)));
        i += batch;
    }
    const uint32_t pos = consumed <= pos0 ? pos0 - consumed : 0xFFFFFFFFu;
    o[len - 1] = (uint8_t)(__ldg(w) & 0xFF);
    if (!(pos == 8 && u == 0)) atomicAdd(bad, 1u);   // a valid stream lands back on state 0 at bit 8
}

int fse_args(uint64_t n, uint64_t block_size, uint64_t seg, uint64_t* bs_out) {
    if (seg < 64 || seg > 65536 || (seg & (seg - 1))) { B200_SET_ERR("fse: seg_size must be a power of two in [64, 65536]"); return B200_ERR_ARG; }
    if (block_size && block_size < n && (block_size % seg || block_size % 4096)) {
        B200_SET_ERR("fse: block_size must be a multiple of seg_size and of 4096"); return B200_ERR_ARG;
    }
    *bs_out = eff_block(n, block_size, seg);
    return B200_OK;
}

}  // namespace

extern "C" int b200_fse_layout_for(uint64_t n, uint64_t block_size, uint64_t seg_size, b200_fse_layout* L) {
    uint64_t bs;
    B200_TRY(fse_args(n, block_size, seg_size, &bs));
    L->nblocks = n ? (n + bs - 1) / bs : 1;
    L->segs_per_block = bs / seg_size;
    L->nsegs = L->nblocks * L->segs_per_block;
    uint64_t o = 64;
    L->off_freq = o;     o += align8(L->nblocks * 256 * 4);
    L->off_norm = o;     o += align8(L->nblocks * 256 * 2);
    L->off_tt = o;       o += align8(L->nblocks * 256 * 4);
    L->off_seg_bits = o; o += align8(L->nsegs * 4);
    L->off_seg_word = o; o += align8((L->nsegs + 1) * 8);
    L->bytes = o;
    return B200_OK;
}

extern "C" uint64_t b200_fse_max_words(uint64_t n, uint64_t seg_size) {
    const uint64_t nsegs = (n + seg_size - 1) / seg_size + 1;
    return n / 8 + nsegs * 2 + 4;
}

static int fse_front(b200_ctx* ctx, const uint8_t* d_in, uint64_t n, uint64_t block_size, uint64_t seg_size, uint8_t* d_side,
                     uint64_t side_bytes, b200_fse_layout* L, uint64_t* bs) {
    if (n == 0) { B200_SET_ERR("fse: empty input"); return B200_ERR_DOMAIN; }
    if ((reinterpret_cast<uintptr_t>(d_in) & 15) || (reinterpret_cast<uintptr_t>(d_side) & 7)) { B200_SET_ERR("fse: d_in must be 16-byte and d_side 8-byte aligned"); return B200_ERR_ARG; }
    B200_TRY(b200_fse_layout_for(n, block_size, seg_size, L));
    if (side_bytes < L->bytes) { B200_SET_ERR("fse: side buffer too small"); return B200_ERR_CAPACITY; }
    B200_TRY(fse_args(n, block_size, seg_size, bs));
    CUDA_TRY(cudaMemsetAsync(d_side, 0, L->off_norm, ctx->stream));  // info + histogram
    const uint32_t tpb = (uint32_t)((*bs + HIST_TILE - 1) / HIST_TILE);
    byte_hist_kernel<<<(unsigned)(L->nblocks * tpb), 256, 0, ctx->stream>>>(d_in, n, *bs, tpb, reinterpret_cast<uint32_t*>(d_side + L->off_freq));
    fse_tables_kernel<<<(unsigned)L->nblocks, 32, 0, ctx->stream>>>(reinterpret_cast<const uint32_t*>(d_side + L->off_freq), nullptr,
                                                                   reinterpret_cast<uint16_t*>(d_side + L->off_norm),
                                                                   reinterpret_cast<uint32_t*>(d_side + L->off_tt));
    ctx->launches += 2;
    CUDA_TRY(cudaGetLastError());
    return B200_OK;
}

extern "C" int b200_fse_normalize_dev(b200_ctx* ctx, const uint8_t* d_in, uint64_t n, uint64_t block_size, uint8_t* d_side, uint64_t side_bytes) {
    B200_ENTER(ctx);
    b200_fse_layout L; uint64_t bs;
    return fse_front(ctx, d_in, n, block_size, 1024, d_side, side_bytes, &L, &bs);
}

extern "C" int b200_fse_encode_dev(b200_ctx* ctx, const uint8_t* d_in, uint64_t n, uint64_t block_size, uint64_t seg_size,
                                   uint64_t* d_words, uint64_t words_capacity, uint8_t* d_side, uint64_t side_bytes,
                                   uint64_t* h_total_words) {
    B200_ENTER(ctx);
    b200_fse_layout L; uint64_t bs;
    B200_TRY(fse_front(ctx, d_in, n, block_size, seg_size, d_side, side_bytes, &L, &bs));
    const uint32_t spb = (uint32_t)L.segs_per_block;
    const uint32_t tpb = (spb + SEG_TILE - 1) / SEG_TILE;
    const uint32_t threads = spb < SEG_TILE ? ((spb + 31) / 32 * 32) : SEG_TILE;
    const uint32_t stride = (uint32_t)seg_stride_words(seg_size);
    uint64_t* scratch;
    B200_TRY(b200_scratch(ctx, 5, (size_t)L.nsegs * stride * 8 + 64, reinterpret_cast<void**>(&scratch)));
    uint32_t* seg_bits = reinterpret_cast<uint32_t*>(d_side + L.off_seg_bits);
    uint64_t* seg_word = reinterpret_cast<uint64_t*>(d_side + L.off_seg_word);
    uint64_t* info = reinterpret_cast<uint64_t*>(d_side);
    CUDA_TRY(cudaMemsetAsync(seg_bits, 0, L.nsegs * 4, ctx->stream));
    B200_TIMED_BEGIN(ctx, B200_K_FSE_ENCODE);
    fse_encode_kernel<<<(unsigned)(L.nblocks * tpb), threads, 0, ctx->stream>>>(
        d_in, n, bs, (uint32_t)seg_size, spb, tpb, reinterpret_cast<const uint16_t*>(d_side + L.off_norm),
        reinterpret_cast<const uint32_t*>(d_side + L.off_tt), scratch, stride, seg_bits);
    B200_TIMED_END(ctx);
    const uint64_t ntiles = (L.nsegs + OFF_TILE - 1) / OFF_TILE;
    uint64_t* tile_tot;
    B200_TRY(b200_scratch(ctx, 6, 64 + (size_t)ntiles * 8 + 64, reinterpret_cast<void**>(&tile_tot)));
    tile_tot += 8;   // (the first 64 bytes of slot 6 are the decoder's bad-segment counter)
    fse_tile_scan_kernel<<<(unsigned)ntiles, 256, 0, ctx->stream>>>(seg_bits, L.nsegs, seg_word, tile_tot);
    fse_tile_base_kernel<<<1, 1024, 0, ctx->stream>>>(tile_tot, ntiles, L.nsegs, seg_word, words_capacity, info);
    fse_gather_kernel<<<(unsigned)((L.nsegs + 7) / 8), 256, 0, ctx->stream>>>(scratch, stride, seg_bits, seg_word, tile_tot, L.nsegs, d_words, info);
    ctx->launches += 4;
    CUDA_TRY(cudaGetLastError());
    if (h_total_words) {
        uint64_t* pin; B200_TRY(b200_pinned(ctx, 64, reinterpret_cast<void**>(&pin)));
        CUDA_TRY(cudaMemcpyAsync(pin, info, 16, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        *h_total_words = pin[0];
        if (pin[1]) { B200_SET_ERR("fse: stream needs %llu words, capacity %llu", (unsigned long long)pin[0], (unsigned long long)words_capacity); return B200_ERR_CAPACITY; }
    }
    return B200_OK;
}

extern "C" int b200_fse_decode_dev(b200_ctx* ctx, const uint64_t* d_words, const uint8_t* d_side, uint64_t side_bytes,
                                   uint64_t n, uint64_t block_size, uint64_t seg_size, uint8_t* d_out, uint32_t* h_bad_segments) {
    B200_ENTER(ctx);
    if (n == 0) { if (h_bad_segments) *h_bad_segments = 0; return B200_OK; }
    b200_fse_layout L; uint64_t bs;
    B200_TRY(b200_fse_layout_for(n, block_size, seg_size, &L));
    B200_TRY(fse_args(n, block_size, seg_size, &bs));
    if (side_bytes < L.bytes) { B200_SET_ERR("fse decode: side buffer too small"); return B200_ERR_CAPACITY; }
    const uint32_t spb = (uint32_t)L.segs_per_block;
    const uint32_t tpb = (spb + SEG_TILE - 1) / SEG_TILE;
    const uint32_t threads = spb < SEG_TILE ? ((spb + 31) / 32 * 32) : SEG_TILE;
    uint32_t* d_bad;
    B200_TRY(b200_scratch(ctx, 6, 64, reinterpret_cast<void**>(&d_bad)));
    CUDA_TRY(cudaMemsetAsync(d_bad, 0, 4, ctx->stream));
    B200_TIMED_BEGIN(ctx, B200_K_FSE_DECODE);
    fse_decode_kernel<<<(unsigned)(L.nblocks * tpb), threads, 0, ctx->stream>>>(
        d_words, n, bs, (uint32_t)seg_size, spb, tpb, reinterpret_cast<const uint32_t*>(d_side + L.off_tt),
        reinterpret_cast<const uint32_t*>(d_side + L.off_seg_bits), reinterpret_cast<const uint64_t*>(d_side + L.off_seg_word), d_out, d_bad);
    B200_TIMED_END(ctx);
    ctx->launches += 1;
    CUDA_TRY(cudaGetLastError());
    if (h_bad_segments) {
        uint32_t* pin; B200_TRY(b200_pinned(ctx, 64, reinterpret_cast<void**>(&pin)));
        CUDA_TRY(cudaMemcpyAsync(pin, d_bad, 4, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        *h_bad_segments = pin[0];
    }
    return B200_OK;
}

// Rebuild what the decoder needs from a container's stored tables: tt from the normalised
// counts, seg_word from seg_bits. d_side holds norm and seg_bits already (layout L).
extern "C" int b200_fse_rebuild_index_dev(b200_ctx* ctx, uint64_t n, uint64_t block_size, uint64_t seg_size,
                                          uint8_t* d_side, uint64_t side_bytes) {
    B200_ENTER(ctx);
    b200_fse_layout L;
    B200_TRY(b200_fse_layout_for(n, block_size, seg_size, &L));
    if (side_bytes < L.bytes) { B200_SET_ERR("fse: side buffer too small"); return B200_ERR_CAPACITY; }
    uint64_t* info = reinterpret_cast<uint64_t*>(d_side);
    fse_tables_kernel<<<(unsigned)L.nblocks, 32, 0, ctx->stream>>>(nullptr, reinterpret_cast<const uint16_t*>(d_side + L.off_norm),
                                                                  reinterpret_cast<uint16_t*>(d_side + L.off_norm),
                                                                  reinterpret_cast<uint32_t*>(d_side + L.off_tt));
    const uint64_t ntiles = (L.nsegs + OFF_TILE - 1) / OFF_TILE;
    uint64_t* tile_tot;
    B200_TRY(b200_scratch(ctx, 6, 64 + (size_t)ntiles * 8 + 64, reinterpret_cast<void**>(&tile_tot)));
    tile_tot += 8;
    uint64_t* seg_word = reinterpret_cast<uint64_t*>(d_side + L.off_seg_word);
    fse_tile_scan_kernel<<<(unsigned)ntiles, 256, 0, ctx->stream>>>(reinterpret_cast<const uint32_t*>(d_side + L.off_seg_bits), L.nsegs, seg_word, tile_tot);
    fse_tile_base_kernel<<<1, 1024, 0, ctx->stream>>>(tile_tot, ntiles, L.nsegs, seg_word, ~0ull, info);
    fse_add_base_kernel<<<(unsigned)((L.nsegs + 255) / 256), 256, 0, ctx->stream>>>(seg_word, L.nsegs, tile_tot);
    ctx->launches += 4;
    CUDA_TRY(cudaGetLastError());
    return B200_OK;
}

// normalizeFrequencyTable / buildTransitionTable on one table (main.zig:106-189): exactly one of
// h_freq (raw counts -> normalise, then build) and h_norm_in (already normalised) is given.
extern "C" int b200_fse_tables_host(b200_ctx* ctx, const uint32_t* h_freq, const uint16_t* h_norm_in,
                                    uint16_t* h_norm_out, uint32_t* h_tt_out) {
    B200_ENTER(ctx);
    if ((h_freq != nullptr) == (h_norm_in != nullptr)) { B200_SET_ERR("fse: give either counts or normalised counts"); return B200_ERR_ARG; }
    uint8_t* d; uint8_t* pin;
    B200_TRY(b200_scratch(ctx, 6, 4096, reinterpret_cast<void**>(&d)));
    B200_TRY(b200_pinned(ctx, 4096, reinterpret_cast<void**>(&pin)));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    uint32_t* d_freq = reinterpret_cast<uint32_t*>(d + 64);
    uint16_t* d_nin = reinterpret_cast<uint16_t*>(d + 64 + 1024);
    uint16_t* d_nout = reinterpret_cast<uint16_t*>(d + 64 + 1536);
    uint32_t* d_tt = reinterpret_cast<uint32_t*>(d + 64 + 2048);
    if (h_freq) { memcpy(pin, h_freq, 1024); CUDA_TRY(cudaMemcpyAsync(d_freq, pin, 1024, cudaMemcpyHostToDevice, ctx->stream)); }
    else { memcpy(pin, h_norm_in, 512); CUDA_TRY(cudaMemcpyAsync(d_nin, pin, 512, cudaMemcpyHostToDevice, ctx->stream)); }
    fse_tables_kernel<<<1, 32, 0, ctx->stream>>>(h_freq ? d_freq : nullptr, h_freq ? nullptr : d_nin, d_nout, d_tt);
    ctx->launches += 1;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(pin + 1024, d_nout, 512, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(pin + 2048, d_tt, 1024, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (h_norm_out) memcpy(h_norm_out, pin + 1024, 512);
    if (h_tt_out) memcpy(h_tt_out, pin + 2048, 1024);
    return B200_OK;
}
