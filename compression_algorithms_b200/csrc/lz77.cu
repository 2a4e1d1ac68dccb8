// LZ77 match finder + greedy parse + token emission + decoders for sm_100a.
//
// Reference being replaced (bit-exact on the same block segmentation):
//   algorithms/lz77/lz77.c      hash :13-41, insert :55-86, find :94-108,
//                               greedy parse :264-345, decoder :347-377
//   algorithms/deflate/lz77.c   insert :77-145, find :147-174, byte tokens :176-197,
//                               greedy parse :199-280
//
// The reference's 2^20-slot open-addressed table is emulated slot-exactly with the
// "lazy expiry" rule: an entry placed for position i is
// live at time P iff i >= P - W, plus the slot-0 exception caused by the reference's
// early is_full flip. No eviction stores, no clearing between blocks (epoch tags).
//
// v1 layout: one warp owns one 2^20(+guard)-slot table of 8-byte slots
// {pattern, epoch<<22 | index} in HBM and walks its blocks sequentially; the 32
// lanes read 32 consecutive slots of a probe chain per step (256 B, coalesced) and
// pick the first non-live / first matching slot with ballot + ffs.
#include "common.cuh"
#include "../../include/b200comp.h"
#include <stdlib.h>

namespace {

constexpr uint32_t TABLE_SLOTS = 1u << 20;
constexpr uint32_t GUARD = 65536u + 64u;
constexpr uint32_t NONE = 0xFFFFFFFFu;
constexpr uint32_t IDX_BITS = 22;               // blocks up to 4 MiB
constexpr uint32_t IDX_MASK = (1u << IDX_BITS) - 1;
constexpr uint32_t MAX_EPOCH = (1u << (32 - IDX_BITS)) - 1;
constexpr uint32_t CLRQ = 1024;                 // pending slot-0 clear times per warp
constexpr uint64_t MAX_BLOCK = 1ull << IDX_BITS;

template <int V> struct Cfg;
template <> struct Cfg<0> { static constexpr uint32_t W = 1u << 14, MAXLEN = 15; };
template <> struct Cfg<1> { static constexpr uint32_t W = 1u << 15, MAXLEN = 31; };

__device__ __forceinline__ uint32_t byte_at(const uint8_t* __restrict__ d, uint32_t len, uint32_t p) {
    return p < len ? (uint32_t)__ldg(d + p) : 0u;  // bytes past the block read as 0 (U1)
}
__device__ __forceinline__ uint32_t word_at(const uint8_t* __restrict__ d, uint32_t len, uint32_t p) {
    return byte_at(d, len, p) | (byte_at(d, len, p + 1) << 8) | (byte_at(d, len, p + 2) << 16) | (byte_at(d, len, p + 3) << 24);
}

// slot tag: blocks up to 4 MiB carry an epoch (no clearing between blocks); larger blocks
// (whole-buffer calls of the standalone codec) use tag = index + 1 on a table that the host
// clears before the launch, one block per warp
template <bool BIG> __device__ __forceinline__ bool tag_live(uint32_t tag, uint32_t epoch, uint32_t W, uint32_t p) {
    if (BIG) return tag != 0u && (uint64_t)(tag - 1u) + W >= p;
    return (tag >> IDX_BITS) == epoch && (tag & IDX_MASK) + W >= p;
}
template <bool BIG> __device__ __forceinline__ uint32_t tag_index(uint32_t tag) { return BIG ? tag - 1u : (tag & IDX_MASK); }
template <bool BIG> __device__ __forceinline__ uint32_t tag_make(uint32_t epoch, uint32_t q) { return BIG ? q + 1u : ((epoch << IDX_BITS) | q); }

// ------------------------------------------------------------------ parse (v1)
template <int V, bool BIG>
__global__ void __launch_bounds__(128) lz77_parse_kernel(const uint8_t* __restrict__ in, uint64_t n, uint64_t bs, uint64_t nblocks,
                                                        uint2* __restrict__ tables, uint32_t* __restrict__ clrq, uint32_t epoch0,
                                                        uint8_t* __restrict__ scratch, uint64_t stride,
                                                        uint64_t* __restrict__ block_sizes, uint64_t* __restrict__ block_bytes,
                                                        uint32_t* __restrict__ err) {
    constexpr uint32_t W = Cfg<V>::W, MAXLEN = Cfg<V>::MAXLEN;
    const unsigned lane = threadIdx.x & 31;
    const uint64_t gw = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    uint2* T = tables + gw * (uint64_t)(TABLE_SLOTS + GUARD);
    uint32_t* Q = clrq + gw * CLRQ;
    uint32_t epoch = epoch0;

    for (uint64_t b = gw; b < nblocks; b += nwarps) {
        ++epoch;
        const uint8_t* d = in + b * bs;
        const uint32_t len = (uint32_t)(n - b * bs < bs ? n - b * bs : bs);
        uint8_t* out = scratch + b * stride;
        // FIFO of times at which slot 0 must be cleared (after that insert). The
        // reference reads the never-written ring cell at insert W-1 (lz77.c:70-76).
        uint32_t qh = 0, qt = 1, next_clear = W - 1;
        if (lane == 0) Q[0] = W - 1;
        uint32_t p = 0, pf = 0;
        uint64_t acc = 0; uint32_t accbits = 0; uint32_t opos = 0;  // V==0: word index; V==1: byte index

        while (p < len) {
            // warm L2 with the home slot of the next positions (each position once)
            if (p + 32 > pf) {
                const uint32_t q = pf + lane;
                if (q < len) {
                    const uint2* a = &T[lz_hash(word_at(d, len, q))];
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
                }
                pf += 32;
            }
            const uint32_t pat = word_at(d, len, p);
            const uint32_t h = lz_hash(pat);
            // ---- find(pat): first slot from h that is non-live or holds pat (no wrap)
            uint32_t m = NONE, s = h, stop_slot = 0;
            bool stop_dead = false;
            for (;;) {
                const uint2 e = __ldcg(&T[s + lane]);
                const bool live = tag_live<BIG>(e.y, epoch, W, p);
                const bool stop = !live || e.x == pat;
                const unsigned mask = __ballot_sync(0xffffffffu, stop);
                if (mask) {
                    const int k = __ffs(mask) - 1;
                    const uint32_t idx = __shfl_sync(0xffffffffu, tag_index<BIG>(e.y), k);
                    const bool lv = __shfl_sync(0xffffffffu, (int)live, k) != 0;
                    m = lv ? idx : NONE;
                    stop_slot = s + k; stop_dead = !lv;
                    break;
                }
                s += 32;
            }
            const bool reject = (m == NONE) || (V ? (p - m >= W - 1) : (p - m == W));
            uint32_t mlen = 1;
            if (!reject) {
                bool neq = true;
                if (lane < MAXLEN - 4) neq = byte_at(d, len, m + 4 + lane) != byte_at(d, len, p + 4 + lane);
                const unsigned mask = __ballot_sync(0xffffffffu, neq);
                mlen = 4 + (uint32_t)(__ffs(mask) - 1);
            }
            // ---- emit token
            if (lane == 0) {
                if (V == 0) {
                    if (reject) { acc |= (uint64_t)(byte_at(d, len, p) << 1) << accbits; accbits += 9; }
                    else { acc |= (uint64_t)(1u | ((p - m) << 1) | (mlen << 15)) << accbits; accbits += 19; }
                    if (accbits >= 32) {
                        reinterpret_cast<uint32_t*>(out)[opos++] = (uint32_t)acc;
                        acc >>= 32; accbits -= 32;
                    }
                } else {
                    if (reject) { *reinterpret_cast<uint16_t*>(out + opos) = (uint16_t)(byte_at(d, len, p) << 8); opos += 2; }
                    else {
                        const uint32_t off = p - m;
                        *reinterpret_cast<uint16_t*>(out + opos) = (uint16_t)(1u | ((off & 0xFF) << 8));
                        *reinterpret_cast<uint16_t*>(out + opos + 2) = (uint16_t)((off >> 8) | (mlen << 8));
                        opos += 4;
                    }
                }
            }
            // ---- insert every covered position, in order
            for (uint32_t q = p; q < p + mlen; ++q) {
                const uint32_t pq = (q == p) ? pat : word_at(d, len, q);
                uint32_t slot = 0, t;
                bool have = false;
                if (q == p && (V == 0 || stop_slot < TABLE_SLOTS)) {
                    // the find walk already proved every slot before stop_slot live
                    if (stop_dead) { slot = stop_slot; have = true; }
                    t = stop_slot;
                } else {
                    t = lz_hash(pq);
                }
                while (!have) {
                    const uint32_t at = V ? ((t + lane) & (TABLE_SLOTS - 1)) : (t + lane);  // deflate insert wraps (deflate/lz77.c:99-101)
                    const uint2 e = __ldcg(&T[at]);
                    const bool live = tag_live<BIG>(e.y, epoch, W, q);
                    const unsigned mask = __ballot_sync(0xffffffffu, !live);
                    if (mask) {
                        const uint32_t k = (uint32_t)(__ffs(mask) - 1);
                        slot = V ? ((t + k) & (TABLE_SLOTS - 1)) : (t + k);
                        have = true;
                    }
                    t += 32;
                }
                if (lane == 0) __stcg(&T[slot], make_uint2(pq, tag_make<BIG>(epoch, q)));
                if (slot == 0) {
                    if (qt - qh < CLRQ) { if (lane == 0) Q[qt % CLRQ] = q + W; }
                    else if (lane == 0) atomicOr(err, 2u);
                    if (qt == qh) next_clear = q + W;
                    ++qt;
                }
                if (q == next_clear) {
                    if (lane == 0) __stcg(&T[0], make_uint2(0u, 0u));  // kills even an entry placed just now
                    ++qh;
                    __syncwarp();
                    next_clear = (qh != qt) ? __ldcg(&Q[qh % CLRQ]) : NONE;
                }
                __syncwarp();
            }
            p += mlen;
        }
        if (lane == 0) {
            if (V == 0) {
                const uint64_t bits = (uint64_t)opos * 32 + accbits;
                reinterpret_cast<uint32_t*>(out)[opos] = (uint32_t)acc;        // tail bits, zero padded
                reinterpret_cast<uint32_t*>(out)[opos + 1] = 0;                // covers the bit_index/8+1 byte (lz77.c:341)
                block_sizes[b] = bits;
                block_bytes[b] = bits / 8 + 1;
            } else {
                block_sizes[b] = opos;
                block_bytes[b] = opos;
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------ offsets (single CTA)
__global__ void __launch_bounds__(1024) lz77_offsets_kernel(const uint64_t* __restrict__ block_bytes, uint64_t nblocks,
                                                            uint64_t* __restrict__ block_off, uint64_t capacity,
                                                            uint64_t* __restrict__ info) {
    __shared__ uint64_t warp_tot[33];
    const uint64_t total = cta_exscan_1024(nblocks, warp_tot,
        [&](uint64_t b) { return block_bytes[b]; }, [&](uint64_t b, uint64_t ex) { block_off[b] = ex; });
    if (threadIdx.x == 0) { block_off[nblocks] = total; info[0] = total; info[1] = total > capacity ? 1 : 0; }
}

// ------------------------------------------------------------------ compaction
// Copies each block's tokens from its fixed-stride scratch slot to its final,
// arbitrarily aligned offset. One CTA per (block, 32 KiB piece). 16-byte aligned
// destination stores assembled from aligned source words with funnel shifts.
__global__ void __launch_bounds__(256) lz77_gather_kernel(const uint8_t* __restrict__ scratch, uint64_t stride,
                                                          const uint64_t* __restrict__ block_bytes, const uint64_t* __restrict__ block_off,
                                                          uint32_t pieces, uint8_t* __restrict__ out, const uint64_t* __restrict__ info) {
    if (info[1]) return;
    const uint64_t b = blockIdx.x / pieces, piece = blockIdx.x % pieces;
    const uint64_t nb = block_bytes[b];
    const uint64_t PIECE = 32768;
    const uint64_t lo = piece * PIECE;
    if (lo >= nb) return;
    const uint64_t hi = lo + PIECE < nb ? lo + PIECE : nb;
    const uint8_t* src = scratch + b * stride + lo;   // 16-byte aligned (stride and PIECE are)
    uint8_t* dst = out + block_off[b] + lo;
    const uint64_t cnt = hi - lo;
    // head: bytes until dst is 16-byte aligned
    uint64_t head = (16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15;
    if (head > cnt) head = cnt;
    if (threadIdx.x < head) dst[threadIdx.x] = src[threadIdx.x];
    const uint64_t body = (cnt - head) / 16;
    const uint32_t sh = (uint32_t)(head & 3) * 8;           // src misalignment of the body in bits
    const uint32_t* s32 = reinterpret_cast<const uint32_t*>(src + (head & ~(uint64_t)3));
    uint4* d128 = reinterpret_cast<uint4*>(dst + head);
    for (uint64_t i = threadIdx.x; i < body; i += 256) {
        const uint32_t* a = s32 + i * 4;
        const uint32_t w0 = a[0], w1 = a[1], w2 = a[2], w3 = a[3];
        const uint32_t w4 = sh ? a[4] : 0;
        uint4 v;
        v.x = __funnelshift_r(w0, w1, sh); v.y = __funnelshift_r(w1, w2, sh);
        v.z = __funnelshift_r(w2, w3, sh); v.w = __funnelshift_r(w3, w4, sh);
        d128[i] = v;
    }
    const uint64_t done = head + body * 16;
    if (threadIdx.x < cnt - done) dst[done + threadIdx.x] = src[done + threadIdx.x];
}

// ------------------------------------------------------------------ decoders
// One HALF-warp (16 lanes) per block, 8 blocks per CTA, no shared memory: up to 128 blocks in
// flight per SM hide the latency of the one dependent L2 read per match (the copy source was
// written by the same lanes a moment ago), and the two halves of a warp have their reads in
// flight together. The token stream is held in a 1024-bit register window (two u32 per lane, the
// second prefetched); runs of up to 16 literal tokens are recognised with one ballot (token k of a
// run sits at a fixed stride) and stored with one coalesced byte store; a match is copied 16 bytes
// per step, overlapping matches (offset < length) resolved by the period rule
// out[o+k] = out[o - off + k % off]. Source bytes are read with ld.cg because they were written by
// other lanes of the same group.
template <int V, uint32_t G>
__global__ void __launch_bounds__(128) lz77_decode_kernel(const uint8_t* __restrict__ stream, const uint64_t* __restrict__ block_off,
                                                         const uint64_t* __restrict__ block_sizes, uint64_t n, uint64_t bs,
                                                         uint64_t nblocks, uint8_t* __restrict__ out) {
    const unsigned lane = threadIdx.x & 31, gl = lane & (G - 1), grp = lane / G;
    const unsigned gmask = (G == 32 ? 0xFFFFFFFFu : ((1u << (G & 31)) - 1u)) << ((G * grp) & 31);
    const uint64_t b = ((uint64_t)blockIdx.x * 4 + (threadIdx.x >> 5)) * (32 / G) + grp;
    if (b >= nblocks) return;                   // the whole group leaves; every sync below names only the group
    const uint32_t len = (uint32_t)(n - b * bs < bs ? n - b * bs : bs);
    const uint8_t* tk = stream + block_off[b];
    uint8_t* gout = out + b * bs;
    const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(tk) & 3);
    const uint32_t* wbase = reinterpret_cast<const uint32_t*>(tk - mis);
    const uint64_t total_bits = (uint64_t)mis * 8 + (V ? block_sizes[b] * 8 : block_sizes[b]);
    const uint32_t nwords = (uint32_t)((total_bits + 31) >> 5);
    constexpr uint32_t STRIDE = V ? 16 : 9, MATCH_BITS = V ? 32 : 19;
    uint32_t wo = 0;                       // word offset of the window start
    uint32_t t = mis * 8;                  // bit cursor relative to the window start
    uint64_t tend = total_bits;            // end of the stream relative to the window start
    uint32_t cur = gl < nwords ? __ldg(wbase + gl) : 0u;
    uint32_t nxt = G + gl < nwords ? __ldg(wbase + G + gl) : 0u;
    uint32_t o = 0;

    // value of `nb` <= 32 bits at window bit position p < 32 * (2G - 1) (may differ per lane)
    auto extract = [&](uint32_t p, uint32_t nb) -> uint32_t {
        const uint32_t i0 = p >> 5, i1 = i0 + 1;
        const uint32_t a0 = __shfl_sync(gmask, cur, i0 & (G - 1), G), a1 = __shfl_sync(gmask, nxt, i0 & (G - 1), G);
        const uint32_t b0 = __shfl_sync(gmask, cur, i1 & (G - 1), G), b1 = __shfl_sync(gmask, nxt, i1 & (G - 1), G);
        const uint32_t lo = i0 < G ? a0 : a1, hi = i1 < G ? b0 : (i1 < 2 * G ? b1 : 0u);
        const uint32_t v = __funnelshift_r(lo, hi, p & 31);
        return nb >= 32 ? v : (v & ((1u << nb) - 1u));
    };

    for (;;) {
        if (V ? (t >= tend) : (o >= len)) break;
        // ---- a run of literal tokens
        const uint32_t p = t + STRIDE * gl;
        const uint32_t v = extract(p, STRIDE);
        const bool lit = (uint64_t)p + STRIDE <= tend && (V ? (v & 0xFFu) == 0u : ((v & 1u) == 0u && o + gl < len));
        const unsigned stop = __ballot_sync(gmask, !lit) >> ((G * grp) & 31);
        const uint32_t run = stop ? (uint32_t)(__ffs(stop) - 1) : G;
        if (gl < run && o + gl < len) gout[o + gl] = (uint8_t)(V ? (v >> 8) : (v >> 1));
        o += run; t += run * STRIDE;
        // ---- then at most one match token
        if (run < G && (V ? (t < tend) : (o < len))) {
            const uint32_t mv = extract(t, MATCH_BITS);
            const bool is_match = V ? ((mv & 0xFFu) != 0u) : ((mv & 1u) != 0u);
            if (is_match) {
                const uint32_t off = V ? ((mv >> 8) & 0xFFFFu) : ((mv >> 1) & 0x3FFFu);
                const uint32_t ml = V ? (mv >> 24) : ((mv >> 15) & 0xFu);
                __syncwarp(gmask);
                if (off != 0 && off <= o) {
                    for (uint32_t k = gl; k < ml && o + k < len; k += G) {
                        const uint8_t c = __ldcg(gout + (o - off + (k % off)));
                        gout[o + k] = c;
                    }
                }
                o += ml; t += MATCH_BITS;
                if (ml == 0 && !V) break;   // corrupt stream: no progress possible
            }
        }
        __syncwarp(gmask);
        if (t >= 32 * G) {                  // first window register consumed: slide by G words
            wo += G; t -= 32 * G; tend -= 32 * G;
            cur = nxt;
            nxt = wo + G + gl < nwords ? __ldg(wbase + wo + G + gl) : 0u;
        }
    }
}

// Deflate-variant decoder, token-parallel: one warp per block, 32 two-byte units of the token stream per
// step, one unit per lane. Tokens are 2 or 4 bytes, so every token starts on a unit; a unit whose first
// byte is non-zero is a match head unless it is the tail (offset high byte, length) of the match before
// it, and a unit is a tail iff the run of "first byte non-zero" units that ends just before it has odd
// length (see deflate_huff.cu) -- one ballot classifies the 32 units. An exclusive scan of the output
// lengths places every token; the step's OUTPUT BYTES are then spread over the lanes (64 per round): a lane
// looks up the unit its byte belongs to and stores the literal or the byte at x - offset. Sources written in an
// earlier step -- nearly all of them in text -- need no waiting; a round that reads the step's own output runs
// behind a __syncwarp and chases sources inside its 32 bytes over shuffles. Source bytes are read with ld.cg
// because other lanes wrote them a moment ago.
template <int MODE>
__global__ void __launch_bounds__(128) lz77_decode_units_kernel(const uint8_t* __restrict__ stream, const uint64_t* __restrict__ block_off,
                                                               const uint64_t* __restrict__ block_sizes, uint64_t n, uint64_t bs,
                                                               uint64_t nblocks, uint8_t* __restrict__ out) {
    const unsigned lane = threadIdx.x & 31;
    const uint64_t b = (uint64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (b >= nblocks) return;
    const uint32_t len = (uint32_t)(n - b * bs < bs ? n - b * bs : bs);
    // three kernels share the blocks: MODE 0 (bytes spread over the lanes) takes the literal-heavy ones, stream >= 1.1 x
    // block (from 1.9 x on: lz77_decode_units64_kernel); MODE 1 (a lane copies its own match) the others (measured on B200, 1 GB: text 5.8 against 6.8 ms, acgt 6.1 against 4.5)
    {
        const uint64_t s10 = block_sizes[b] * 10ull, l = (uint64_t)len;
        if (s10 >= l * 19ull) return;                                     // nearly incompressible: lz77_decode_units64_kernel
        if ((s10 >= l * 11ull) != (MODE == 0)) return;
    }
    const uint16_t* tk = reinterpret_cast<const uint16_t*>(stream + block_off[b]);   // block offsets are even (tokens are 2 or 4 bytes)
    const uint32_t nunits = (uint32_t)(block_sizes[b] >> 1);
    uint8_t* gout = out + b * bs;
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint32_t o = 0;        // output bytes produced by the units before this step
    uint32_t s0 = 0;       // 1 = the first unit of this step is the tail of a match
    // the units of this step and of the next three are held in registers: four loads of the token stream
    // are always in flight (ncu: with one step of look-ahead the kernel waited on this load)
    uint32_t nextv = lane < nunits ? __ldg(tk + lane) : 0u;
    uint32_t pf1 = lane + 32 < nunits ? __ldg(tk + lane + 32) : 0u;
    uint32_t pf2 = lane + 64 < nunits ? __ldg(tk + lane + 64) : 0u;
    uint32_t pf3 = lane + 96 < nunits ? __ldg(tk + lane + 96) : 0u;
    for (uint32_t base = 0; base < nunits; base += 32) {
        const uint32_t u = base + lane;
        const bool valid = u < nunits;
        const uint32_t cur = nextv;
        nextv = pf1; pf1 = pf2; pf2 = pf3;
        pf3 = u + 128 < nunits ? __ldg(tk + u + 128) : 0u;
        uint32_t nxt = __shfl_down_sync(0xffffffffu, cur, 1);
        const uint32_t nv0 = __shfl_sync(0xffffffffu, nextv, 0);
        if (lane == 31) nxt = nv0;
        const bool h = valid && (cur & 0xFFu) != 0u;
        const uint32_t m = __ballot_sync(0xffffffffu, h);
        if (m == 0u && s0 == 0u) {                                         // 32 literals (incompressible stretches): no scan, no copies
            if (valid && o + lane < len) gout[o + lane] = (uint8_t)(cur >> 8);
            o += nunits - base < 32u ? nunits - base : 32u;
            __syncwarp();
            continue;
        }
        // run of h-units that ends just before this lane
        const uint32_t zeros_below = ~m & lt_mask;
        const uint32_t tail = zeros_below ? ((lane - 32u + (uint32_t)__clz((int)zeros_below)) & 1u)   // lane - 1 - (31 - clz)
                                          : (s0 ^ (lane & 1u));
        const uint32_t z32 = ~m;
        s0 = z32 ? ((uint32_t)__clz((int)z32) & 1u) : s0;                  // state of the unit after lane 31: run = 31 - (31 - clz)
        const bool is_lit = valid && !tail && !h, is_head = valid && !tail && h;
        const uint32_t ml = is_head ? (nxt >> 8) & 0xFFu : 0u;
        const uint32_t off = is_head ? ((cur >> 8) | ((nxt & 0xFFu) << 8)) : 0u;
        const uint32_t outlen = is_lit ? 1u : ml;
        const uint32_t incl = warp_incl_scan_u32(outlen);
        const uint32_t myo = o + incl - outlen;
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        const bool copy = is_head && ml != 0u && off != 0u && off <= myo;   // an offset of 0 or beyond the start copies nothing (as lz77_decode_kernel)
        // The step's output bytes, 64 per round and one or two per lane: the byte's unit is the first one whose inclusive
        // length sum lies above it (five shuffles), a literal brings its value along, a match byte at x reads x - offset.
        // Sources below the step's first byte are final and visible (nearly all of them in text): those rounds load and
        // store without waiting. A round that reads this step's own output runs behind a __syncwarp, 32 bytes at a time,
        // and follows sources inside those 32 bytes by pointer doubling over shuffles (period-k runs: at most five rounds).
        const uint32_t pk = off | (is_lit ? 0x10000u : 0u) | (copy ? 0x20000u : 0u) | ((cur >> 8) << 24);
        if (MODE == 0) {
            for (uint32_t r0 = 0; r0 < total; r0 += 64) {
            uint32_t xs[2], sv[2], fl[2];                   // position, source position / literal value, flags: 1 = store, 2 = load first
#pragma unroll
            for (uint32_t q = 0; q < 2; ++q) {
                const uint32_t xr = r0 + 32u * q + lane;
                uint32_t j = 0;
#pragma unroll
                for (uint32_t st = 16; st > 0; st >>= 1) {
                    const uint32_t ic = __shfl_sync(0xffffffffu, incl, (j + st - 1u) & 31u);
                    if (ic <= xr) j += st;
                }
                const uint32_t pj = __shfl_sync(0xffffffffu, pk, j & 31u);
                const uint32_t x = o + xr;
                const bool inb = xr < total && x < len;
                xs[q] = x;
                if (pj & 0x10000u) { sv[q] = pj >> 24; fl[q] = inb ? 1u : 0u; }
                else { sv[q] = x - (pj & 0xFFFFu); fl[q] = (inb && (pj & 0x20000u)) ? 3u : (inb ? 4u : 0u); }   // 4: a match that copies nothing
            }
            const bool dep = ((fl[0] & 2u) && sv[0] >= o) || ((fl[1] & 2u) && sv[1] >= o);
            if (!__any_sync(0xffffffffu, dep)) {
                uint32_t v0 = sv[0], v1 = sv[1];
                if (fl[0] & 2u) v0 = __ldcg(gout + sv[0]);
                if (fl[1] & 2u) v1 = __ldcg(gout + sv[1]);
                if (fl[0] & 1u) gout[xs[0]] = (uint8_t)v0;
                if (fl[1] & 1u) gout[xs[1]] = (uint8_t)v1;
            } else {
#pragma unroll 1
            for (uint32_t q = 0; q < 2; ++q) {
                __syncwarp();                               // the bytes stored so far are visible
                const uint32_t sub = o + r0 + 32u * q;
                uint32_t val = sv[q], ptr = lane;
                bool done = true;
                if (fl[q] & 2u) {
                    if (sv[q] < sub) val = __ldcg(gout + sv[q]);
                    else { ptr = sv[q] - sub; done = false; }
                } else if (fl[q] & 4u) val = __ldcg(gout + xs[q]);     // (its bytes stay what they were)
                while (__any_sync(0xffffffffu, !done)) {
                    const uint32_t v2 = __shfl_sync(0xffffffffu, val, ptr);
                    const uint32_t p2 = __shfl_sync(0xffffffffu, ptr, ptr);
                    const bool d2 = __shfl_sync(0xffffffffu, done ? 1u : 0u, ptr) != 0u;
                    if (!done) { if (d2) { val = v2; done = true; } else ptr = p2; }
                }
                if (fl[q] & 1u) gout[xs[q]] = (uint8_t)val;
            }
            }
            }
        } else {
            // long steps (match-heavy input): a match is copied by its own lane as soon as all of its source bytes lie below
            // the first byte that is not final yet
            if (is_lit && myo < len) gout[myo] = (uint8_t)(cur >> 8);
            bool pending = copy;
            const uint32_t need = copy ? myo - off + (ml < off ? ml : off) : 0u;   // end of the source bytes this match reads
            for (;;) {
                const uint32_t pm = __ballot_sync(0xffffffffu, pending);
                if (!pm) break;
                const uint32_t wm = __shfl_sync(0xffffffffu, myo, __ffs(pm) - 1);  // every byte below the first unfinished match is final
                __syncwarp();                                                      // ... and visible
                if (pending && need <= wm) {
                    const uint8_t* src = gout + (myo - off);
                    uint8_t* dst = gout + myo;
                    const uint32_t cnt = myo + ml <= len ? ml : (myo < len ? len - myo : 0u);
                    if (off >= 8u) {                                               // eight independent reads per batch
                        for (uint32_t k0 = 0; k0 < cnt; k0 += 8) {
                            uint8_t c[8];
#pragma unroll
                            for (uint32_t q = 0; q < 8; ++q) c[q] = k0 + q < cnt ? __ldcg(src + k0 + q) : (uint8_t)0;
#pragma unroll
                            for (uint32_t q = 0; q < 8; ++q) if (k0 + q < cnt) dst[k0 + q] = c[q];
                        }
                    } else {                                                       // short period: out[o+k] = out[o-off + k % off]
                        uint8_t c[8];
#pragma unroll
                        for (uint32_t q = 0; q < 8; ++q) c[q] = q < off ? __ldcg(src + q) : (uint8_t)0;
                        uint32_t si = 0;
                        for (uint32_t k = 0; k < cnt; ++k) {
                            uint8_t v = c[0];
#pragma unroll
                            for (uint32_t q = 1; q < 8; ++q) if (si == q) v = c[q];
                            dst[k] = v;
                            if (++si == off) si = 0;
                        }
                    }
                    pending = false;
                }
                __syncwarp();
            }
        }
        o += total;
        __syncwarp();
    }
}

// The byte-per-lane decoder at 64 units per step (a lane holds unit l of two rows of 32) for nearly incompressible blocks: a step
// of 64 literals is two stores; the rest as lz77_decode_units_kernel<0> with ONE packed scan of both rows' output lengths.
// (Measured on B200, 1 GB: near-random input 2.09 -> 1.58 ms; text is SLOWER this way, 6.3 against 5.8 ms - the longer step is a
// longer chain of dependent shuffles and loads per warp - so text stays with 32 units per step.)
__global__ void __launch_bounds__(128) lz77_decode_units64_kernel(const uint8_t* __restrict__ stream, const uint64_t* __restrict__ block_off,
                                                                 const uint64_t* __restrict__ block_sizes, uint64_t n, uint64_t bs,
                                                                 uint64_t nblocks, uint8_t* __restrict__ out) {
    const unsigned lane = threadIdx.x & 31;
    const uint64_t b = (uint64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (b >= nblocks) return;
    const uint32_t len = (uint32_t)(n - b * bs < bs ? n - b * bs : bs);
    if (block_sizes[b] * 10ull < (uint64_t)len * 19ull) return;          // only nearly incompressible blocks (stream >= 1.9 x block)
    const uint16_t* tk = reinterpret_cast<const uint16_t*>(stream + block_off[b]);
    const uint32_t nunits = (uint32_t)(block_sizes[b] >> 1);
    uint8_t* gout = out + b * bs;
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint32_t o = 0, s0 = 0;
    auto ld = [&](uint32_t u) -> uint32_t { return u < nunits ? (uint32_t)__ldg(tk + u) : 0u; };
    uint32_t a0 = ld(lane), a1 = ld(lane + 32), b0 = ld(lane + 64), b1 = ld(lane + 96), c0 = ld(lane + 128), c1 = ld(lane + 160);
    for (uint32_t base = 0; base < nunits; base += 64) {
        const uint32_t cur0 = a0, cur1 = a1;
        a0 = b0; a1 = b1; b0 = c0; b1 = c1;
        c0 = ld(base + 192 + lane); c1 = ld(base + 224 + lane);
        uint32_t nx0 = __shfl_down_sync(0xffffffffu, cur0, 1), nx1 = __shfl_down_sync(0xffffffffu, cur1, 1);
        const uint32_t f1 = __shfl_sync(0xffffffffu, cur1, 0), f2 = __shfl_sync(0xffffffffu, a0, 0);
        if (lane == 31) { nx0 = f1; nx1 = f2; }
        const bool v0 = base + lane < nunits, v1 = base + 32 + lane < nunits;
        const bool h0 = v0 && (cur0 & 0xFFu) != 0u, h1 = v1 && (cur1 & 0xFFu) != 0u;
        const uint32_t m0 = __ballot_sync(0xffffffffu, h0), m1 = __ballot_sync(0xffffffffu, h1);
        if ((m0 | m1) == 0u && s0 == 0u) {                                   // 64 literals
            if (v0 && o + lane < len) gout[o + lane] = (uint8_t)(cur0 >> 8);
            if (v1 && o + 32 + lane < len) gout[o + 32 + lane] = (uint8_t)(cur1 >> 8);
            o += nunits - base < 64u ? nunits - base : 64u;
            __syncwarp();
            continue;
        }
        // a unit is a match tail iff the run of "first byte non-zero" units that ends just before it is odd
        const uint32_t zb0 = ~m0 & lt_mask, zb1 = ~m1 & lt_mask;
        const uint32_t sA = ~m0 ? ((uint32_t)__clz((int)~m0) & 1u) : s0;     // state behind row 0
        const uint32_t tail0 = zb0 ? ((lane - 32u + (uint32_t)__clz((int)zb0)) & 1u) : (s0 ^ (lane & 1u));
        const uint32_t tail1 = zb1 ? ((lane - 32u + (uint32_t)__clz((int)zb1)) & 1u) : (sA ^ (lane & 1u));
        s0 = ~m1 ? ((uint32_t)__clz((int)~m1) & 1u) : sA;
        const bool lit0 = v0 && !tail0 && !h0, head0 = v0 && !tail0 && h0;
        const bool lit1 = v1 && !tail1 && !h1, head1 = v1 && !tail1 && h1;
        const uint32_t ml0 = head0 ? (nx0 >> 8) & 0xFFu : 0u, ml1 = head1 ? (nx1 >> 8) & 0xFFu : 0u;
        const uint32_t off0 = head0 ? ((cur0 >> 8) | ((nx0 & 0xFFu) << 8)) : 0u, off1 = head1 ? ((cur1 >> 8) | ((nx1 & 0xFFu) << 8)) : 0u;
        const uint32_t ol0 = lit0 ? 1u : ml0, ol1 = lit1 ? 1u : ml1;
        const uint32_t incl = warp_incl_scan_u32(ol0 | (ol1 << 16));         // both rows at once (a row's sum is below 32 * 255)
        const uint32_t tot = __shfl_sync(0xffffffffu, incl, 31);
        const uint32_t total0 = tot & 0xFFFFu, total = total0 + (tot >> 16);
        const uint32_t myo0 = o + (incl & 0xFFFFu) - ol0, myo1 = o + total0 + (incl >> 16) - ol1;
        const bool copy0 = head0 && ml0 != 0u && off0 != 0u && off0 <= myo0, copy1 = head1 && ml1 != 0u && off1 != 0u && off1 <= myo1;
        const uint32_t pk0 = off0 | (lit0 ? 0x10000u : 0u) | (copy0 ? 0x20000u : 0u) | ((cur0 >> 8) << 24);
        const uint32_t pk1 = off1 | (lit1 ? 0x10000u : 0u) | (copy1 ? 0x20000u : 0u) | ((cur1 >> 8) << 24);
        for (uint32_t r0 = 0; r0 < total; r0 += 64) {
            uint32_t xs[2], sv[2], fl[2];
#pragma unroll
            for (uint32_t q = 0; q < 2; ++q) {
                const uint32_t xr = r0 + 32u * q + lane;
                const bool row = xr >= total0;
                const uint32_t key = row ? xr - total0 : xr;
                uint32_t j = 0;
#pragma unroll
                for (uint32_t st = 16; st > 0; st >>= 1) {
                    const uint32_t ic = __shfl_sync(0xffffffffu, incl, (j + st - 1u) & 31u);
                    if ((row ? ic >> 16 : ic & 0xFFFFu) <= key) j += st;
                }
                const uint32_t p0 = __shfl_sync(0xffffffffu, pk0, j & 31u), p1 = __shfl_sync(0xffffffffu, pk1, j & 31u);
                const uint32_t pj = row ? p1 : p0;
                const uint32_t x = o + xr;
                const bool inb = xr < total && x < len;
                xs[q] = x;
                if (pj & 0x10000u) { sv[q] = pj >> 24; fl[q] = inb ? 1u : 0u; }
                else { sv[q] = x - (pj & 0xFFFFu); fl[q] = (inb && (pj & 0x20000u)) ? 3u : (inb ? 4u : 0u); }
            }
            const bool dep = ((fl[0] & 2u) && sv[0] >= o) || ((fl[1] & 2u) && sv[1] >= o);
            if (!__any_sync(0xffffffffu, dep)) {
                uint32_t w0 = sv[0], w1 = sv[1];
                if (fl[0] & 2u) w0 = __ldcg(gout + sv[0]);
                if (fl[1] & 2u) w1 = __ldcg(gout + sv[1]);
                if (fl[0] & 1u) gout[xs[0]] = (uint8_t)w0;
                if (fl[1] & 1u) gout[xs[1]] = (uint8_t)w1;
            } else {
#pragma unroll 1
                for (uint32_t q = 0; q < 2; ++q) {
                    __syncwarp();
                    const uint32_t sub = o + r0 + 32u * q;
                    uint32_t val = sv[q], ptr = lane;
                    bool done = true;
                    if (fl[q] & 2u) {
                        if (sv[q] < sub) val = __ldcg(gout + sv[q]);
                        else { ptr = sv[q] - sub; done = false; }
                    } else if (fl[q] & 4u) val = __ldcg(gout + xs[q]);
                    while (__any_sync(0xffffffffu, !done)) {
                        const uint32_t v2 = __shfl_sync(0xffffffffu, val, ptr);
                        const uint32_t p2 = __shfl_sync(0xffffffffu, ptr, ptr);
                        const bool d2 = __shfl_sync(0xffffffffu, done ? 1u : 0u, ptr) != 0u;
                        if (!done) { if (d2) { val = v2; done = true; } else ptr = p2; }
                    }
                    if (fl[q] & 1u) gout[xs[q]] = (uint8_t)val;
                }
            }
        }
        o += total;
        __syncwarp();
    }
}

inline uint64_t round16(uint64_t x) { return (x + 15) & ~(uint64_t)15; }

}  // namespace

bool lz77_v2_supported(uint64_t bs);
int lz77_v2_launch(b200_ctx* ctx, int variant, const uint8_t* d_in, uint64_t n, uint64_t bs, uint64_t nblocks,
                   uint8_t* scratch, uint64_t stride, uint64_t* d_block_sizes, uint64_t* block_bytes, uint32_t* dbg_tok,
                   const uint32_t* blist = nullptr, const uint32_t* bcount = nullptr);
// deflate variant, blocks of at most 65536 bytes (two expiry phases): cluster-local placement by slot sweeps (lz77_v4.cu);
// blocks it cannot take are listed and go through lz77_v2_launch
int lz77_v4_launch(b200_ctx* ctx, const uint8_t* d_in, uint64_t n, uint64_t bs, uint64_t nblocks,
                   uint8_t* scratch, uint64_t stride, uint64_t* d_block_sizes, uint64_t* block_bytes, uint32_t* dbg_tok, int mode);
// blocks of at most 65536 bytes: occupancy-decided finds + lane-serial simulation of the mixed clusters (lz77_v3.cu)
int lz77_v3_launch(b200_ctx* ctx, int variant, const uint8_t* d_in, uint64_t n, uint64_t bs, uint64_t nblocks,
                   uint8_t* scratch, uint64_t stride, uint64_t* d_block_sizes, uint64_t* block_bytes, uint32_t* dbg_tok);

// token-parallel decoder for any block size (lz77_pdec.cu)
int lz77_pdec_launch(b200_ctx* ctx, int variant, const uint8_t* d_stream, const uint64_t* d_block_off, const uint64_t* d_block_sizes,
                     uint64_t n, uint64_t bs, uint64_t nblocks, uint8_t* d_out);

extern "C" uint64_t b200_lz77_block_stride(uint64_t block_size) { return round16(2 * block_size + 16); }

// scratch slots used: 1 tables, 2 clear queues, 3 token scratch, 4 block_bytes/info/err
static int lz77_encode_impl(b200_ctx* ctx, int variant, const uint8_t* d_in, uint64_t n, uint64_t block_size,
                            uint8_t* d_out, uint64_t out_capacity, uint64_t* d_block_sizes, uint64_t* d_block_off,
                            uint64_t* h_total_bytes, uint32_t* dbg_tok) {
    if (variant != 0 && variant != 1) { B200_SET_ERR("lz77: variant must be 0 or 1"); return B200_ERR_ARG; }
    if (n == 0) {
        CUDA_TRY(cudaMemsetAsync(d_block_off, 0, 2 * sizeof(uint64_t), ctx->stream));
        if (h_total_bytes) *h_total_bytes = 0;
        return B200_OK;
    }
    const uint64_t bs = (block_size == 0 || block_size > n) ? n : block_size;
    if (bs >= 0xFFFFFFF0ull) { B200_SET_ERR("lz77: block of %llu bytes exceeds the 4 GiB limit", (unsigned long long)bs); return B200_ERR_ARG; }
    const bool big = bs > MAX_BLOCK;   // index does not fit the epoch tag: clear the table, one block per warp per launch
    const uint64_t nblocks = (n + bs - 1) / bs;
    const uint64_t stride = b200_lz77_block_stride(bs);

    // v2: whole table in one SM's shared memory (blocks <= 64 KiB); v1: table in HBM (any block <= 4 MiB)
    const bool use_v2 = lz77_v2_supported(bs) && !getenv("B200_LZ_FORCE_V1");
    uint8_t* scratch; uint64_t* misc;
    B200_TRY(b200_scratch(ctx, B200_SLOT(ctx, 3), (size_t)(nblocks * stride + 64), reinterpret_cast<void**>(&scratch)));
    B200_TRY(b200_scratch(ctx, B200_SLOT(ctx, 4), (size_t)(nblocks * 8 + 64), reinterpret_cast<void**>(&misc)));
    uint64_t* info = misc;                       // [0] total, [1] overflow
    uint32_t* err = reinterpret_cast<uint32_t*>(misc + 2);
    uint64_t* block_bytes = misc + 4;
    CUDA_TRY(cudaMemsetAsync(misc, 0, 32, ctx->stream));
    if (use_v2) {
        if (dbg_tok && bs > 65536) { B200_SET_ERR("lz77: token dump needs blocks <= 65536"); return B200_ERR_ARG; }
        B200_TIMED_BEGIN(ctx, B200_K_LZ_PARSE);
        // deflate variant, blocks of at most 64 KiB: B200_LZ_V4=1 lz77_v4_kernel, =0 lz77_v2_kernel, unset: a sample of the input decides
        const char* v4e = getenv("B200_LZ_V4");
        // (below a few dozen blocks the sample kernel and the second launch cost more than the choice can gain)
        uint64_t v4_min_blocks = 64;
        if (const char* e = getenv("B200_LZ_V4_MIN_BLOCKS")) { const long long v = atoll(e); if (v >= 0) v4_min_blocks = (uint64_t)v; }
        const int v4mode = (bs <= 65536 && variant == 1 && !getenv("B200_LZ_V3")) ? (v4e ? (v4e[0] == '1' ? 1 : 0) : (nblocks >= v4_min_blocks ? 2 : 0)) : 0;
        if (v4mode) B200_TRY(lz77_v4_launch(ctx, d_in, n, bs, nblocks, scratch, stride, d_block_sizes, block_bytes, dbg_tok, v4mode));
        else if (bs <= 65536 && getenv("B200_LZ_V3")) B200_TRY(lz77_v3_launch(ctx, variant, d_in, n, bs, nblocks, scratch, stride, d_block_sizes, block_bytes, dbg_tok));
        else B200_TRY(lz77_v2_launch(ctx, variant, d_in, n, bs, nblocks, scratch, stride, d_block_sizes, block_bytes, dbg_tok));
        B200_TIMED_END(ctx);
    } else {
        if (dbg_tok) { B200_SET_ERR("lz77: token dump needs the shared-memory path (block <= 65536)"); return B200_ERR_ARG; }
        int wps = 32;
        if (const char* e = getenv("B200_LZ_WARPS_PER_SM")) { int v = atoi(e); if (v >= 1 && v <= 64) wps = v; }
        uint64_t nwarps = (uint64_t)ctx->sm_count * wps;
        if (nwarps > nblocks) nwarps = nblocks;
        nwarps = (nwarps + 3) / 4 * 4;  // 4 warps per CTA
        if (big && nwarps > 256) nwarps = 256;
        const uint64_t per_warp = (nblocks + nwarps - 1) / nwarps;
        const size_t table_bytes = (size_t)nwarps * (TABLE_SLOTS + GUARD) * sizeof(uint2);
        uint2* tables; uint32_t* clrq;
        const bool fresh = ctx->cap[B200_SLOT(ctx, 1)] < table_bytes;
        B200_TRY(b200_scratch(ctx, B200_SLOT(ctx, 1), table_bytes, reinterpret_cast<void**>(&tables)));
        B200_TRY(b200_scratch(ctx, B200_SLOT(ctx, 2), (size_t)nwarps * CLRQ * 4, reinterpret_cast<void**>(&clrq)));
        uint32_t& ep = ctx->lz_epoch[ctx->bank & 1];   // one table arena (and epoch counter) per bank
        const unsigned grid = (unsigned)(nwarps / 4);
        if (big) {
            // waves of nwarps blocks; the epoch tags of earlier calls are wiped, so restart them
            B200_TIMED_BEGIN(ctx, B200_K_LZ_PARSE);
            for (uint64_t first = 0; first < nblocks; first += nwarps) {
                const uint64_t cnt = nblocks - first < nwarps ? nblocks - first : nwarps;
                CUDA_TRY(cudaMemsetAsync(tables, 0, table_bytes, ctx->stream));
                const uint64_t n_here = (first + cnt) * bs < n ? cnt * bs : n - first * bs;
                if (variant == 0)
                    lz77_parse_kernel<0, true><<<grid, 128, 0, ctx->stream>>>(d_in + first * bs, n_here, bs, cnt, tables, clrq, 0, scratch + first * stride, stride, d_block_sizes + first, block_bytes + first, err);
                else
                    lz77_parse_kernel<1, true><<<grid, 128, 0, ctx->stream>>>(d_in + first * bs, n_here, bs, cnt, tables, clrq, 0, scratch + first * stride, stride, d_block_sizes + first, block_bytes + first, err);
                ctx->launches += 1;
            }
            B200_TIMED_END(ctx);
            ep = MAX_EPOCH;   // forces a clear before the next epoch-tagged call
        } else {
            if (fresh || ep + per_warp > MAX_EPOCH) {
                CUDA_TRY(cudaMemsetAsync(tables, 0, ctx->cap[B200_SLOT(ctx, 1)], ctx->stream));
                ep = 0;
            }
            B200_TIMED_BEGIN(ctx, B200_K_LZ_PARSE);
            if (variant == 0)
                lz77_parse_kernel<0, false><<<grid, 128, 0, ctx->stream>>>(d_in, n, bs, nblocks, tables, clrq, ep, scratch, stride, d_block_sizes, block_bytes, err);
            else
                lz77_parse_kernel<1, false><<<grid, 128, 0, ctx->stream>>>(d_in, n, bs, nblocks, tables, clrq, ep, scratch, stride, d_block_sizes, block_bytes, err);
            B200_TIMED_END(ctx);
            ep += (uint32_t)per_warp;
        }
    }
    lz77_offsets_kernel<<<1, 1024, 0, ctx->stream>>>(block_bytes, nblocks, d_block_off, out_capacity, info);
    const uint32_t pieces = (uint32_t)((stride + 32767) / 32768);
    lz77_gather_kernel<<<(unsigned)(nblocks * pieces), 256, 0, ctx->stream>>>(scratch, stride, block_bytes, d_block_off, pieces, d_out, info);
    ctx->launches += 3;
    CUDA_TRY(cudaGetLastError());
    if (h_total_bytes) {
        uint64_t* pin; B200_TRY(b200_pinned(ctx, 64, reinterpret_cast<void**>(&pin)));
        CUDA_TRY(cudaMemcpyAsync(pin, misc, 32, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        *h_total_bytes = pin[0];
        if (pin[1]) { B200_SET_ERR("lz77: stream needs %llu bytes, capacity %llu", (unsigned long long)pin[0], (unsigned long long)out_capacity); return B200_ERR_CAPACITY; }
        if (reinterpret_cast<uint32_t*>(pin + 2)[0]) { B200_SET_ERR("lz77: slot-0 clear queue overflow"); return B200_ERR_DOMAIN; }
    }
    return B200_OK;
}

extern "C" int b200_lz77_encode_dev(b200_ctx* ctx, int variant, const uint8_t* d_in, uint64_t n, uint64_t block_size,
                                    uint8_t* d_out, uint64_t out_capacity, uint64_t* d_block_sizes, uint64_t* d_block_off,
                                    uint64_t* h_total_bytes) {
    B200_ENTER(ctx);
    return lz77_encode_impl(ctx, variant, d_in, n, block_size, d_out, out_capacity, d_block_sizes, d_block_off, h_total_bytes, nullptr);
}

// Test hook: same as b200_lz77_encode_dev, additionally dumping the per-position token
// candidates of the match finder (0 = literal, else offset | len << 16) for every block,
// 65536 entries per block. Only for blocks <= 65536 bytes.
extern "C" int b200_lz77_encode_debug_dev(b200_ctx* ctx, int variant, const uint8_t* d_in, uint64_t n, uint64_t block_size,
                                          uint8_t* d_out, uint64_t out_capacity, uint64_t* d_block_sizes, uint64_t* d_block_off,
                                          uint64_t* h_total_bytes, uint32_t* d_tok) {
    B200_ENTER(ctx);
    return lz77_encode_impl(ctx, variant, d_in, n, block_size, d_out, out_capacity, d_block_sizes, d_block_off, h_total_bytes, d_tok);
}

extern "C" int b200_lz77_decode_dev(b200_ctx* ctx, int variant, const uint8_t* d_stream, const uint64_t* d_block_off,
                                    const uint64_t* d_block_sizes, uint64_t n, uint64_t block_size, uint8_t* d_out) {
    B200_ENTER(ctx);
    if (variant != 0 && variant != 1) { B200_SET_ERR("lz77: variant must be 0 or 1"); return B200_ERR_ARG; }
    if (n == 0) return B200_OK;
    const uint64_t bs = (block_size == 0 || block_size > n) ? n : block_size;
    const uint64_t nblocks = (n + bs - 1) / bs;
    // lanes per block: a whole warp while all blocks are resident at once (shortest chain per block),
    // a half-warp beyond that (twice the blocks in flight instead of a second wave); measured on B200:
    // 300 MB 23.5 vs 27.4 ms/GB, 1 GB 18.9 vs 15.2 ms/GB
    uint32_t G = nblocks * 32 > (uint64_t)ctx->sm_count * 2048 ? 16u : 32u;
    if (const char* e = getenv("B200_LZ_DEC_G")) { const int v = atoi(e); if (v == 8 || v == 16 || v == 32) G = (uint32_t)v; }
    const char* seq = getenv("B200_LZ_DEC_SERIAL");     // keep the token-serial decoder reachable for comparison
    {
        // one warp per block starves the GPU when blocks are few and large (a 4 MiB block decoded at 0.5 GB/s, the whole
        // buffer as one block at 15 MB/s): from PDEC_MIN_BLOCK bytes per block on, every kernel works on the whole stream
        const char* pd = getenv("B200_LZ_PDEC");        // 0 / 1: never / always
        uint64_t min_block = variant ? (1ull << 20) : (1ull << 17);
        if (const char* e = getenv("B200_LZ_PDEC_MIN_BLOCK")) { const long long v = atoll(e); if (v > 0) min_block = (uint64_t)v; }
        // ... and below that size when the blocks are too few to give every SM a handful of warps (measured on B200, 32 MiB:
        // 256 KiB blocks of the deflate variant 7.8 vs 14.9 GB/s, 64 KiB blocks of the standalone variant 8.5 vs 15.8 GB/s)
        const bool few = nblocks < 4ull * (uint64_t)ctx->sm_count && bs >= (variant ? (1ull << 18) : (1ull << 12));
        const bool use_pdec = pd ? pd[0] == '1' : ((bs >= min_block || few) && bs < (1ull << 30) && !(seq && seq[0] == '1'));
        if (use_pdec) {
            B200_TIMED_BEGIN(ctx, B200_K_LZ_DECODE);
            const int rc = lz77_pdec_launch(ctx, variant, d_stream, d_block_off, d_block_sizes, n, bs, nblocks, d_out);
            B200_TIMED_END(ctx);
            return rc;
        }
    }
    if (variant == 1 && !(seq && seq[0] == '1')) {
        B200_TIMED_BEGIN(ctx, B200_K_LZ_DECODE);
        lz77_decode_units_kernel<0><<<(unsigned)((nblocks + 3) / 4), 128, 0, ctx->stream>>>(d_stream, d_block_off, d_block_sizes, n, bs, nblocks, d_out);
        lz77_decode_units64_kernel<<<(unsigned)((nblocks + 3) / 4), 128, 0, ctx->stream>>>(d_stream, d_block_off, d_block_sizes, n, bs, nblocks, d_out);
        lz77_decode_units_kernel<1><<<(unsigned)((nblocks + 3) / 4), 128, 0, ctx->stream>>>(d_stream, d_block_off, d_block_sizes, n, bs, nblocks, d_out);
        B200_TIMED_END(ctx);
        ctx->launches += 3;
        CUDA_TRY(cudaGetLastError());
        return B200_OK;
    }
    const unsigned grid = (unsigned)((nblocks * G + 127) / 128);   // 128 / G blocks per CTA
    B200_TIMED_BEGIN(ctx, B200_K_LZ_DECODE);
#define LZ_DEC(V, GG) lz77_decode_kernel<V, GG><<<grid, 128, 0, ctx->stream>>>(d_stream, d_block_off, d_block_sizes, n, bs, nblocks, d_out)
    if (variant == 0) { if (G == 8) LZ_DEC(0, 8); else if (G == 16) LZ_DEC(0, 16); else LZ_DEC(0, 32); }
    else { if (G == 8) LZ_DEC(1, 8); else if (G == 16) LZ_DEC(1, 16); else LZ_DEC(1, 32); }
    B200_TIMED_END(ctx);
    ctx->launches += 1;
    CUDA_TRY(cudaGetLastError());
    return B200_OK;
}
