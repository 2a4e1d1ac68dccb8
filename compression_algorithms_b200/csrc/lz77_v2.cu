// LZ77 v2: the reference's 2^20-slot hash table emulated slot-exactly inside ONE SM's
// shared memory, for blocks of up to 65536 bytes (the reference's BUFFER_SIZE,
// algorithms/deflate/deflate.h:8).
//
// Why this is exact (DESIGN.md "LZ77 v2" has the full argument):
//  * every position is inserted exactly once, in order (lz77.c:295,333-336;
//    deflate/lz77.c:228,267-270) and find is read-only, so F(p) = find(word(p)) after
//    inserts 0..p-1 is a pure function of the data; the parse only selects which F(p)
//    are used.
//  * with expiry, the set of live slots is always a subset of the slots occupied when
//    nothing ever expires, and that set does not depend on insertion order. So the
//    no-expiry occupancy bitmap (built with parallel atomicOr linear probing) bounds
//    every probe walk; rank(slot) in that bitmap is a dense "compact slot" index, and the
//    whole table becomes u16 T[n] (key = position+1, 0 = never used) = 128 KiB.
//  * maximal runs of occupied slots ("clusters") never interact, so the slot space is
//    cut into 32 ranges at cluster ends and each warp simulates one range over its own
//    time-ordered list of positions.
//  * inside a warp 32 consecutive list entries walk the table speculatively; a lane keeps
//    a cursor (all slots before it proven live at its time), re-validates it after every
//    commit round and commits once no lower uncommitted lane shares its cursor. That is
//    the sequential semantics, because a placement only ever turns a dead slot live.
//  * the cluster touching slot 0 (and, for the wrapping deflate insert, the one touching
//    slot 2^20-1) is simulated serially with the reference's early slot-0 clear
//    (lz77.c:70-85, U10).
//
// Phases per block (one persistent CTA of 1024 threads per SM):
//   P0 load block -> smem | P1 occupancy bitmap | P2 rank prefix, range cuts
//   P3 compact index + stable partition into 32 per-range lists (global, L2 resident)
//   P4 table simulation -> tok[p] = 0 (literal) | offset | len<<16   (global, L2 resident)
//   P5 greedy parse: per-64-position chunk exit functions (backward DP), entry offsets
//   P6 token emission into the block's scratch slot (then lz77_offsets/gather as in v1)
#include "common.cuh"
#include "../../include/b200comp.h"

namespace {

constexpr uint32_t SLOTS = 1u << 20;
constexpr uint32_t GUARD_BITS = 65536u;
constexpr uint32_t BM_WORDS = (SLOTS + GUARD_BITS) / 32;     // 34816
constexpr uint32_t PRE_CHUNK = 8;                            // words per rank-prefix entry
constexpr uint32_t PRE_N = BM_WORDS / PRE_CHUNK;             // 4352
constexpr uint32_t NONE = 0xFFFFFFFFu;
constexpr uint32_t LONER = 0xFFFFFFFFu;                      // P3 marker: entry alone in its cluster (not a valid hash | range << 21)
constexpr uint32_t MAXB = 65536;                             // max block bytes on this path
constexpr uint32_t NTHREADS = 1024;
constexpr uint32_t CARRY_BYTES = (MAXB + 64) * 4 + 32768 * 4 + MAXB * 4 + 32768 * 4;
constexpr uint32_t NR = 32;                                  // ranges == warps
constexpr uint32_t MAXF = 16;                                // forced cuts (two per hot chain) at most
constexpr uint32_t CROWD_MIN = 12;                           // same-home entries in a batch from which the slot-ordered placement is used
constexpr uint32_t FULL_WORD_WEIGHT = 6;                      // extra cost units per slot of a fully occupied bitmap word (range balancing, P2)
constexpr uint32_t RUN8_WEIGHT = 4;                           // extra cost units per slot beyond the seventh of a run of occupied slots
constexpr uint32_t HOT_MIN_WORDS = 7;                         // fully occupied words (of 8) in a 256-slot chunk from which its chain counts as hot (swept 2..8 on B200)

// shared memory layout (bytes)
constexpr uint32_t OFF_DATA = 0;
constexpr uint32_t SZ_DATA = MAXB + 128;                     // zero pad behind the block (U1)
constexpr uint32_t OFF_BIG = OFF_DATA + SZ_DATA;
constexpr uint32_t SZ_BIG = BM_WORDS * 4;                    // 139264: bitmap | T+B1 | adv+exit | staging
constexpr uint32_t OFF_PRE = OFF_BIG + SZ_BIG;               // must directly follow BIG (staging may spill)
constexpr uint32_t SZ_PRE = (PRE_N + 4) * 4;
constexpr uint32_t OFF_MISC = OFF_PRE + SZ_PRE;
constexpr uint32_t SZ_MISC = 8192;
constexpr uint32_t SMEM_BYTES = OFF_MISC + SZ_MISC;          // 230,544 <= 232,448

constexpr uint32_t T_ENTRIES = MAXB + 64;
constexpr uint32_t OFF_T = 0;                                // inside BIG
constexpr uint32_t OFF_B1 = T_ENTRIES * 2;                   // inside BIG, u16[2052]
// adv / exitof are indexed through PADX: every 64-position chunk is followed by 4 pad bytes, so that the
// phases in which thread t walks chunk t (stride 68 bytes = 17 words) are free of bank conflicts
constexpr uint32_t PADDED = MAXB + (MAXB >> 6) * 4;          // 69632
constexpr uint32_t OFF_ADV = 0;                              // inside BIG, u8[PADDED]
constexpr uint32_t OFF_EXIT = PADDED;                        // inside BIG, u8[PADDED]: 2 * 69632 = 139264 = SZ_BIG
constexpr uint32_t OFF_STAGE = PADDED;                       // inside BIG (after adv), u32 words, V0 only (may spill ~4 KB into PRE)
#define PADX(p) ((p) + (((p) >> 6) << 2))

struct Misc {
    uint32_t cut[NR + 1];
    uint32_t top_start, sp_lo_end, sp_hi_start, pad0;
    uint32_t rstart[NR + 1];
    uint32_t fcut[MAXF], nfcut;  // cuts forced just before / after a hot chain
    uint32_t ncar, cq_h, cq_t, ecarry;   // slices of a large block: live carried entries, clear-queue cursors, parse carry
    uint32_t cnt[NR][NR];        // [warp][range]
    uint32_t clr[64];            // slot-0 clear times (warp 0)
    uint8_t  sexit[NR][32];      // super-chunk exit functions
    uint8_t  sentry[NR + 1];
    uint32_t scan[34];
    uint32_t entry_pad;
    uint32_t p1_next;                            // P1: next row of 32 positions to hand out
    uint8_t  rtab[(SLOTS + GUARD_BITS) >> 12];   // range of the first slot of every 4096-slot bin (P3 starts its search there)
};
static_assert(sizeof(Misc) <= SZ_MISC, "misc region too small");

template <int V> struct Cfg;
template <> struct Cfg<0> { static constexpr uint32_t W = 1u << 14, MAXLEN = 15; };
template <> struct Cfg<1> { static constexpr uint32_t W = 1u << 15, MAXLEN = 31; };

__device__ __forceinline__ uint32_t sm_word(const uint8_t* data, uint32_t p) {
    // unaligned little-endian 4-byte read from shared memory (pad behind the block is zero)
    const uint32_t* a = reinterpret_cast<const uint32_t*>(data + (p & ~3u));
    return __funnelshift_r(a[0], a[1], (p & 3u) * 8);
}

// length of the common prefix of data[m..] and data[q..], at least 4 (the hashed word), capped
// at MAXLEN: compares 4 bytes per step (lz77.c:302-311, deflate/lz77.c:238-247)
template <uint32_t MAXLEN>
__device__ __forceinline__ uint32_t match_len(const uint8_t* data, uint32_t m, uint32_t q) {
    uint32_t l = 4;
#pragma unroll 1
    while (l < MAXLEN) {
        const uint32_t x = sm_word(data, m + l) ^ sm_word(data, q + l);
        if (x) { l += (uint32_t)(__ffs(x) - 1) >> 3; break; }
        l += 4;
    }
    return l < MAXLEN ? l : MAXLEN;
}

__device__ __forceinline__ uint32_t bm_rank(const uint32_t* bm, const uint16_t* pre16, uint32_t s) {
    // set bits below slot s: prefix of the 4-word half chunk (u16: a slice has at most 65536 occupied slots, and the
    // prefix of an occupied slot is below that) + masked popcounts of the half chunk (one 16-byte read)
    const uint32_t wi = s >> 5, j = wi & 3u;
    const uint4 q = *reinterpret_cast<const uint4*>(bm + (wi & ~3u));
    const uint32_t wd[4] = {q.x, q.y, q.z, q.w};
    uint32_t r = pre16[wi >> 2];
    const uint32_t below = (1u << (s & 31)) - 1u;
#pragma unroll
    for (uint32_t k = 0; k < 4; ++k) r += __popc(wd[k] & (k < j ? 0xFFFFFFFFu : (k == j ? below : 0u)));
    return r;
}

// Long walks (hot chains). From cursor kk (every slot before it proven live at the lane's time)
// to the first slot with T <= dthr, one aligned group of 32 slots per step: the 64 bytes of the
// group are read with four 16-byte loads, the dead slots found with packed u16 compares. Whole
// groups are skipped while the group's lower bound B1[g] proves every slot live and, for a lane
// whose find is still pending, the signature filter S1[g] proves the pattern absent. Every group
// that is read gets its exact minimum published to B1 (T only grows, so any snapshot minimum stays
// a valid lower bound). Slots of a neighbouring range may be in the group; the walk ends inside
// the lane's own cluster, so they never decide anything.
__device__ __forceinline__ uint32_t walk_groups(const uint16_t* T, uint16_t* B1, const uint32_t* S1, const uint8_t* data,
                                                uint32_t kk, uint32_t dthr, uint32_t w, uint32_t sig, bool& pend, uint32_t& fm,
                                                uint32_t& iters) {
    const uint32_t thr2 = dthr | (dthr << 16);
    for (;;) {
        uint32_t g = kk >> 5;
        const uint32_t j0 = kk & 31u;
        if (j0 == 0) {
            while (B1[g] > dthr && !(pend && ((S1[g] >> sig) & 1u))) ++g;   // unused tail groups have B1 == 0
            kk = g << 5;
        }
        ++iters;
        const uint4* gp = reinterpret_cast<const uint4*>(T + (g << 5));
        const uint4 q0 = gp[0], q1 = gp[1], q2 = gp[2], q3 = gp[3];
        const uint32_t wd[16] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, q3.x, q3.y, q3.z, q3.w};
        // dead flags of the 32 slots, two instructions per word: acc gets the flag of the even slot 2i in bit
        // 15-i and that of the odd slot 2i+1 in bit 31-i (the walk only needs the first dead slot at/after j0)
        uint32_t acc = 0, mn2 = 0xFFFFFFFFu;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            acc = (acc << 1) + __vsetleu2(wd[i], thr2);             // vset: 1 per halfword that is <= dthr
            mn2 = __vminu2(mn2, wd[i]);
        }
        B1[g] = (uint16_t)min(mn2 & 0xFFFFu, mn2 >> 16);
        const uint32_t ev = (acc & 0xFFFFu) & ((1u << (16u - ((j0 + 1u) >> 1))) - 1u);   // even slots 2i >= j0
        const uint32_t od = (acc >> 16) & ((1u << (16u - (j0 >> 1))) - 1u);              // odd slots 2i+1 >= j0
        const uint32_t s_ev = ev ? 2u * ((uint32_t)__clz(ev) - 16u) : 64u;
        const uint32_t s_od = od ? 2u * ((uint32_t)__clz(od) - 16u) + 1u : 64u;
        const bool dm = (ev | od) != 0u;
        const uint32_t stop = dm ? min(s_ev, s_od) : 32u;
        if (pend) {
#pragma unroll 1
            for (uint32_t j = j0; j < stop; ++j) {
                const uint32_t v = T[(g << 5) + j];
                if (sm_word(data, v - 1) == w) { fm = v - 1; pend = false; break; }
            }
        }
        if (dm) return (g << 5) + stop;
        kk = (g + 1) << 5;
    }
}

// DBG: phase stamps and per-warp cycle counters (tools/lz_stats.py); the clock reads cost a few
// per cent, so the production instantiation leaves them out
#define CLK() (DBG ? clock64() : 0ll)
template <int V, bool DBG>
__global__ void __launch_bounds__(NTHREADS, 1) lz77_v2_kernel(const uint8_t* __restrict__ in, uint64_t n, uint32_t bs, uint32_t nblocks,
                                                             uint32_t* __restrict__ lists_all, uint32_t* __restrict__ tok_all,
                                                             uint8_t* __restrict__ carry_all,
                                                             uint8_t* __restrict__ scratch, uint64_t stride,
                                                             uint64_t* __restrict__ block_sizes, uint64_t* __restrict__ block_bytes,
                                                             uint32_t* __restrict__ dbg_tok,
                                                             const uint32_t* __restrict__ blist, const uint32_t* __restrict__ bcount) {
    constexpr uint32_t W = Cfg<V>::W, MAXLEN = Cfg<V>::MAXLEN;
    extern __shared__ __align__(16) uint8_t smem[];
    uint8_t* data = smem + OFF_DATA;
    uint8_t* big = smem + OFF_BIG;
    uint32_t* bm = reinterpret_cast<uint32_t*>(big);
    uint32_t* pre = reinterpret_cast<uint32_t*>(smem + OFF_PRE);
    uint16_t* pre16 = reinterpret_cast<uint16_t*>(smem + OFF_PRE);   // P2/P3: rank prefix per 4-word half chunk
    Misc* ms = reinterpret_cast<Misc*>(smem + OFF_MISC);
    uint16_t* T = reinterpret_cast<uint16_t*>(big + OFF_T);
    uint16_t* B1 = reinterpret_cast<uint16_t*>(big + OFF_B1);
    uint8_t* adv = big + OFF_ADV;
    uint8_t* exitof = big + OFF_EXIT;
    uint32_t* stage = reinterpret_cast<uint32_t*>(big + OFF_STAGE);

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint32_t* lists = lists_all + (uint64_t)blockIdx.x * MAXB;
    uint32_t* tokb = tok_all + (uint64_t)blockIdx.x * MAXB;
    // state handed from one slice of a large block to the next (global, per CTA)
    uint8_t* carry = carry_all + (uint64_t)blockIdx.x * CARRY_BYTES;
    uint32_t* rawmap = reinterpret_cast<uint32_t*>(carry);                        // [T_ENTRIES] raw slot of every compact slot
    uint32_t* raw_carry = rawmap + T_ENTRIES;                                     // [32768] raw slot of every carried position, NONE = cleared
    uint32_t* cs = raw_carry + 32768;                                             // [MAXB] compact slot every new position was placed in (NONE = wiped); 65535 is a valid slot, hence 32 bits
    uint32_t* ccar = cs + MAXB;                                                   // [32768] compact slot of every carried position (NONE = none)

    uint32_t* dbg_stats = dbg_tok ? dbg_tok + (uint64_t)nblocks * MAXB : nullptr;   // [block][8 phase stamps + 32 warps x 4]
    // blist != nullptr: only the blocks the v4 kernel (lz77_v4.cu) handed back, listed in blist[0 .. *bcount)
    const uint32_t nb_eff = blist ? *bcount : nblocks;
    for (uint32_t bi = blockIdx.x; bi < nb_eff; bi += gridDim.x) {
        const uint32_t b = blist ? blist[bi] : bi;
        const long long t_begin = CLK();
#define PHASE_STAMP(k) do { if (DBG && dbg_stats && tid == 0) dbg_stats[(uint64_t)b * 136 + (k)] = (uint32_t)(clock64() - t_begin); } while (0)
        const uint8_t* bsrc = in + (uint64_t)b * bs;
        const uint32_t blen = (uint32_t)(n - (uint64_t)b * bs < bs ? n - (uint64_t)b * bs : bs);
        // A block larger than 65536 positions is simulated in SLICES so that positions stay 16-bit: a slice
        // holds the W positions before it (the "carried" entries: still live, already placed, their raw slots
        // handed over by the previous slice) and up to 65536 - W new positions. Everything below works on
        // positions relative to the first carried one; a 64 KiB block is a single slice with nothing carried.
        uint32_t n0 = 0;            // first new position of the slice (absolute in the block)
        uint64_t out_units = 0;     // bytes (V1) / bits (V0) of the block's tokens written by earlier slices
        for (;;) {
        const uint32_t C = n0 ? W : 0u;                       // carried positions: relative [0, C)
        const uint32_t B0 = n0 - C;                           // absolute position of relative 0
        const uint32_t avail = blen - B0;                     // bytes of the block from there on
        const uint32_t len = avail < MAXB ? avail : MAXB;     // new positions: relative [C, len)
        const bool last_slice = B0 + len >= blen;
        const uint8_t* src = bsrc + B0;
        const uint32_t p_start = C + (C ? ms->ecarry : 0u);   // where the greedy parse enters the slice (read before P5 rewrites it)

        // ---------------- P0: slice -> shared memory (with the real bytes behind it, zero pad behind the block), zero bitmap
        {
            const bool al = (reinterpret_cast<uintptr_t>(src) & 15) == 0;
            for (uint32_t i = tid * 16; i < len + 128 && i < SZ_DATA; i += NTHREADS * 16) {
                if (al && i + 16 <= avail) *reinterpret_cast<uint4*>(data + i) = __ldg(reinterpret_cast<const uint4*>(src + i));
                else for (uint32_t k = 0; k < 16 && i + k < SZ_DATA; ++k) data[i + k] = (i + k < avail) ? __ldg(src + i + k) : 0;
            }
            for (uint32_t i = tid; i < BM_WORDS; i += NTHREADS) bm[i] = 0;
            for (uint32_t i = tid; i < BM_WORDS / 32; i += NTHREADS) pre[i] = 0;   // P1's "word is full" summary (pre is free until P2)
            if (tid == 0) { ms->ncar = 0; ms->p1_next = C; }
        }
        __syncthreads();
        if (C) {   // the carried entries sit where the previous slice left them
            uint32_t cnt = 0;
            for (uint32_t i = tid; i < C; i += NTHREADS) {
                const uint32_t r = raw_carry[i];
                if (r != NONE) { atomicOr(&bm[r >> 5], 1u << (r & 31)); ++cnt; }
            }
            cnt = warp_sum_u32(cnt);
            if (lane == 0 && cnt) atomicAdd(&ms->ncar, cnt);
            __syncthreads();
        }
        const uint32_t nslots = ms->ncar + (len - C);         // occupied compact slots of the slice

        PHASE_STAMP(0);
        // ---------------- P1: no-expiry occupancy (order independent) by atomic linear probing
        // (a per-home "frontier hint" table that lets later walks skip the full words of a hot chain was
        // tried and did not pay: 175 K -> 190 K cycles; the phase is bound by the latency of the
        // dependent hash -> read -> atomicOr chain of the 64 positions each thread owns)
        // The k-th occurrence of a hot 4-gram would walk k/32 words of its chain; a summary bit per bitmap
        // word ("known to be full", set by whoever takes its last free bit; bits are never cleared in this
        // phase, so a set summary bit is always right) lets a walk jump to the first word that may have room.
        uint32_t* summ = pre;
        for (;;) {
            // 32 positions at a time from a shared counter: the warps that hold many occurrences of a hot 4-gram
            // (long probe walks, lost races) take fewer rows, and all warps reach the barrier together
            uint32_t row = 0;
            if (lane == 0) row = atomicAdd(&ms->p1_next, 32u);
            row = __shfl_sync(0xffffffffu, row, 0);
            if (row >= len) break;
            const uint32_t i = row + lane;
            if (i >= len) continue;
            uint32_t s = lz_hash(sm_word(data, i));
            tokb[i] = s;                                   // P3 reads the hash back instead of computing it again
            for (;;) {
                const uint32_t wi = s >> 5;
                const uint32_t free_bits = ~bm[wi] & (0xFFFFFFFFu << (s & 31));
                if (free_bits) {
                    const uint32_t bit = 1u << (__ffs(free_bits) - 1);
                    const uint32_t old = atomicOr(&bm[wi], bit);
                    if (!(old & bit)) {
                        if ((old | bit) == 0xFFFFFFFFu) atomicOr(&summ[wi >> 5], 1u << (wi & 31));
                        break;
                    }
                    continue;  // lost the race for that bit: look again in the same word
                }
                uint32_t nw = wi + 1;
                uint32_t open = ~summ[nw >> 5] & (0xFFFFFFFFu << (nw & 31));
                while (!open) { nw = ((nw >> 5) + 1) << 5; open = ~summ[nw >> 5]; }   // the guard words behind the table never fill up
                s = ((nw & ~31u) + (uint32_t)(__ffs(open) - 1)) << 5;
                if (V == 1 && s >= SLOTS) s = 0;   // the deflate insert wraps (deflate/lz77.c:99-101)
            }
        }
        __syncthreads();

        PHASE_STAMP(1);
        // ---------------- P2: rank prefix per 8-word chunk, range cuts, special cluster bounds
        {
            // thread t owns chunks [5t, 5t+5)
            uint32_t c0 = tid * 5, mine = 0, mine_cost = 0;
            uint32_t part[5], half[5], cost[5], run8[5], fullw[7];   // fullw[j]: fully occupied words of chunk c0 - 1 + j; half: first four words
#pragma unroll
            for (int j = 0; j < 7; ++j) {
                const uint32_t ch = c0 + j - 1;     // wraps for c0 == 0, j == 0: out of range, counts 0
                uint32_t s = 0, s4 = 0, full = 0, r8 = 0;
                if (ch < PRE_N) for (uint32_t w = 0; w < PRE_CHUNK; ++w) {
                    const uint32_t x = bm[ch * PRE_CHUNK + w];
                    s += __popc(x); if (w == 3) s4 = s; full += x == 0xFFFFFFFFu;
                    uint32_t y = x & (x >> 1); y &= y >> 2; y &= y >> 4;   // slots that end a run of eight or more (inside the word)
                    r8 += __popc(y);
                }
                fullw[j] = full;
                if (j >= 1 && j <= 5) { part[j - 1] = s; half[j - 1] = s4; run8[j - 1] = r8; mine += s; }
            }
            // Range cuts. A chunk with HOT_MIN_WORDS fully occupied bitmap words belongs to the chain of a hot
            // 4-gram: such a chain gets a cut just before and just after it, so that it sits (almost) alone in a
            // warp whose batches are then "crowded" and take the slot-ordered placement of P4. The remaining
            // cuts give the other ranges equal numbers of ordinary entries. Every cut is moved up to the next
            // cluster end (clusters never interact).
            auto free_at_or_after = [&](uint32_t sl) {
                for (;;) {
                    const uint32_t z = ~bm[sl >> 5] & (0xFFFFFFFFu << (sl & 31));
                    if (z) return (sl & ~31u) + (uint32_t)(__ffs(z) - 1);
                    sl = (sl & ~31u) + 32;
                    if (sl >= SLOTS + GUARD_BITS) return sl;
                }
            };
            if (tid == 0) ms->nfcut = 0;
            if (tid >= 1 && tid < NR) ms->cut[tid] = SLOTS + GUARD_BITS;   // a cut nobody sets leaves an empty range
            __syncthreads();
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                const bool hot = fullw[k + 1] >= HOT_MIN_WORDS, hot_prev = fullw[k] >= HOT_MIN_WORDS, hot_next = fullw[k + 2] >= HOT_MIN_WORDS;
                // An entry of a medium-hot chain (a run of occupied slots too short to be cut out, 32 .. 223 slots)
                // costs a whole commit round of its warp, an ordinary entry a thirty-second of a batch: slots in
                // fully occupied bitmap words are weighted accordingly
                cost[k] = (c0 + k < PRE_N && !hot) ? part[k] + RUN8_WEIGHT * run8[k] + FULL_WORD_WEIGHT * 32u * fullw[k + 1] : 0u;
                mine_cost += cost[k];
                if (hot && c0 + k < PRE_N) {
                    if (!hot_prev && c0 + k >= 1) { const uint32_t i = atomicAdd(&ms->nfcut, 1u); if (i < MAXF) ms->fcut[i] = free_at_or_after((c0 + k - 1) * PRE_CHUNK * 32); }
                    if (!hot_next && c0 + k + 1 < PRE_N) { const uint32_t i = atomicAdd(&ms->nfcut, 1u); if (i < MAXF) ms->fcut[i] = free_at_or_after((c0 + k + 1) * PRE_CHUNK * 32); }
                }
            }
            const uint32_t incl = warp_incl_scan_u32(mine);
            if (lane == 31) ms->scan[warp] = incl;
            __syncthreads();
            if (warp == 0) {
                const uint32_t t = ms->scan[lane];
                const uint32_t ti = warp_incl_scan_u32(t);
                ms->scan[lane] = ti - t;
            }
            __syncthreads();
            uint32_t run = ms->scan[warp] + incl - mine;
#pragma unroll
            for (int k = 0; k < 5; ++k) { if (c0 + k < PRE_N) { pre16[2 * (c0 + k)] = (uint16_t)run; pre16[2 * (c0 + k) + 1] = (uint16_t)(run + half[k]); } run += part[k]; }
            const uint32_t cincl = warp_incl_scan_u32(mine_cost);
            __syncthreads();
            if (lane == 31) ms->scan[warp] = cincl;
            __syncthreads();
            if (warp == 0) {
                const uint32_t t = ms->scan[lane];
                const uint32_t ti = warp_incl_scan_u32(t);
                ms->scan[lane] = ti - t;
                if (lane == 31) ms->scan[32] = ti;
            }
            __syncthreads();
            const uint32_t nforced = ms->nfcut < MAXF ? ms->nfcut : MAXF;
            const uint32_t nq = NR - nforced;                                    // ranges shared out by entry count
            {
                const uint64_t ctot = ms->scan[32] ? ms->scan[32] : 1;
                uint64_t crun = (uint64_t)ms->scan[warp] + cincl - mine_cost;
#pragma unroll 1
                for (int k = 0; k < 5; ++k) {
                    if (cost[k]) {
                        uint32_t r = (uint32_t)((crun * nq + ctot - 1) / ctot);      // first quantile at/after this chunk's start
                        if (r == 0) r = 1;
                        while (r < nq && ctot * r < (crun + cost[k]) * nq) { ms->cut[r] = free_at_or_after((c0 + k) * PRE_CHUNK * 32); ++r; }
                    }
                    crun += cost[k];
                }
            }
            if (tid < nforced) ms->cut[nq + tid] = ms->fcut[tid];
            __syncthreads();
            {                                                                    // 31 cuts from two sources: rank-sort them (one thread each)
                uint32_t v = 0, rk = 0;
                if (tid >= 1 && tid < NR) {
                    v = ms->cut[tid];
                    rk = 1;
                    for (uint32_t j = 1; j < NR; ++j) { const uint32_t o = ms->cut[j]; rk += (o < v || (o == v && j < tid)) ? 1u : 0u; }
                }
                __syncthreads();
                if (tid >= 1 && tid < NR) ms->cut[rk] = v;
            }
            if (tid == 0) {
                uint32_t s = 0;   // cut[0]: end of the cluster that touches slot 0
                for (;;) {
                    const uint32_t z = ~bm[s >> 5];
                    if (z) { s += (uint32_t)(__ffs(z) - 1); break; }
                    s += 32;
                    if (s >= SLOTS + GUARD_BITS) break;
                }
                ms->cut[0] = s;
            }
            if (tid == NR) {
                ms->cut[NR] = SLOTS + GUARD_BITS;
                uint32_t ts = SLOTS;
                if (V == 1 && ((bm[(SLOTS - 1) >> 5] >> 31) & 1u)) {
                    ts = SLOTS - 1;   // walk down to the first slot of the run that ends at the last slot
                    while (ts > 0) {
                        if ((ts & 31) == 0 && bm[(ts - 1) >> 5] == 0xFFFFFFFFu) { ts -= 32; continue; }
                        if (!((bm[(ts - 1) >> 5] >> ((ts - 1) & 31)) & 1u)) break;
                        --ts;
                    }
                }
                ms->top_start = ts;
            }
        }
        __syncthreads();
        if (tid < ((SLOTS + GUARD_BITS) >> 12)) {
            const uint32_t slot = tid << 12;
            uint32_t r = 0;
#pragma unroll
            for (uint32_t st = 16; st > 0; st >>= 1) if (ms->cut[r + st] <= slot) r += st;
            ms->rtab[tid] = (uint8_t)r;
        }
        if (tid == 0) {
            ms->sp_lo_end = bm_rank(bm, pre16, ms->cut[0]);
            ms->sp_hi_start = ms->top_start < SLOTS ? bm_rank(bm, pre16, ms->top_start) : 0xFFFFFFFFu;
        }
        for (uint32_t i = tid; i < NR * NR; i += NTHREADS) (&ms->cnt[0][0])[i] = 0;
        __syncthreads();

        PHASE_STAMP(2);
        // ---------------- P3: compact index + stable partition by range into per-range lists
        const uint32_t slice = ((len - C + NTHREADS - 1) / NTHREADS) * 32;   // new positions per warp (multiple of 32)
        const uint32_t p_lo = (C + warp * slice < len) ? C + warp * slice : len, p_hi = (p_lo + slice < len) ? p_lo + slice : len;
        const uint32_t top_start = ms->top_start;
        // pass 1: counts per (warp, range)
        uint32_t h_next = p_lo + lane < p_hi ? tokb[p_lo + lane] : 0u;
        for (uint32_t base = p_lo; base < p_hi; base += 32) {
            const uint32_t i = base + lane;
            const uint32_t h = h_next;
            if (i + 32 < p_hi) h_next = tokb[i + 32];        // prefetch
            if (i < p_hi) {
                // A position whose home slot is a cluster of its own (both neighbours free in the no-expiry
                // occupancy) is the only entry that ever touches that slot: it lands there, its find sees a slot
                // that was never used (-> literal candidate) and no other probe walk reaches it. Such entries
                // (a third of enwik-shaped text) skip the partition and the simulation altogether. Not on a
                // slice that hands its slots over (cs[] would be missing).
                const uint32_t wi = h >> 5, bi = h & 31u, wv = bm[wi];
                const uint32_t below = bi ? (wv >> (bi - 1u)) & 1u : (wi ? bm[wi - 1] >> 31 : 1u);
                const uint32_t above = bi != 31u ? (wv >> (bi + 1u)) & 1u : (bm[wi + 1] & 1u);
                if (last_slice && !below && !above && h != SLOTS - 1u) tokb[i] = LONER;
                else {
                uint32_t r = ms->rtab[h >> 12];   // largest r with cut[r] <= h (0 also for the cluster below cut[0])
                while (ms->cut[r + 1] <= h) ++r;  // (cut[NR] lies beyond every slot)
                if (V == 1 && h >= top_start) r = 0;
                tokb[i] = h | (r << 21);    // kept for pass 2 (the token array is free until P4)
                atomicAdd(&ms->cnt[warp][r], 1u);
                }
            }
        }
        __syncthreads();
        if (tid < NR) {           // totals per range -> range start offsets (thread r sums column r)
            uint32_t t = 0;
            for (uint32_t w = 0; w < NR; ++w) t += ms->cnt[w][tid];
            const uint32_t incl = warp_incl_scan_u32(t);
            ms->rstart[tid] = incl - t;
            if (tid == NR - 1) ms->rstart[NR] = incl;
        }
        __syncthreads();
        {   // turn cnt[w][r] into the running write cursor of (warp w, range r)
            uint32_t mycur = 0;
            if (tid < NR * NR) {
                const uint32_t w = tid >> 5, r = tid & 31;
                mycur = ms->rstart[r];
                for (uint32_t w2 = 0; w2 < w; ++w2) mycur += ms->cnt[w2][r];
            }
            __syncthreads();
            if (tid < NR * NR) ms->cnt[tid >> 5][tid & 31] = mycur;
        }
        __syncthreads();
        // pass 2: ordered scatter (warp = contiguous time slice, lanes in position order)
        uint32_t hr_next = p_lo + lane < p_hi ? tokb[p_lo + lane] : 0u;
        for (uint32_t base = p_lo; base < p_hi; base += 32) {
            const uint32_t i = base + lane;
            const bool in_range = i < p_hi;
            const uint32_t hr = hr_next;
            if (i + 32 < p_hi) hr_next = tokb[i + 32];       // prefetch
            const bool loner = in_range && hr == LONER;
            // (a loner keeps its LONER mark in tokb: P5 / P6 read it as "literal candidate")
            const bool valid = in_range && !loner;
            const uint32_t r = valid ? hr >> 21 : 0xFFu;
            // MATCH.ANY takes one step per distinct value: three matches on pieces of the range id (4 + 4 + 2 values, the
            // invalid lanes share a fifth in the first) instead of one on up to 33 values (two matches, 8 + 4: 44.6 ms per
            // GB, three: 44.2); issued first, their latency overlaps the rank below
            const uint32_t peers = __match_any_sync(0xffffffffu, r >> 3) & __match_any_sync(0xffffffffu, (r >> 1) & 3u) & __match_any_sync(0xffffffffu, r & 1u);
            uint32_t c = 0;
            if (valid) c = bm_rank(bm, pre16, hr & 0x1FFFFFu);
            const uint32_t myrank = __popc(peers & lt_mask);
            uint32_t basepos = 0;
            if (valid) basepos = ms->cnt[warp][r];
            __syncwarp();
            if (valid) {
                lists[basepos + myrank] = i | (c << 16);
                if (myrank == 0) ms->cnt[warp][r] = basepos + __popc(peers);
            }
            __syncwarp();
        }
        if (C) {                   // compact slot of every carried entry (the bitmap is still alive here)
            for (uint32_t i = tid; i < C; i += NTHREADS) {
                const uint32_t r = raw_carry[i];
                ccar[i] = r != NONE ? bm_rank(bm, pre16, r) : NONE;
            }
        }
        if (!last_slice) {         // raw slot of every compact slot, for the hand-over at the end of the slice
            for (uint32_t k = 0; k < 5; ++k) {
                const uint32_t ch = tid * 5 + k;
                if (ch >= PRE_N) break;
                uint32_t r = pre16[2 * ch];
                for (uint32_t wq = 0; wq < PRE_CHUNK; ++wq) {
                    uint32_t x = bm[ch * PRE_CHUNK + wq];
                    while (x) { const uint32_t bpos = (uint32_t)(__ffs(x) - 1); x &= x - 1; rawmap[r++] = ((ch * PRE_CHUNK + wq) << 5) + bpos; }
                }
            }
        }
        __syncthreads();

        PHASE_STAMP(3);
        // ---------------- P4: table simulation. T[k] = position+1 of the entry in compact slot k, 0 = never used
        for (uint32_t i = tid; i < (T_ENTRIES + 2052) / 2 + 2; i += NTHREADS) reinterpret_cast<uint32_t*>(big)[i] = 0;   // T and B1
        uint32_t* S1 = pre;   // rank prefix is dead: per-group pattern-signature filter, u32[2052]
        for (uint32_t i = tid; i < 2052; i += NTHREADS) S1[i] = 0;
        __syncthreads();
        if (C) {                   // the carried entries: live at the start of the slice, in the slots they were placed in
            for (uint32_t i = tid; i < C; i += NTHREADS) {
                const uint32_t c = ccar[i];
                if (c != NONE) { T[c] = (uint16_t)(i + 1); atomicOr(&S1[c >> 5], 1u << ((sm_word(data, i) * 0x9E3779B1u) >> 27)); }   // and its pattern signature
            }
            __syncthreads();
        }
        {
            const uint32_t sp_lo_end = ms->sp_lo_end, sp_hi_start = ms->sp_hi_start;
            uint32_t cur = ms->rstart[warp];
            const uint32_t end = ms->rstart[warp + 1];
            const long long t_p4 = CLK();
            uint32_t st_rounds = 0, st_coop = 0, st_entries = end - cur, st_iters = 0;
            long long tq = 0, tc = 0, tm = 0, tl = 0, t_mark;
            // slot-0 clear queue (only warp 0 ever uses it): ABSOLUTE clear times, carried from slice to slice
            uint32_t qh = 0, qt = 0;
            if (warp == 0) {
                if (C == 0) { if (lane == 0) ms->clr[0] = W - 1; qt = 1; }
                else { qh = ms->cq_h; qt = ms->cq_t; }
            }
            __syncwarp();
            // the list entries cur .. cur+31 (ent) and cur+32 .. cur+63 (pf) are kept in registers, pf being
            // re-read right after every advance: the L2 latency of the list stays off the critical path
            uint32_t ent = cur + lane < end ? lists[cur + lane] : 0u;
            uint32_t pf = cur + 32 + lane < end ? lists[cur + 32 + lane] : 0u;
            auto advance = [&](uint32_t by) {
                const uint32_t a = __shfl_sync(0xffffffffu, ent, (lane + by) & 31u), c = __shfl_sync(0xffffffffu, pf, (lane + by) & 31u);
                ent = lane + by < 32 ? a : c;
                cur += by;
                pf = cur + 32 + lane < end ? lists[cur + 32 + lane] : 0u;
            };
            // ---- a range that is ONE pure chain (every entry has the same home and the same 4-byte pattern: constant or
            // few-symbol input, chains of tens of thousands of entries) needs no table: find() only looks at the home
            // slot, whose occupant changes exactly when an entry arrives after the occupant has expired (that entry finds
            // nothing and takes the slot: lz77.c:55-108 with first-fit from the home). Verified on the fly; any other
            // entry in the list and the simulation below starts over (it rewrites every token candidate).
            if (C == 0 && last_slice && end - cur >= 64) {
                const uint32_t e0 = lists[cur];
                const uint32_t kk0 = e0 >> 16, w0 = sm_word(data, e0 & 0xFFFFu);
                bool pure = !(kk0 < sp_lo_end || kk0 >= sp_hi_start);
                uint32_t occ = NONE;                                   // position of the home slot's occupant
                for (uint32_t c2 = cur; pure && c2 < end; c2 += 32) {
                    const bool in = c2 + lane < end;
                    const uint32_t e = in ? lists[c2 + lane] : e0;
                    const uint32_t q = e & 0xFFFFu;
                    if (__any_sync(0xffffffffu, (e >> 16) != kk0 || sm_word(data, q) != w0)) { pure = false; break; }
                    uint32_t m = NONE;
                    bool res = !in;
                    for (;;) {
                        if (!res && occ != NONE && occ + W >= q) { m = occ; res = true; }   // the occupant is live: match
                        const uint32_t un = __ballot_sync(0xffffffffu, !res);
                        if (!un) break;
                        const int f = __ffs(un) - 1;                   // first entry that finds the slot dead: literal, takes the slot
                        occ = __shfl_sync(0xffffffffu, q, f);
                        if ((int)lane == f) res = true;
                    }
                    if (in) {
                        uint32_t tk = 0;
                        const bool reject = (m == NONE) || (V ? (q - m >= W - 1) : (q - m == W));
                        if (!reject) tk = (q - m) | (match_len<MAXLEN>(data, m, q) << 16);
                        tokb[q] = tk;
                    }
                }
                if (pure) cur = end;
            }
            while (cur < end) {
                const uint32_t lanes = end - cur < 32 ? end - cur : 32;
                uint32_t q = 0, kk = 0;
                bool special = false;
                if (lane < lanes) {
                    q = ent & 0xFFFFu; kk = ent >> 16;
                    special = kk < sp_lo_end || kk >= sp_hi_start;
                }
                const uint32_t spmask = __ballot_sync(0xffffffffu, special);
                if (spmask & 1u) {
                    // ---- serial path for the cluster that touches slot 0 / the table end: exact reference order
                    const uint32_t q0 = __shfl_sync(0xffffffffu, q, 0), c0 = __shfl_sync(0xffffffffu, kk, 0);
                    if (lane == 0) {
                        const uint32_t q0a = B0 + q0;                      // absolute time
                        auto clear0 = [&]() {                              // the early clear wipes whatever sits in slot 0
                            if (sp_lo_end) { const uint32_t v0 = T[0]; if (v0 && !last_slice) cs[v0 - 1] = NONE; T[0] = 0; B1[0] = 0; }
                        };
                        while (qh < qt && ms->clr[qh & 63] < q0a) { clear0(); ++qh; }
                        const uint32_t dthr = q0 > W ? q0 - W : 0;
                        const uint32_t w = sm_word(data, q0);
                        uint32_t k = c0, m = NONE;
                        bool ran_off = false;
                        for (;;) {
                            const uint32_t v = T[k];
                            if (v <= dthr) break;
                            if (sm_word(data, v - 1) == w) { m = v - 1; break; }
                            if (k + 1 == nslots) { ran_off = true; break; }   // find does not wrap (lz77.c:102, deflate/lz77.c:168)
                            ++k;
                        }
                        uint32_t e = ran_off ? 0 : k;                      // the wrapping insert continues at slot 0
                        for (;;) { if (T[e] <= dthr) break; ++e; if (e == nslots) e = 0; }
                        if (q0 != 65535u) T[e] = (uint16_t)(q0 + 1);
                        if (!last_slice) cs[q0] = e;
                        if (e == 0 && sp_lo_end) { ms->clr[qt & 63] = q0a + W; ++qt; }
                        if (qh < qt && ms->clr[qh & 63] == q0a) { clear0(); ++qh; }
                        // token candidate
                        uint32_t tk = 0;
                        const bool reject = (m == NONE) || (V ? (q0 - m >= W - 1) : (q0 - m == W));
                        if (!reject) tk = (q0 - m) | (match_len<MAXLEN>(data, m, q0) << 16);
                        tokb[q0] = tk;
                    }
                    qh = __shfl_sync(0xffffffffu, qh, 0); qt = __shfl_sync(0xffffffffu, qt, 0);
                    advance(1);
                    __syncwarp();
                    continue;
                }
                const uint32_t L = spmask ? (uint32_t)(__ffs(spmask) - 1) : lanes;   // stop before the first special entry
                bool done = lane >= L;
                const uint32_t dthr = q > W ? q - W : 0;
                const uint32_t w = done ? 0u : sm_word(data, q);
                const uint32_t sig = (w * 0x9E3779B1u) >> 27;   // 5-bit pattern signature for the per-group filter
                uint32_t fm = NONE; bool pend = true;
                // ---- crowded batch (many entries with the same home: the chain of a hot 4-gram). The rounds
                // below would commit about one of them per round. The sequential outcome is instead computed
                // slot by slot: going through the slots that are dead by the end of the batch in ascending
                // order, a slot goes to the EARLIEST entry that is still unplaced, whose home is at/before the
                // slot and for which the slot is already dead -- exactly what the entry itself would have taken,
                // because every lower dead slot it could reach has by then gone to an earlier entry.
                {
                    // crowd detection: how many entries share the home of lane 0 / of the middle lane
                    const uint32_t home_a = __shfl_sync(0xffffffffu, kk, 0), home_b = __shfl_sync(0xffffffffu, kk, L >> 1);
                    const uint32_t same_a = __ballot_sync(0xffffffffu, !done && kk == home_a);
                    const uint32_t same_b = __ballot_sync(0xffffffffu, !done && kk == home_b);
                    // (valid while no placement of the batch can expire inside the batch: its time span is below W)
                    const uint32_t q_first = __shfl_sync(0xffffffffu, q, 0), q_last = __shfl_sync(0xffffffffu, q, L - 1);
                    if (q_last - q_first < W && (__popc(same_a) >= CROWD_MIN || __popc(same_b) >= CROWD_MIN)) {
                        uint16_t* aslot = reinterpret_cast<uint16_t*>(&ms->cnt[warp][0]);   // [32] slot given to lane j (P3 counters are dead)
                        uint16_t* aqpos = aslot + 32;                                       // [32] position of lane j
                        const uint32_t home = kk;
                        const uint32_t home_s = done ? 0xFFFFFFFFu : home, dthr_s = done ? 0xFFFFFFFFu : dthr;
                        const uint32_t dmax = __shfl_sync(0xffffffffu, dthr, L - 1);        // entries are in time order
                        uint32_t unplaced = __ballot_sync(0xffffffffu, !done);
                        uint32_t my_e = NONE;
                        uint32_t g = 0;
                        while (unplaced) {
                            // nothing below the lowest home of an unplaced entry can be taken any more
                            const uint32_t hmin = __reduce_min_sync(0xffffffffu, ((unplaced >> lane) & 1u) ? home : 0xFFFFFFFFu);
                            if ((hmin >> 5) > g) g = hmin >> 5;
                            if (B1[g] > dmax) { ++g; continue; }                            // every slot of the group outlives the batch
                            const uint32_t sl = (g << 5) + lane;
                            const uint32_t tv = T[sl];
                            const uint32_t mn = __reduce_min_sync(0xffffffffu, tv);
                            if (lane == 0) B1[g] = (uint16_t)mn;
                            uint32_t lo = 0;                                                // entries for which the slot is still live
#pragma unroll
                            for (uint32_t st = 16; st > 0; st >>= 1) {
                                const uint32_t d = __shfl_sync(0xffffffffu, dthr_s, (lo + st - 1) & 31u);
                                if (d < tv) lo += st;
                            }
                            const uint32_t qmask = lo >= 32 ? 0u : (0xFFFFFFFFu << lo);
                            const bool cand = tv <= dmax && sl >= hmin && (qmask & unplaced) != 0u;
                            uint32_t cm = __ballot_sync(0xffffffffu, cand);
                            // Fast path: when no unplaced entry has its home strictly inside this group (so the same set
                            // E of entries may take any of its slots) and the candidates' release indices lo are
                            // non-decreasing in slot order (slots of a chain die in the order they were filled), the
                            // slot-by-slot greedy has a closed form: the k-th candidate takes the entry of E-rank
                            // rho_k = max(rank_E(lo_k), rho_{k-1} + 1) = k + max_{i <= k}(rank_E(lo_i) - i), one prefix
                            // maximum for the whole group instead of one ballot + shuffle round per slot.
                            if (cm & (cm - 1u)) {                                           // two or more candidates
                                const uint32_t gs = g << 5;
                                const uint32_t am_g = __ballot_sync(0xffffffffu, home_s <= gs);
                                const uint32_t inside = __ballot_sync(0xffffffffu, home_s > gs && home_s < gs + 32u) & unplaced;
                                uint32_t lom = cand ? lo : 0u;                              // inclusive prefix maximum of lo over the candidates
#pragma unroll
                                for (uint32_t d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, lom, d); if (lane >= d) lom = max(lom, t); }
                                const uint32_t lom_prev = __shfl_up_sync(0xffffffffu, lom, 1);
                                const bool viol = cand && lane > 0 && lo < lom_prev;
                                if (!inside && !__ballot_sync(0xffffffffu, viol)) {
                                    const uint32_t E = unplaced & am_g;
                                    const int k = (int)__popc(cm & lt_mask);
                                    int v = cand ? (int)__popc(E & ((1u << lo) - 1u)) - k : -64;    // cand implies lo < 32
#pragma unroll
                                    for (uint32_t d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, v, d); if (lane >= d) v = max(v, t); }
                                    uint32_t rho = (uint32_t)(v + k);
                                    const bool take = cand && rho < (uint32_t)__popc(E);
                                    uint32_t e = 0;
                                    if (take) {                                             // position of the rho-th set bit of E
                                        uint32_t mk = E, c;
                                        c = __popc(mk & 0xFFFFu); if (rho >= c) { rho -= c; e = 16; mk >>= 16; }
                                        c = __popc(mk & 0xFFu);   if (rho >= c) { rho -= c; e += 8; mk >>= 8; }
                                        c = __popc(mk & 0xFu);    if (rho >= c) { rho -= c; e += 4; mk >>= 4; }
                                        c = __popc(mk & 0x3u);    if (rho >= c) { rho -= c; e += 2; mk >>= 2; }
                                        if (rho >= (mk & 1u)) e += 1;
                                        aslot[e] = (uint16_t)sl;
                                    }
                                    const uint32_t assigned = __reduce_or_sync(0xffffffffu, take ? (1u << e) : 0u);
                                    __syncwarp();
                                    if ((assigned >> lane) & 1u) my_e = aslot[lane];
                                    unplaced &= ~assigned;
                                    cm = 0;
                                }
                            }
                            while (cm && unplaced) {
                                const int j = __ffs(cm) - 1;
                                cm &= cm - 1;
                                const uint32_t sj = (g << 5) + (uint32_t)j;
                                const uint32_t am = __ballot_sync(0xffffffffu, home_s <= sj);   // entries whose home is at/before the slot
                                const uint32_t ej = __shfl_sync(0xffffffffu, qmask, j) & am & unplaced;
                                if (ej) {
                                    const int pick = __ffs(ej) - 1;
                                    if ((int)lane == pick) my_e = sj;
                                    unplaced &= ~(1u << pick);
                                }
                            }
                            ++g;
                        }
                        if (!done) { aslot[lane] = (uint16_t)my_e; aqpos[lane] = (uint16_t)q; atomicOr(&S1[my_e >> 5], 1u << sig); }
                        __syncwarp();
                        if (!done) {
                            // find: first slot of [home, my_e) holding the pattern; a slot that was dead before the
                            // batch has been given to an earlier entry of the batch (its pattern is read from there)
                            uint32_t k2 = home;
#pragma unroll 1
                            while (k2 < my_e) {
                                if (!((S1[k2 >> 5] >> sig) & 1u)) { k2 = (k2 | 31u) + 1u; continue; }   // nothing with this signature in the (rest of the) group
                                const uint32_t v = T[k2];
                                if (v > dthr) { if (sm_word(data, v - 1) == w) { fm = v - 1; break; } }
                                else {
                                    for (uint32_t j = 0; j < lane; ++j) if (aslot[j] == k2) {
                                        if (sm_word(data, aqpos[j]) == w) fm = aqpos[j];
                                        break;
                                    }
                                    if (fm != NONE) break;
                                }
                                ++k2;
                            }
                        }
                        __syncwarp();
                        if (!done) { if (q != 65535u) T[my_e] = (uint16_t)(q + 1); if (!last_slice) cs[q] = my_e; done = true; }
                        __syncwarp();
                        ++st_coop;
                    }
                }
                while (__ballot_sync(0xffffffffu, !done)) {
                    ++st_rounds;
                    t_mark = CLK();
                    // advance every cursor to the first slot that is dead at the lane's time: a few
                    // single-slot probes (most walks are 1-3 slots long), then whole aligned groups of
                    // 32 slots per step (walk_groups), all lanes walking privately and concurrently
                    if (!done) {
                        bool need = true;
#pragma unroll 1
                        for (int it = 0; it < 3; ++it) {
                            const uint32_t v = T[kk];
                            if (v <= dthr) { need = false; break; }
                            if (pend && sm_word(data, v - 1) == w) { fm = v - 1; pend = false; }
                            ++kk;
                        }
                        if (need) { ++st_coop; kk = walk_groups(T, B1, S1, data, kk, dthr, w, sig, pend, fm, st_iters); }
                    }
                    tq += CLK() - t_mark; t_mark = CLK();
                    __syncwarp();
                    tc += CLK() - t_mark; t_mark = CLK();
                    const uint32_t key = done ? 0xFFFFFFFFu : kk;   // the finished lanes share one key: MATCH.ANY takes one step per distinct value
                    const uint32_t peers = __match_any_sync(0xffffffffu, key);
                    const bool blocked = !done && (peers & lt_mask) != 0;
                    const uint32_t cmask = __ballot_sync(0xffffffffu, blocked);
                    const uint32_t firstc = cmask ? (uint32_t)(__ffs(cmask) - 1) : 32u;
                    if (!done && lane < firstc) {
                        if (q != 65535u) T[kk] = (uint16_t)(q + 1);      // position 65535 is never looked up again within the slice
                        if (!last_slice) cs[q] = kk;
                        atomicOr(&S1[kk >> 5], 1u << sig);
                        done = true;
                    }
                    __syncwarp();
                    tm += CLK() - t_mark;
                }
                t_mark = CLK();
                // token candidates of the whole batch at once (outside the commit rounds)
                if (lane < L) {
                    uint32_t tk = 0;
                    const uint32_t m = fm;
                    const bool reject = (m == NONE) || (V ? (q - m >= W - 1) : (q - m == W));
                    if (!reject) tk = (q - m) | (match_len<MAXLEN>(data, m, q) << 16);
                    tokb[q] = tk;
                }
                tl += CLK() - t_mark;
                advance(L);
            }
            if (warp == 0 && lane == 0) { ms->cq_h = qh; ms->cq_t = qt; }
            if (DBG && dbg_stats && lane == 0) {
                uint32_t* o = dbg_stats + (uint64_t)b * 136 + 8 + warp * 4;
                o[0] = (uint32_t)(clock64() - t_p4); o[1] = st_entries | (st_iters << 16); o[2] = st_rounds | ((uint32_t)(tq >> 10) << 16); o[3] = st_coop | ((uint32_t)(tc >> 10) << 16);
                if (warp < 8) dbg_stats[(uint64_t)b * 136 + 7] = 0;
                atomicMax(&dbg_stats[(uint64_t)b * 136 + 7], ((uint32_t)(tm >> 10) << 16) | (uint32_t)(tl >> 10));
            }
        }
        __syncthreads();

        PHASE_STAMP(4);
        // ---------------- P5: greedy parse. adv[p] = bytes consumed by the token that would start at p
        for (uint32_t i0 = tid; i0 < len; i0 += 8 * NTHREADS) {   // eight independent reads in flight per thread
            uint32_t tv[8];
#pragma unroll
            for (uint32_t k = 0; k < 8; ++k) { const uint32_t i = i0 + k * NTHREADS; tv[k] = i < len ? tokb[i] : 0u; if (tv[k] == LONER) tv[k] = 0u; }
#pragma unroll
            for (uint32_t k = 0; k < 8; ++k) {
                const uint32_t i = i0 + k * NTHREADS;
                if (i < len) {
                    if (dbg_tok && C == 0) dbg_tok[(uint64_t)b * MAXB + i] = tv[k];
                    // positions before the parse entry (carried ones, and those covered by the previous slice's last
                    // token) are stepped over one by one and never emitted
                    adv[PADX(i)] = (uint8_t)((i >= p_start && (tv[k] >> 16)) ? (tv[k] >> 16) : 1u);
                }
            }
        }
        __syncthreads();
        const uint32_t nchunks = (len + 63) >> 6;
        if (tid < nchunks) {   // exit function of chunk tid by backward DP over its 64 positions
            const uint32_t lo = tid << 6, cend = lo + 64;
            const uint32_t hi = cend < len ? cend : len;
            for (uint32_t p = hi; p-- > lo;) {
                const uint32_t nx = p + adv[PADX(p)];
                exitof[PADX(p)] = (uint8_t)(nx >= cend ? nx - cend : nx >= hi ? 0u : exitof[PADX(nx)]);   // (a ragged block's last chunk ends at len: nothing behind it was written)
            }
        }
        __syncthreads();
        {   // entry offset of every chunk: 32 super-chunks of 32 chunks
            const uint32_t s = warp;
            uint32_t e = lane;
            for (uint32_t k = 0; k < 32; ++k) {
                const uint32_t ch = s * 32 + k;
                const uint32_t p = (ch << 6) + e;
                if (ch < nchunks && p < len && e < 31) e = exitof[PADX(p)];
            }
            ms->sexit[s][lane] = (uint8_t)e;
        }
        __syncthreads();
        if (tid == 0) {
            uint32_t e = 0;
            for (uint32_t s = 0; s < NR; ++s) { ms->sentry[s] = (uint8_t)e; e = ms->sexit[s][e & 31u]; }
        }
        __syncthreads();
        uint8_t* centry = reinterpret_cast<uint8_t*>(pre) + 8192;   // rank prefix is dead: u8[1024] chunk entry offsets (clear of the V0 staging spill)
        if (lane == 0) {
            uint32_t e = ms->sentry[warp];
            for (uint32_t k = 0; k < 32; ++k) {
                const uint32_t ch = warp * 32 + k;
                centry[ch] = (uint8_t)e;
                const uint32_t p = (ch << 6) + e;
                if (ch < nchunks && p < len) e = exitof[PADX(p)];
                if (ch + 1 == nchunks) ms->ecarry = e;      // positions of the next slice covered by this slice's last token
            }
        }
        __syncthreads();
        // per-chunk output size, CTA exclusive scan. V1 also notes, for every token start, its byte
        // offset inside the chunk's output (in 2-byte units, exitof is dead now) so that P6 can emit
        // one position per thread with coalesced reads
        uint8_t* orel = exitof;
        if (V == 1) {
            for (uint32_t i = tid; i < (PADDED >> 2); i += NTHREADS) reinterpret_cast<uint32_t*>(orel)[i] = 0xFFFFFFFFu;
            __syncthreads();
        }
        uint32_t my_units = 0;   // bytes (V1) or bits (V0)
        if (tid < nchunks) {
            const uint32_t cend = (tid << 6) + 64;
            const uint32_t hi = cend < len ? cend : len;
            for (uint32_t p = (tid << 6) + centry[tid]; p < hi; p += adv[PADX(p)]) {
                if (p < p_start) continue;            // (the for's increment still advances by adv[p] == 1)
                const bool lit = adv[PADX(p)] == 1;   // matches are at least 4 long
                if (V == 1) orel[PADX(p)] = (uint8_t)(my_units >> 1);
                my_units += V ? (lit ? 2u : 4u) : (lit ? 9u : 19u);
            }
        }
        uint32_t my_off;
        {
            const uint32_t incl = warp_incl_scan_u32(my_units);
            if (lane == 31) ms->scan[warp] = incl;
            __syncthreads();
            if (warp == 0) {
                const uint32_t t = ms->scan[lane];
                const uint32_t ti = warp_incl_scan_u32(t);
                ms->scan[lane] = ti - t;
                if (lane == 31) ms->scan[32] = ti;
            }
            __syncthreads();
            my_off = ms->scan[warp] + incl - my_units;
        }
        const uint32_t total_units = ms->scan[32];
        uint8_t* out = scratch + (uint64_t)b * stride + (V ? out_units : (out_units >> 5) * 4);   // V0: word holding the first bit

        PHASE_STAMP(5);
        // ---------------- P6: emission
        if (V == 1) {
            uint32_t* coff = pre + 3072;          // u32[1024] output offset of every chunk (the P4 tables are dead)
            if (tid < nchunks) coff[tid] = my_off;
            __syncthreads();
            for (uint32_t p0 = tid; p0 < len; p0 += 4 * NTHREADS) {
                uint32_t tk[4];
#pragma unroll
                for (uint32_t k = 0; k < 4; ++k) { const uint32_t p = p0 + k * NTHREADS; tk[k] = p < len ? tokb[p] : 0u; if (tk[k] == LONER) tk[k] = 0u; }
#pragma unroll
                for (uint32_t k = 0; k < 4; ++k) {
                    const uint32_t p = p0 + k * NTHREADS;
                    if (p >= len) break;
                    const uint32_t r = orel[PADX(p)];
                    if (r == 0xFFu) continue;     // not a token start of the greedy parse
                    const uint32_t o = coff[p >> 6] + 2u * r, t = tk[k];
                    if (t == 0) *reinterpret_cast<uint16_t*>(out + o) = (uint16_t)((uint32_t)data[p] << 8);
                    else {
                        const uint32_t off = t & 0xFFFFu, ml = t >> 16;
                        *reinterpret_cast<uint16_t*>(out + o) = (uint16_t)(1u | ((off & 0xFFu) << 8));
                        *reinterpret_cast<uint16_t*>(out + o + 2) = (uint16_t)((off >> 8) | (ml << 8));
                    }
                }
            }
            if (tid == 0 && last_slice) { block_sizes[b] = out_units + total_units; block_bytes[b] = out_units + total_units; }
        } else {
            // LSB-first bit stream staged in shared memory (exitof is dead now), then stored coalesced
            const uint32_t sh0 = (uint32_t)(out_units & 31);   // bits of the first word that belong to the previous slice
            const uint32_t nwords = ((sh0 + total_units) >> 5) + 2;
            __syncthreads();
            for (uint32_t i = tid; i < nwords; i += NTHREADS) stage[i] = 0;
            __syncthreads();
            if (tid < nchunks && my_units) {
                const uint32_t cend = (tid << 6) + 64;
                const uint32_t hi = cend < len ? cend : len;
                uint32_t wi = (sh0 + my_off) >> 5, have = (sh0 + my_off) & 31;
                bool first = have != 0;
                uint64_t acc = 0;
                for (uint32_t p = (tid << 6) + centry[tid]; p < hi; p += adv[PADX(p)]) {
                    if (p < p_start) continue;
                    uint32_t t = tokb[p];
                    if (t == LONER) t = 0u;
                    uint32_t v, nb;
                    if (t == 0) { v = (uint32_t)data[p] << 1; nb = 9; }
                    else { v = 1u | ((t & 0xFFFFu) << 1) | ((t >> 16) << 15); nb = 19; }
                    acc |= (uint64_t)v << have;
                    have += nb;
                    if (have >= 32) {
                        if (first) { atomicOr(&stage[wi], (uint32_t)acc); first = false; } else stage[wi] = (uint32_t)acc;
                        ++wi; acc >>= 32; have -= 32;
                    }
                }
                if (have) atomicOr(&stage[wi], (uint32_t)acc);
            }
            __syncthreads();
            for (uint32_t i = tid; i < nwords; i += NTHREADS) {
                if (i == 0 && sh0) reinterpret_cast<uint32_t*>(out)[0] |= stage[0];   // word shared with the previous slice (same CTA, earlier)
                else reinterpret_cast<uint32_t*>(out)[i] = stage[i];
            }
            if (tid == 0 && last_slice) { block_sizes[b] = out_units + total_units; block_bytes[b] = (out_units + total_units) / 8 + 1; }
        }
        out_units += total_units;
        if (!last_slice) {
            // hand-over: the last W positions of this slice are the next slice's carried entries; their raw slots come
            // from the compact slot each was placed in (0xFFFF = wiped by the early slot-0 clear)
            __syncthreads();
            for (uint32_t i = tid; i < W; i += NTHREADS) {
                const uint32_t c = cs[len - W + i];
                raw_carry[i] = c != NONE ? rawmap[c] : NONE;
            }
        }
        __syncthreads();   // smem is reused by the next slice / block
        PHASE_STAMP(6);
        if (last_slice) break;
        n0 = B0 + len;
        }   // slices
    }
}

}  // namespace

// blocks above 65536 bytes are simulated in slices (positions stay 16-bit inside a slice)
bool lz77_v2_supported(uint64_t bs) { return bs <= (1ull << 31); }

// scratch slots: 13 = lists, 14 = tok
int lz77_v2_launch(b200_ctx* ctx, int variant, const uint8_t* d_in, uint64_t n, uint64_t bs, uint64_t nblocks,
                   uint8_t* scratch, uint64_t stride, uint64_t* d_block_sizes, uint64_t* block_bytes, uint32_t* dbg_tok,
                   const uint32_t* blist, const uint32_t* bcount) {
    static bool attr_done_dev[64] = {};   // the attribute is per device
    bool& attr_done = attr_done_dev[ctx->device >= 0 && ctx->device < 64 ? ctx->device : 0];
    if (!attr_done) {
        CUDA_TRY(cudaFuncSetAttribute(lz77_v2_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        CUDA_TRY(cudaFuncSetAttribute(lz77_v2_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        CUDA_TRY(cudaFuncSetAttribute(lz77_v2_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        CUDA_TRY(cudaFuncSetAttribute(lz77_v2_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        attr_done = true;
    }
    uint64_t grid = (uint64_t)ctx->sm_count;
    if (grid > nblocks) grid = nblocks;
    uint32_t *lists, *tok; uint8_t* carry;
    B200_TRY(b200_scratch(ctx, B200_SLOT(ctx, 13), (size_t)grid * MAXB * 4 + 64, reinterpret_cast<void**>(&lists)));
    B200_TRY(b200_scratch(ctx, B200_SLOT(ctx, 14), (size_t)grid * MAXB * 4 + 64, reinterpret_cast<void**>(&tok)));
    B200_TRY(b200_scratch(ctx, B200_SLOT(ctx, 15), bs > MAXB ? (size_t)grid * CARRY_BYTES + 64 : 64, reinterpret_cast<void**>(&carry)));
#define LZ_V2_LAUNCH(V, D) lz77_v2_kernel<V, D><<<(unsigned)grid, NTHREADS, SMEM_BYTES, ctx->stream>>>(d_in, n, (uint32_t)bs, (uint32_t)nblocks, lists, tok, carry, scratch, stride, d_block_sizes, block_bytes, dbg_tok, blist, bcount)
    if (dbg_tok) { if (variant == 0) LZ_V2_LAUNCH(0, true); else LZ_V2_LAUNCH(1, true); }
    else { if (variant == 0) LZ_V2_LAUNCH(0, false); else LZ_V2_LAUNCH(1, false); }
    CUDA_TRY(cudaGetLastError());
    return B200_OK;
}
