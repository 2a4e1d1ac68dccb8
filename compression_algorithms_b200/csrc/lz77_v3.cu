// LZ77 v3: the reference's 2^20-slot hash table emulated slot-exactly inside ONE SM's shared
// memory for blocks of up to 65536 bytes (the reference's BUFFER_SIZE, algorithms/deflate/deflate.h:8),
// with the table SIMULATED only where the outcome is not already decided by the occupancy.
//
// Facts used (DESIGN.md "LZ77 v3"; tools/proto_v3.c is the CPU model that was checked against a serial
// table for every position):
//  * F(p) = find(word(p)) after inserts 0..p-1 is a pure function of the data (every position is inserted
//    exactly once, in order: lz77.c:295,333-336; deflate/lz77.c:228,267-270), the parse only selects.
//  * In the occupancy that results when nothing ever expires (order independent, built with atomicOr linear
//    probing), a maximal run of occupied slots -- a CLUSTER -- has exactly as many entries as slots, its live
//    slots are always a subset of it, and no probe walk ever leaves it. rank(slot) is a dense "compact slot".
//  * An entry whose home is the FIRST slot of its cluster is a HEAD. Only heads of that home can ever take
//    the home slot (nothing spills in from below), an arriving head takes it exactly when its occupant has
//    expired, and find() of a head looks at the home slot first. So, when all heads of a home carry the same
//    4-byte pattern, find(head p) = the current occupant e of the home slot (if e >= p - W), where the
//    occupant changes only at the first arrival after e + W: a "jump chain" of at most len / W + 1 links,
//    computed with one minimum per link -- no table at all. A head alone in its cluster (a LONER) never
//    finds anything.
//  * Only a cluster that also holds an INTRUDER (an entry whose home lies strictly inside it) or two patterns
//    with the same head home needs its placements simulated, and there only the intruders (and the heads of
//    a two-pattern home) need find(). Those "mixed" clusters (28 % of the positions of enwik-shaped text,
//    a handful of long chains plus ~2000 small ones) are simulated SERIALLY, one lane per group of clusters
//    (1024 lanes), with a liveness bitmask + first-fit placement + FIFO expiry: no commit rounds, no conflicts.
//  * The cluster touching slot 0 (and, for the wrapping deflate insert, the table end) is simulated by one
//    lane with the reference's early slot-0 clear (lz77.c:70-85, U10).
//
// Phases per block (one persistent CTA of 1024 threads per SM):
//   P0 load | P1 occupancy bitmap | P2 rank prefix | P3 classify: loner / head / intruder (c, ca per position)
//   P4 mixed + two-pattern marks, first head per home | P5 jump chains | P6 stable partition of the mixed
//   entries into 1024 time-ordered lane lists | P7 lane-serial simulation | P8 greedy parse | P9 emission
#include "common.cuh"
#include "../../include/b200comp.h"

namespace {

constexpr uint32_t SLOTS = 1u << 20;
constexpr uint32_t GUARD_BITS = 65536u;
constexpr uint32_t BM_WORDS = (SLOTS + GUARD_BITS) / 32;     // 34816
constexpr uint32_t PRE_CHUNK = 8;                            // words per rank-prefix entry
constexpr uint32_t PRE_N = BM_WORDS / PRE_CHUNK;             // 4352
constexpr uint32_t NONE = 0xFFFFFFFFu;
constexpr uint32_t LONER = 0xFFFFFFFFu;                      // work[] marker (c == ca == 65535 cannot be a non-loner head)
constexpr uint32_t MAXB = 65536;
constexpr uint32_t NTHREADS = 1024;
constexpr uint32_t NBIN = 1024;                              // lane lists: bin = compact cluster start >> 6
constexpr uint32_t BIN_SPECIAL = NBIN;                       // slot-0 / table-end cluster
constexpr uint32_t LONG_LIST_MIN = 48;                           // list length from which a warp takes the list
constexpr uint32_t CNT_STRIDE = 1026;                        // u16 counters per warp row (even: pairs share a u32)

// shared memory layout (bytes)
constexpr uint32_t OFF_DATA = 0;
constexpr uint32_t SZ_DATA = MAXB + 128;                     // zero pad behind the block (U1)
constexpr uint32_t OFF_BIG = OFF_DATA + SZ_DATA;
constexpr uint32_t SZ_BIG = BM_WORDS * 4;                    // 139264: bitmap | cur | counters | T + M | adv + exit | staging
constexpr uint32_t OFF_PRE = OFF_BIG + SZ_BIG;               // must directly follow BIG (V0 staging may spill)
constexpr uint32_t SZ_PRE = (PRE_N + 4) * 4;                 // 17424: rank prefix | mixed + two-pattern bits | chunk entries, offsets
constexpr uint32_t OFF_MISC = OFF_PRE + SZ_PRE;
constexpr uint32_t SZ_MISC = 8192;
constexpr uint32_t SMEM_BYTES = OFF_MISC + SZ_MISC;          // 230,544 <= 232,448

constexpr uint32_t OFF_T = 0;                                // inside BIG: u16[65536] position of the entry in a compact slot
constexpr uint32_t OFF_M = MAXB * 2;                         // inside BIG: u32[2048] liveness bits of the compact slots
static_assert(OFF_M + MAXB / 8 <= SZ_BIG, "T + M must fit");
static_assert(32 * CNT_STRIDE * 2 <= SZ_BIG, "partition counters must fit");
constexpr uint32_t PADDED = MAXB + (MAXB >> 6) * 4;          // 69632
constexpr uint32_t OFF_ADV = 0;
constexpr uint32_t OFF_EXIT = PADDED;
constexpr uint32_t OFF_STAGE = PADDED;
#define PADX(p) ((p) + (((p) >> 6) << 2))

struct Misc3 {
    uint32_t scan[34];
    uint32_t cut0, top_start, sp_lo_end, sp_hi_start;
    uint32_t p1_next, nlong, next_long, pad2;
    uint32_t clr[64];            // slot-0 clear times
    uint8_t  sexit[32][32];
    uint8_t  sentry[36];
    uint32_t binstart[NBIN + 4]; // first list entry of every lane list; [1024] = special list, [1025] = end
    uint16_t longbin[NBIN];      // lists of more than LONG_LIST_MIN entries: simulated by a whole warp each (stage B of P7)
};
static_assert(sizeof(Misc3) <= SZ_MISC, "misc region too small");

template <int V> struct Cfg;
template <> struct Cfg<0> { static constexpr uint32_t W = 1u << 14, MAXLEN = 15; };
template <> struct Cfg<1> { static constexpr uint32_t W = 1u << 15, MAXLEN = 31; };

__device__ __forceinline__ uint32_t sm_word(const uint8_t* data, uint32_t p) {
    const uint32_t* a = reinterpret_cast<const uint32_t*>(data + (p & ~3u));
    return __funnelshift_r(a[0], a[1], (p & 3u) * 8);
}

// common prefix of data[m..] and data[q..], at least 4 (the hashed word), capped at MAXLEN
// (lz77.c:302-311, deflate/lz77.c:238-247)
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

// token candidate of position q whose find() returned m: 0 = literal, else offset | len << 16
// (reject rules: lz77.c:290 distance == W; deflate/lz77.c:223 distance >= W - 1)
template <int V>
__device__ __forceinline__ uint32_t candidate(const uint8_t* data, uint32_t m, uint32_t q) {
    constexpr uint32_t W = Cfg<V>::W;
    if (m == NONE) return 0u;
    if (V ? (q - m >= W - 1) : (q - m == W)) return 0u;
    return (q - m) | (match_len<Cfg<V>::MAXLEN>(data, m, q) << 16);
}

__device__ __forceinline__ uint32_t bm_rank(const uint32_t* bm, const uint16_t* pre16, uint32_t s) {
    const uint32_t wi = s >> 5, j = wi & 3u;
    const uint4 q = *reinterpret_cast<const uint4*>(bm + (wi & ~3u));
    const uint32_t wd[4] = {q.x, q.y, q.z, q.w};
    uint32_t r = pre16[wi >> 2];
    const uint32_t below = (1u << (s & 31)) - 1u;
#pragma unroll
    for (uint32_t k = 0; k < 4; ++k) r += __popc(wd[k] & (k < j ? 0xFFFFFFFFu : (k == j ? below : 0u)));
    return r;
}

// minimum of a u16 cell of shared memory (two cells per word): a plain read decides nearly every call
// (positions arrive in ascending order, only the first head of a home has to write)
__device__ __forceinline__ void cur_min16(uint32_t* cur32, uint32_t c, uint32_t p) {
    uint32_t* a = cur32 + (c >> 1);
    const uint32_t sh = (c & 1u) * 16u;
    uint32_t old = *reinterpret_cast<volatile uint32_t*>(a);
    while (((old >> sh) & 0xFFFFu) > p) {
        const uint32_t nw = (old & ~(0xFFFFu << sh)) | (p << sh);
        const uint32_t prev = atomicCAS(a, old, nw);
        if (prev == old) break;
        old = prev;
    }
}

#define CLK3() (DBG ? clock64() : 0ll)
template <int V, bool DBG>
__global__ void __launch_bounds__(NTHREADS, 1) lz77_v3_kernel(const uint8_t* __restrict__ in, uint64_t n, uint32_t bs, uint32_t nblocks,
                                                             uint32_t* __restrict__ work_all, uint32_t* __restrict__ cand_all,
                                                             uint32_t* __restrict__ la_all, uint32_t* __restrict__ lb_all,
                                                             uint8_t* __restrict__ scratch, uint64_t stride,
                                                             uint64_t* __restrict__ block_sizes, uint64_t* __restrict__ block_bytes,
                                                             uint32_t* __restrict__ dbg_tok) {
    constexpr uint32_t W = Cfg<V>::W;
    extern __shared__ __align__(16) uint8_t smem[];
    uint8_t* data = smem + OFF_DATA;
    uint8_t* big = smem + OFF_BIG;
    uint32_t* bm = reinterpret_cast<uint32_t*>(big);
    uint32_t* pre = reinterpret_cast<uint32_t*>(smem + OFF_PRE);
    uint16_t* pre16 = reinterpret_cast<uint16_t*>(smem + OFF_PRE);
    Misc3* ms = reinterpret_cast<Misc3*>(smem + OFF_MISC);
    uint32_t* cur32 = reinterpret_cast<uint32_t*>(big);                 // P4/P5: u16[65536] first unresolved head of every home
    volatile uint16_t* cur16 = reinterpret_cast<volatile uint16_t*>(big);
    uint16_t* cnt16 = reinterpret_cast<uint16_t*>(big);                 // P6: [32][CNT_STRIDE]
    uint32_t* cnt32 = reinterpret_cast<uint32_t*>(big);
    uint16_t* T = reinterpret_cast<uint16_t*>(big + OFF_T);
    uint32_t* M = reinterpret_cast<uint32_t*>(big + OFF_M);
    uint32_t* MIX = pre;                                                // P4..P7: u32[2048] cluster (by compact start) needs a simulation
    uint32_t* NONUNI = pre + 2048;                                      //         u32[2048] head home with two patterns
    uint8_t* adv = big + OFF_ADV;
    uint8_t* exitof = big + OFF_EXIT;
    uint32_t* stage = reinterpret_cast<uint32_t*>(big + OFF_STAGE);

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint32_t* work = work_all + (uint64_t)blockIdx.x * MAXB;            // hash, then c | ca << 16 (LONER)
    uint32_t* cand = cand_all + (uint64_t)blockIdx.x * MAXB;            // token candidate of every position
    uint32_t* la = la_all + (uint64_t)blockIdx.x * MAXB;                // lane lists: position | c << 16
    uint32_t* lb = lb_all + (uint64_t)blockIdx.x * MAXB;                //   flags (1 find, 2 head) until placed, then position | slot << 16

    uint32_t* dbg_stats = dbg_tok ? dbg_tok + (uint64_t)nblocks * MAXB : nullptr;   // [block][136]
    for (uint32_t b = blockIdx.x; b < nblocks; b += gridDim.x) {
        const long long t_begin = CLK3();
#define PHASE_STAMP3(k) do { if (DBG && dbg_stats && tid == 0) dbg_stats[(uint64_t)b * 136 + (k)] = (uint32_t)(clock64() - t_begin); } while (0)
        const uint8_t* src = in + (uint64_t)b * bs;
        const uint32_t len = (uint32_t)(n - (uint64_t)b * bs < bs ? n - (uint64_t)b * bs : bs);
        const uint32_t nslots = len;

        // ---------------- P0: block -> shared memory, zero pad behind it, zero bitmap
        {
            const bool al = (reinterpret_cast<uintptr_t>(src) & 15) == 0;
            for (uint32_t i = tid * 16; i < len + 128 && i < SZ_DATA; i += NTHREADS * 16) {
                if (al && i + 16 <= len) *reinterpret_cast<uint4*>(data + i) = __ldg(reinterpret_cast<const uint4*>(src + i));
                else for (uint32_t k = 0; k < 16 && i + k < SZ_DATA; ++k) data[i + k] = (i + k < len) ? __ldg(src + i + k) : 0;
            }
            for (uint32_t i = tid; i < BM_WORDS; i += NTHREADS) bm[i] = 0;
            for (uint32_t i = tid; i < BM_WORDS / 32; i += NTHREADS) pre[i] = 0;   // P1's "word is full" summary
            if (tid == 0) { ms->p1_next = 0; ms->nlong = 0; ms->next_long = 0; }
        }
        __syncthreads();

        // ---------------- P1: no-expiry occupancy (order independent) by atomic linear probing; rows of 32
        // positions from a shared counter; one summary bit per bitmap word ("known to be full", never cleared)
        // lets the k-th occurrence of a hot 4-gram jump to the first word that may have room
        {
            uint32_t* summ = pre;
            for (;;) {
                uint32_t row = 0;
                if (lane == 0) row = atomicAdd(&ms->p1_next, 32u);
                row = __shfl_sync(0xffffffffu, row, 0);
                if (row >= len) break;
                const uint32_t i = row + lane;
                if (i >= len) continue;
                uint32_t s = lz_hash(sm_word(data, i));
                work[i] = s;
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
                        continue;
                    }
                    uint32_t nw = wi + 1;
                    uint32_t open = ~summ[nw >> 5] & (0xFFFFFFFFu << (nw & 31));
                    while (!open) { nw = ((nw >> 5) + 1) << 5; open = ~summ[nw >> 5]; }
                    s = ((nw & ~31u) + (uint32_t)(__ffs(open) - 1)) << 5;
                    if (V == 1 && s >= SLOTS) s = 0;   // the deflate insert wraps (deflate/lz77.c:99-101)
                }
            }
        }
        __syncthreads();
        PHASE_STAMP3(0);

        // ---------------- P2: rank prefix per 4-word half chunk, bounds of the special cluster(s)
        {
            const uint32_t c0 = tid * 5;
            uint32_t part[5], half[5], mine = 0;
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                const uint32_t ch = c0 + k;
                uint32_t s = 0, s4 = 0;
                if (ch < PRE_N) {
                    const uint4 q0 = *reinterpret_cast<const uint4*>(bm + ch * PRE_CHUNK), q1 = *reinterpret_cast<const uint4*>(bm + ch * PRE_CHUNK + 4);
                    s4 = __popc(q0.x) + __popc(q0.y) + __popc(q0.z) + __popc(q0.w);
                    s = s4 + __popc(q1.x) + __popc(q1.y) + __popc(q1.z) + __popc(q1.w);
                }
                part[k] = s; half[k] = s4; mine += s;
            }
            const uint32_t incl = warp_incl_scan_u32(mine);
            __syncthreads();                       // (summary bits in pre are dead from here on)
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
            if (tid == 0) {
                uint32_t s = 0;   // end of the cluster that touches slot 0
                for (;;) {
                    const uint32_t z = ~bm[s >> 5];
                    if (z) { s += (uint32_t)(__ffs(z) - 1); break; }
                    s += 32;
                    if (s >= SLOTS + GUARD_BITS) break;
                }
                ms->cut0 = s;
            }
            if (tid == 32) {
                uint32_t ts = SLOTS;
                if (V == 1 && ((bm[(SLOTS - 1) >> 5] >> 31) & 1u)) {
                    ts = SLOTS - 1;   // first slot of the run that ends at the last slot
                    while (ts > 0) {
                        if ((ts & 31) == 0 && bm[(ts - 1) >> 5] == 0xFFFFFFFFu) { ts -= 32; continue; }
                        if (!((bm[(ts - 1) >> 5] >> ((ts - 1) & 31)) & 1u)) break;
                        --ts;
                    }
                }
                ms->top_start = ts;
            }
            __syncthreads();
            if (tid == 0) {
                ms->sp_lo_end = bm_rank(bm, pre16, ms->cut0);
                ms->sp_hi_start = ms->top_start < SLOTS ? bm_rank(bm, pre16, ms->top_start) : 0xFFFFFFFFu;
            }
        }
        __syncthreads();
        PHASE_STAMP3(1);
        const uint32_t cut0 = ms->cut0, top_start = ms->top_start, sp_lo_end = ms->sp_lo_end, sp_hi_start = ms->sp_hi_start;

        // ---------------- P3: classify every position. work[i] = c | ca << 16 (compact home, compact cluster start),
        // LONER for a head alone in its cluster (cand = literal at once)
        for (uint32_t i0 = tid; i0 < len; i0 += 4 * NTHREADS) {
            uint32_t hv[4];
#pragma unroll
            for (uint32_t k = 0; k < 4; ++k) { const uint32_t i = i0 + k * NTHREADS; hv[k] = i < len ? work[i] : 0u; }
#pragma unroll
            for (uint32_t k = 0; k < 4; ++k) {
                const uint32_t i = i0 + k * NTHREADS;
                if (i >= len) break;
                const uint32_t h = hv[k], wi = h >> 5, bi = h & 31u, wv = bm[wi];
                if (h < cut0 || (V == 1 && h >= top_start)) {
                    const uint32_t c = bm_rank(bm, pre16, h);
                    work[i] = c | (c << 16);
                    continue;
                }
                const uint32_t below = bi ? (wv >> (bi - 1u)) & 1u : bm[wi - 1] >> 31;
                const uint32_t above = bi != 31u ? (wv >> (bi + 1u)) & 1u : (bm[wi + 1] & 1u);
                if (!below && !above) { work[i] = LONER; cand[i] = 0u; continue; }
                const uint32_t c = bm_rank(bm, pre16, h);
                uint32_t ca = c;
                if (below) {   // walk down to the first slot of the cluster
                    uint32_t x = ~wv & ((1u << bi) - 1u), w2 = wi;
                    while (!x) { --w2; x = ~bm[w2]; }
                    const uint32_t a = (w2 << 5) + 32u - (uint32_t)__clz(x);
                    ca = c - (h - a);
                }
                work[i] = c | (ca << 16);
            }
        }
        __syncthreads();
        PHASE_STAMP3(2);

        // ---------------- P4: the bitmap is dead. cur[c] = first head of home c; MIX[ca] for every cluster with an
        // intruder; then NONUNI[c] (and MIX) for a head home that carries two different patterns
        for (uint32_t i = tid; i < MAXB / 2; i += NTHREADS) cur32[i] = 0xFFFFFFFFu;
        for (uint32_t i = tid; i < 4096; i += NTHREADS) pre[i] = 0;          // MIX, NONUNI
        __syncthreads();
        for (uint32_t i0 = tid; i0 < len; i0 += 4 * NTHREADS) {
            uint32_t v[4];
#pragma unroll
            for (uint32_t k = 0; k < 4; ++k) { const uint32_t i = i0 + k * NTHREADS; v[k] = i < len ? work[i] : LONER; }
#pragma unroll
            for (uint32_t k = 0; k < 4; ++k) {
                const uint32_t i = i0 + k * NTHREADS;
                if (v[k] == LONER) continue;
                const uint32_t c = v[k] & 0xFFFFu, ca = v[k] >> 16;
                if (c < sp_lo_end || c >= sp_hi_start) continue;
                if (ca != c) atomicOr(&MIX[ca >> 5], 1u << (ca & 31));
                else cur_min16(cur32, c, i);
            }
        }
        __syncthreads();
        for (uint32_t i0 = tid; i0 < len; i0 += 4 * NTHREADS) {
            uint32_t v[4];
#pragma unroll
            for (uint32_t k = 0; k < 4; ++k) { const uint32_t i = i0 + k * NTHREADS; v[k] = i < len ? work[i] : LONER; }
#pragma unroll
            for (uint32_t k = 0; k < 4; ++k) {
                const uint32_t i = i0 + k * NTHREADS;
                if (v[k] == LONER) continue;
                const uint32_t c = v[k] & 0xFFFFu, ca = v[k] >> 16;
                if (ca != c || c < sp_lo_end || c >= sp_hi_start) continue;
                const uint32_t e0 = cur16[c];
                if (e0 != i && sm_word(data, e0) != sm_word(data, i)) { atomicOr(&NONUNI[c >> 5], 1u << (c & 31)); atomicOr(&MIX[c >> 5], 1u << (c & 31)); }
            }
        }
        __syncthreads();
        PHASE_STAMP3(3);

        // ---------------- P5: jump chains. Every head of a one-pattern home finds the occupant of its home slot: the
        // first head e of the home, as long as e >= p - W; the heads beyond that start the next link (minimum again)
        {
            unsigned long long unres = 0ull;
            for (uint32_t k = 0; k < 64; ++k) {
                const uint32_t i = tid + k * NTHREADS;
                if (i >= len) break;
                const uint32_t v = work[i];
                if (v == LONER) continue;
                const uint32_t c = v & 0xFFFFu, ca = v >> 16;
                if (ca != c || c < sp_lo_end || c >= sp_hi_start) continue;
                if ((NONUNI[c >> 5] >> (c & 31)) & 1u) continue;
                const uint32_t e = cur16[c];
                if (e == i) cand[i] = 0u;
                else if (i - e <= W) cand[i] = candidate<V>(data, e, i);
                else unres |= 1ull << k;
            }
            while (__syncthreads_or(unres != 0ull)) {
                for (unsigned long long m = unres; m; m &= m - 1) { const uint32_t i = tid + (uint32_t)(__ffsll((long long)m) - 1) * NTHREADS; cur16[work[i] & 0xFFFFu] = 0xFFFFu; }
                __syncthreads();
                for (unsigned long long m = unres; m; m &= m - 1) { const uint32_t i = tid + (uint32_t)(__ffsll((long long)m) - 1) * NTHREADS; cur_min16(cur32, work[i] & 0xFFFFu, i); }
                __syncthreads();
                for (unsigned long long m = unres; m; m &= m - 1) {
                    const uint32_t k = (uint32_t)(__ffsll((long long)m) - 1), i = tid + k * NTHREADS;
                    const uint32_t e = cur16[work[i] & 0xFFFFu];
                    if (e == i) { cand[i] = 0u; unres &= ~(1ull << k); }
                    else if (i - e <= W) { cand[i] = candidate<V>(data, e, i); unres &= ~(1ull << k); }
                }
            }
        }
        __syncthreads();
        PHASE_STAMP3(4);

        // ---------------- P6: stable partition of the entries of mixed clusters (and of the special cluster) into
        // 1024 (+1) time-ordered lists: warp = contiguous time slice, count per (warp, bin), ordered scatter
        for (uint32_t i = tid; i < 32 * CNT_STRIDE / 2; i += NTHREADS) cnt32[i] = 0;
        __syncthreads();
        const uint32_t slice = ((len + NTHREADS - 1) / NTHREADS) * 32;
        const uint32_t p_lo = (warp * slice < len) ? warp * slice : len, p_hi = (p_lo + slice < len) ? p_lo + slice : len;
        auto bin_of = [&](uint32_t v) -> uint32_t {
            if (v == LONER) return NONE;
            const uint32_t c = v & 0xFFFFu, ca = v >> 16;
            if (c < sp_lo_end || c >= sp_hi_start) return BIN_SPECIAL;
            return ((MIX[ca >> 5] >> (ca & 31)) & 1u) ? (ca >> 6) : NONE;
        };
        {
            uint32_t v_next = p_lo + lane < p_hi ? work[p_lo + lane] : LONER;
            for (uint32_t base = p_lo; base < p_hi; base += 32) {
                const uint32_t i = base + lane;
                const uint32_t v = v_next;
                v_next = i + 32 < p_hi ? work[i + 32] : LONER;
                const uint32_t bin = i < p_hi ? bin_of(v) : NONE;
                if (bin != NONE) atomicAdd(&cnt32[(warp * CNT_STRIDE + bin) >> 1], (bin & 1u) ? 0x10000u : 1u);
            }
        }
        __syncthreads();
        {
            uint32_t tot = 0;
            for (uint32_t w = 0; w < 32; ++w) { const uint32_t t = cnt16[w * CNT_STRIDE + tid]; cnt16[w * CNT_STRIDE + tid] = (uint16_t)tot; tot += t; }
            uint32_t tot_sp = 0;
            if (tid == 0) for (uint32_t w = 0; w < 32; ++w) { const uint32_t t = cnt16[w * CNT_STRIDE + BIN_SPECIAL]; cnt16[w * CNT_STRIDE + BIN_SPECIAL] = (uint16_t)tot_sp; tot_sp += t; }
            const uint32_t incl = warp_incl_scan_u32(tot);
            if (lane == 31) ms->scan[warp] = incl;
            __syncthreads();
            if (warp == 0) {
                const uint32_t t = ms->scan[lane];
                const uint32_t ti = warp_incl_scan_u32(t);
                ms->scan[lane] = ti - t;
                if (lane == 31) ms->scan[32] = ti;
            }
            __syncthreads();
            ms->binstart[tid] = ms->scan[warp] + incl - tot;
            if (tid == 0) { ms->binstart[NBIN] = ms->scan[32]; ms->binstart[NBIN + 1] = ms->scan[32] + tot_sp; }
        }
        __syncthreads();
        {
            uint32_t v_next = p_lo + lane < p_hi ? work[p_lo + lane] : LONER;
            for (uint32_t base = p_lo; base < p_hi; base += 32) {
                const uint32_t i = base + lane;
                const uint32_t v = v_next;
                v_next = i + 32 < p_hi ? work[i + 32] : LONER;
                const uint32_t bin = i < p_hi ? bin_of(v) : NONE;
                // (MATCH.ANY takes one step per distinct value; the lanes without an entry share one)
                const uint32_t peers = __match_any_sync(0xffffffffu, bin);
                if (bin != NONE) {
                    const uint32_t c = v & 0xFFFFu, ca = v >> 16;
                    const uint32_t pos = ms->binstart[bin] + cnt16[warp * CNT_STRIDE + bin] + __popc(peers & lt_mask);
                    const bool head = ca == c;
                    const bool find = !head || ((NONUNI[c >> 5] >> (c & 31)) & 1u);
                    la[pos] = i | (c << 16);
                    lb[pos] = (find ? 1u : 0u) | (head ? 2u : 0u);
                }
                __syncwarp();
                if (bin != NONE && (peers & lt_mask) == 0u) cnt16[warp * CNT_STRIDE + bin] += (uint16_t)__popc(peers);
                __syncwarp();
            }
        }
        __syncthreads();
        PHASE_STAMP3(5);

        // ---------------- P7: simulation. M = liveness bit per compact slot, T = position of the entry placed there
        for (uint32_t i = tid; i < SZ_BIG / 4; i += NTHREADS) reinterpret_cast<uint32_t*>(big)[i] = 0;
        __threadfence_block();
        __syncthreads();
        {
            const long long t_p7 = CLK3();
            const uint32_t lo = ms->binstart[tid], hi = ms->binstart[tid + 1];
            // ---- stage A: one lane per list, the short lists only (a long list would keep its whole warp waiting)
            const bool is_long = hi - lo > LONG_LIST_MIN;
            if (is_long) ms->longbin[atomicAdd(&ms->nlong, 1u)] = (uint16_t)tid;
            uint32_t i = is_long ? hi : lo, ei = lo;
            uint32_t hint_c = NONE, hint_w = 0;
            uint32_t a_n = i < hi ? la[i] : 0u, f_n = i < hi ? lb[i] : 0u;
            uint32_t st_walk = 0, st_exp = 0;
            while (i < hi) {
                const uint32_t a = a_n, f = f_n;
                if (i + 1 < hi) { a_n = la[i + 1]; f_n = lb[i + 1]; }
                const uint32_t p = a & 0xFFFFu, c = a >> 16;
                // FIFO expiry: the entry of position j is live at time p iff j + W >= p
                while (ei < i) {
                    const uint32_t x0 = lb[ei];
                    if ((x0 & 0xFFFFu) + W >= p) break;
                    const uint32_t s = x0 >> 16;
                    atomicAnd(&M[s >> 5], ~(1u << (s & 31)));
                    if (s >= hint_c && (s >> 5) < hint_w) hint_w = s >> 5;
                    ++ei;
                    if (DBG) ++st_exp;
                }
                if (f & 1u) {   // find: walk the live slots from the home until the pattern or a dead slot
                    const uint32_t w = sm_word(data, p);
                    uint32_t s = c, m = NONE;
                    for (;;) {
                        if (!((M[s >> 5] >> (s & 31)) & 1u)) break;
                        const uint32_t q = T[s];
                        if (sm_word(data, q) == w) { m = q; break; }
                        ++s;
                        if (DBG) ++st_walk;
                    }
                    cand[p] = candidate<V>(data, m, p);
                }
                // first fit: first dead slot at/after the home (never leaves the cluster)
                uint32_t wi, z;
                if (c == hint_c && hint_w > (c >> 5)) { wi = hint_w; z = ~M[wi]; }
                else { wi = c >> 5; z = ~M[wi] & (0xFFFFFFFFu << (c & 31)); }
                while (!z) { ++wi; z = ~M[wi]; }
                const uint32_t e = (wi << 5) + (uint32_t)(__ffs(z) - 1);
                if (f & 2u) { hint_c = c; hint_w = wi; }
                atomicOr(&M[e >> 5], 1u << (e & 31));
                T[e] = (uint16_t)p;
                lb[i] = p | (e << 16);
                ++i;
            }
            const long long t_a = CLK3();
            __syncthreads();     // longbin complete (the clusters of long lists are touched by nobody in stage A)
            // ---- stage B: the long lists (the chains of hot 4-grams), one WARP per list, pulled from a shared counter.
            // All 32 lanes run the same serial simulation on the same values; the lanes serve as the prefetch buffer:
            // 32 list entries (and 32 expiry records) are loaded with one coalesced access and handed out by shuffle,
            // so that no global-memory latency sits on the chain. Side effects by lane 0 only.
            {
                const uint32_t nlong = ms->nlong;
                for (;;) {
                    uint32_t li = 0;
                    if (lane == 0) li = atomicAdd(&ms->next_long, 1u);
                    li = __shfl_sync(0xffffffffu, li, 0);
                    if (li >= nlong) break;
                    const uint32_t bin = ms->longbin[li];
                    const uint32_t llo = ms->binstart[bin], lhi = ms->binstart[bin + 1];
                    uint32_t j = llo, ej = llo;
                    uint32_t ebase = llo;
                    uint32_t ea = ebase + lane < lhi ? la[ebase + lane] : 0u, ef = ebase + lane < lhi ? lb[ebase + lane] : 0u;
                    uint32_t xbase = llo, xcount = 0, xr = 0;        // records of the entries xbase .. xbase + xcount - 1
                    // (hc, hwd, hw): every word of M in [hc >> 5, hwd) has no dead slot at/after hc; hw = this list's view of M[hwd]
                    uint32_t hc = NONE, hwd = 0, hw = 0;
                    while (j < lhi) {
                        if (j - ebase == 32) { ebase += 32; ea = ebase + lane < lhi ? la[ebase + lane] : 0u; ef = ebase + lane < lhi ? lb[ebase + lane] : 0u; }
                        const uint32_t a = __shfl_sync(0xffffffffu, ea, j - ebase), f = __shfl_sync(0xffffffffu, ef, j - ebase);
                        const uint32_t p = a & 0xFFFFu, c = a >> 16;
                        while (ej < j) {
                            if (ej - xbase >= xcount) {      // next batch of records: only entries that are placed (index < j)
                                __syncwarp();
                                xbase = ej; xcount = j - ej < 32 ? j - ej : 32;
                                xr = lane < xcount ? *reinterpret_cast<volatile uint32_t*>(&lb[xbase + lane]) : 0u;
                            }
                            const uint32_t x0 = __shfl_sync(0xffffffffu, xr, ej - xbase);
                            if ((x0 & 0xFFFFu) + W >= p) break;
                            const uint32_t s = x0 >> 16, sw = s >> 5, sbit = 1u << (s & 31);
                            if (lane == 0) atomicAnd(&M[sw], ~sbit);
                            if (s >= hc) {
                                if (sw < hwd) { hwd = sw; hw = ~sbit; }   // that word had no dead slot at/after hc: now exactly this one
                                else if (sw == hwd) hw &= ~sbit;
                            }
                            ++ej;
                        }
                        __syncwarp();
                        if (f & 1u) {   // find (uniform: every lane reads the same words)
                            const uint32_t w = sm_word(data, p);
                            uint32_t s = c, m = NONE;
                            for (;;) {
                                if (!((*reinterpret_cast<volatile uint32_t*>(&M[s >> 5]) >> (s & 31)) & 1u)) break;
                                const uint32_t q = *reinterpret_cast<volatile uint16_t*>(&T[s]);
                                if (sm_word(data, q) == w) { m = q; break; }
                                ++s;
                            }
                            if (lane == 0) cand[p] = candidate<V>(data, m, p);
                        }
                        uint32_t wi, z;
                        if (c == hc) { wi = hwd; z = ~hw; if (wi == (c >> 5)) z &= 0xFFFFFFFFu << (c & 31); }
                        else { wi = c >> 5; z = ~*reinterpret_cast<volatile uint32_t*>(&M[wi]) & (0xFFFFFFFFu << (c & 31)); }
                        while (!z) { ++wi; z = ~*reinterpret_cast<volatile uint32_t*>(&M[wi]); }
                        const uint32_t e = (wi << 5) + (uint32_t)(__ffs(z) - 1);
                        const uint32_t ebit = 1u << (e & 31);
                        if (f & 2u) {                       // a head: its home keys the hint from now on
                            if (c != hc || wi != hwd) { hw = ~z; if (wi == (c >> 5)) hw |= ~(0xFFFFFFFFu << (c & 31)); }
                            hc = c; hwd = wi; hw |= ebit;
                        } else if (e >= hc && (e >> 5) == hwd) hw |= ebit;
                        if (lane == 0) {
                            atomicOr(&M[e >> 5], ebit);
                            T[e] = (uint16_t)p;
                            lb[j] = p | (e << 16);
                        }
                        ++j;
                    }
                    __syncwarp();
                }
            }
            (void)t_a;
            // the slot-0 / table-end cluster: serial, exact reference order, T = position + 1 with lazy expiry (0 = never
            // used; its compact slots [0, sp_lo_end) and [sp_hi_start, nslots) are touched by nobody else)
            if (tid == 0) {
                const uint32_t slo = ms->binstart[NBIN], shi = ms->binstart[NBIN + 1];
                uint32_t qh = 0, qt = 0;
                ms->clr[0] = W - 1; qt = 1;
                for (uint32_t j = slo; j < shi; ++j) {
                    const uint32_t av = la[j], q0 = av & 0xFFFFu, c0 = av >> 16;
                    auto clear0 = [&]() { if (sp_lo_end) T[0] = 0; };
                    while (qh < qt && ms->clr[qh & 63] < q0) { clear0(); ++qh; }
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
                    uint32_t e = ran_off ? 0 : k;                          // the wrapping insert continues at slot 0
                    for (;;) { if (T[e] <= dthr) break; ++e; if (e == nslots) e = 0; }
                    if (q0 != 65535u) T[e] = (uint16_t)(q0 + 1);
                    if (e == 0 && sp_lo_end) { ms->clr[qt & 63] = q0 + W; ++qt; }
                    if (qh < qt && ms->clr[qh & 63] == q0) { clear0(); ++qh; }
                    cand[q0] = candidate<V>(data, m, q0);
                }
            }
            if (DBG && dbg_stats && lane == 0) {
                uint32_t* o = dbg_stats + (uint64_t)b * 136 + 8 + warp * 4;
                o[0] = (uint32_t)(clock64() - t_p7); o[1] = hi - lo; o[2] = (uint32_t)(t_a - t_p7); o[3] = ms->nlong;
            }
        }
        __threadfence_block();
        __syncthreads();
        PHASE_STAMP3(6);

        // ---------------- P8: greedy parse. adv[p] = bytes consumed by the token that would start at p
        for (uint32_t i0 = tid; i0 < len; i0 += 8 * NTHREADS) {
            uint32_t tv[8];
#pragma unroll
            for (uint32_t k = 0; k < 8; ++k) { const uint32_t i = i0 + k * NTHREADS; tv[k] = i < len ? cand[i] : 0u; }
#pragma unroll
            for (uint32_t k = 0; k < 8; ++k) {
                const uint32_t i = i0 + k * NTHREADS;
                if (i < len) {
                    if (dbg_tok) dbg_tok[(uint64_t)b * MAXB + i] = tv[k];
                    adv[PADX(i)] = (uint8_t)((tv[k] >> 16) ? (tv[k] >> 16) : 1u);
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
            for (uint32_t s = 0; s < 32; ++s) { ms->sentry[s] = (uint8_t)e; e = ms->sexit[s][e & 31u]; }
        }
        __syncthreads();
        uint8_t* centry = reinterpret_cast<uint8_t*>(pre) + 8192;   // u8[1024] chunk entry offsets (clear of the V0 staging spill)
        if (lane == 0) {
            uint32_t e = ms->sentry[warp];
            for (uint32_t k = 0; k < 32; ++k) {
                const uint32_t ch = warp * 32 + k;
                centry[ch] = (uint8_t)e;
                const uint32_t p = (ch << 6) + e;
                if (ch < nchunks && p < len) e = exitof[PADX(p)];
            }
        }
        __syncthreads();
        // per-chunk output size, CTA exclusive scan. V1 also notes, for every token start, its offset inside the
        // chunk's output (in 2-byte units, exitof is dead now) so that P9 emits one position per thread
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
        uint8_t* out = scratch + (uint64_t)b * stride;

        // ---------------- P9: emission
        if (V == 1) {
            uint32_t* coff = pre + 3072;          // u32[1024] output offset of every chunk
            if (tid < nchunks) coff[tid] = my_off;
            __syncthreads();
            for (uint32_t p0 = tid; p0 < len; p0 += 4 * NTHREADS) {
                uint32_t tk[4];
#pragma unroll
                for (uint32_t k = 0; k < 4; ++k) { const uint32_t p = p0 + k * NTHREADS; tk[k] = p < len ? cand[p] : 0u; }
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
            if (tid == 0) { block_sizes[b] = total_units; block_bytes[b] = total_units; }
        } else {
            // LSB-first bit stream staged in shared memory (exitof is dead now), then stored coalesced
            const uint32_t nwords = (total_units >> 5) + 2;
            __syncthreads();
            for (uint32_t i = tid; i < nwords; i += NTHREADS) stage[i] = 0;
            __syncthreads();
            if (tid < nchunks && my_units) {
                const uint32_t cend = (tid << 6) + 64;
                const uint32_t hi = cend < len ? cend : len;
                uint32_t wi = my_off >> 5, have = my_off & 31;
                bool first = have != 0;
                uint64_t acc = 0;
                for (uint32_t p = (tid << 6) + centry[tid]; p < hi; p += adv[PADX(p)]) {
                    const uint32_t t = cand[p];
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
            for (uint32_t i = tid; i < nwords; i += NTHREADS) reinterpret_cast<uint32_t*>(out)[i] = stage[i];
            if (tid == 0) { block_sizes[b] = total_units; block_bytes[b] = total_units / 8 + 1; }
        }
        __syncthreads();   // smem is reused by the next block
        PHASE_STAMP3(7);
    }
}

}  // namespace

// scratch slots: 13 = work, 14 = cand (shared with v2's lists / tok), 16 + 17 = lane lists
int lz77_v3_launch(b200_ctx* ctx, int variant, const uint8_t* d_in, uint64_t n, uint64_t bs, uint64_t nblocks,
                   uint8_t* scratch, uint64_t stride, uint64_t* d_block_sizes, uint64_t* block_bytes, uint32_t* dbg_tok) {
    static bool attr_done[64] = {};
    const int dev = ctx->device >= 0 && ctx->device < 64 ? ctx->device : 0;
    if (!attr_done[dev]) {   // the attribute is per device
        CUDA_TRY(cudaFuncSetAttribute(lz77_v3_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        CUDA_TRY(cudaFuncSetAttribute(lz77_v3_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        CUDA_TRY(cudaFuncSetAttribute(lz77_v3_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        CUDA_TRY(cudaFuncSetAttribute(lz77_v3_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        attr_done[dev] = true;
    }
    uint64_t grid = (uint64_t)ctx->sm_count;
    if (grid > nblocks) grid = nblocks;
    uint32_t *work, *cand, *la, *lb;
    B200_TRY(b200_scratch(ctx, B200_SLOT(ctx, 13), (size_t)grid * MAXB * 4 + 64, reinterpret_cast<void**>(&work)));
    B200_TRY(b200_scratch(ctx, B200_SLOT(ctx, 14), (size_t)grid * MAXB * 4 + 64, reinterpret_cast<void**>(&cand)));
    B200_TRY(b200_scratch(ctx, B200_SLOT(ctx, 16), (size_t)grid * MAXB * 4 + 64, reinterpret_cast<void**>(&la)));
    B200_TRY(b200_scratch(ctx, B200_SLOT(ctx, 17), (size_t)grid * MAXB * 4 + 64, reinterpret_cast<void**>(&lb)));
#define LZ_V3_LAUNCH(V, D) lz77_v3_kernel<V, D><<<(unsigned)grid, NTHREADS, SMEM_BYTES, ctx->stream>>>(d_in, n, (uint32_t)bs, (uint32_t)nblocks, work, cand, la, lb, scratch, stride, d_block_sizes, block_bytes, dbg_tok)
    if (dbg_tok) { if (variant == 0) LZ_V3_LAUNCH(0, true); else LZ_V3_LAUNCH(1, true); }
    else { if (variant == 0) LZ_V3_LAUNCH(0, false); else LZ_V3_LAUNCH(1, false); }
    CUDA_TRY(cudaGetLastError());
    return B200_OK;
}
