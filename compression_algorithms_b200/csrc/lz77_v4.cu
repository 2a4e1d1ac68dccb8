// LZ77 v4: the deflate-variant match finder (algorithms/deflate/lz77.c:44-174,199-280) for blocks of at most
// 65536 bytes, slot-exact like v2 (lz77_v2.cu) but without a time-ordered table simulation.
//
// What makes it possible (DESIGN.md "LZ77 v4"):
//  * A block of len <= 2 W positions has only TWO expiry phases. Entry i is live at time P iff i + W >= P
//    (deflate/lz77.c:120-139 as lazy expiry), so while P <= W nothing is dead ("phase 1", entries 0..W), and an
//    entry placed at P > W ("phase 2") outlives the block. Every slot therefore has at most two occupants:
//    e1 (placed in phase 1, dead from e1 + W + 1 on) and e2 (placed in phase 2, never dead).
//  * In the no-expiry occupancy (built with atomicOr probing exactly as in v2) clusters never interact, a
//    cluster has as many entries as slots, and the atomicOr of P1 hands every entry a private slot of ITS OWN
//    cluster. Scattering the entries to the rank of that slot groups them by cluster without any sort or partition.
//  * Inside a cluster the sequential first-fit is a slot sweep: slot j goes to the EARLIEST entry that is still
//    unplaced, whose home is at/before j and for which j is dead (phase 1: always; phase 2: arrival >= release
//    time of e1[j]). Between two consecutive homes the candidates are a fixed time-ordered pool, so phase 1
//    hands the next slots to the next pool entries, and phase 2 is the recurrence
//    idx_j = max(idx_{j-1} + 1, lower_bound(pool, release_j)) = one prefix maximum (release times rise along a
//    run of phase-1 slots because those were filled in time order).
//  * find(p) only ever looks at the slots [home(p), slot(p)): all live at time p by construction, so the find is a
//    bounded scan for the first occupant (e1 if still live, else e2) with p's pattern. No liveness test, no table.
//  * Clusters of up to 16 entries (95 % of the non-trivial ones) are simulated by ONE lane in registers (liveness mask + FIFO
//    expiry pointer; two entries in closed form), 17..64 by a lane after a warp sort, 65..256 by a warp and larger ones by a
//    team of four warps with the sweeps above; a cluster of one entry (a third of text positions) is a literal candidate
//    and costs nothing.
//  * The clusters on slot 0 and on the table's last slot (the reference's early slot-0 clear, U10, and its wrapping insert) are
//    simulated by one thread with the reference's own rules when they have at most 60 slots together. Blocks this kernel does
//    not take (larger special clusters, a cluster above 16383 entries, more than 512 clusters above 16 entries) are listed and
//    run through lz77_v2_kernel.
//
// Phases per block (one persistent CTA of 1024 threads per SM):
//   P0 load | P1 occupancy bitmap, claimed slot per position | P2 rank prefix, cluster-start flags
//   P3 rank of the claimed slot -> (compact slot, displacement) per position; loners marked; the special clusters
//   lane stage: every entry's position (u16) to its compact slot in shared memory, work list by size class, clusters of up to
//   16 entries one lane each | final stage: the larger clusters, all at once (lane batches, warps, four-warp teams)
//   -> F(p) (u16, global, L2 resident)
//   P5 token candidates (match extension) + greedy parse | P6 token emission   (as v2)
#include "common.cuh"
#include "../../include/b200comp.h"

namespace {

constexpr uint32_t SLOTS = 1u << 20;
constexpr uint32_t GUARD_BITS = 65536u;
constexpr uint32_t BM_WORDS = (SLOTS + GUARD_BITS) / 32;     // 34816
constexpr uint32_t PRE_CHUNK = 8;
constexpr uint32_t PRE_N = BM_WORDS / PRE_CHUNK;             // 4352
constexpr uint32_t LONER = 0xFFFFFFFFu;
constexpr uint32_t NONE16 = 0xFFFFu;
constexpr uint32_t MAXB = 65536;
constexpr uint32_t NTHREADS = 1024;
constexpr uint32_t W = 1u << 15, MAXLEN = 31;
constexpr uint32_t CH = 17408;                               // compact slots of the final stage (entries + two occupants per slot)
constexpr uint32_t TMIN = 256;                               // clusters above this: four-warp teams; LMAX+1 .. TMIN: one warp
constexpr uint32_t L2MAX = 64;                               // final stage: clusters of LMAX+1 .. L2MAX entries are sorted by the warp and simulated one lane each
constexpr uint32_t LMAX = 16;                                // largest cluster simulated by one lane (above 16: sorted by the warp first)
constexpr uint32_t NBIG = 512;                               // clusters above LMAX entries per block (kept for the final stage)
constexpr uint32_t WL_S16 = MAXB * 2;                           // lane stage: bare positions by compact slot (u16), the whole block
constexpr uint32_t WL_PRE = 7680;                             // work-list entries inside the PRE region (the big-cluster list follows)
constexpr uint32_t WL_CAP = WL_PRE + (BM_WORDS * 4 - WL_S16) / 2;   // ... plus the tail of BIG right before it; the rest spills to global memory
constexpr uint32_t CL_MAX = 16383;                           // largest cluster (home offsets are 14 bits)
constexpr uint32_t PLACED = 0x80000000u, ISHOME = 0x40000000u;

constexpr uint32_t OFF_DATA = 0;
constexpr uint32_t SZ_DATA = MAXB + 128;
constexpr uint32_t OFF_BIG = OFF_DATA + SZ_DATA;
constexpr uint32_t SZ_BIG = BM_WORDS * 4;                    // 139264: bitmap | S + E1 + E2 | adv + exit
constexpr uint32_t OFF_PRE = OFF_BIG + SZ_BIG;
constexpr uint32_t SZ_PRE = PRE_N * 4 + 16;                  // rank prefix (u16 per 4 words) | work queues | chunk entry / offsets
constexpr uint32_t OFF_FLAGS = OFF_PRE + SZ_PRE;
constexpr uint32_t SZ_FLAGS = 8192 + 64;                     // one bit per compact slot: starts a cluster (+ sentinel)
constexpr uint32_t OFF_MISC = OFF_FLAGS + SZ_FLAGS;
constexpr uint32_t SZ_MISC = 1792;
constexpr uint32_t SMEM_BYTES = OFF_MISC + SZ_MISC;
static_assert(SMEM_BYTES <= 232448, "shared memory layout too large");
static_assert(CH * 8 <= SZ_BIG, "chunk does not fit");
static_assert(8192 + 4 * NBIG * 2 <= WL_PRE * 2 && WL_PRE * 2 + NBIG * 4 <= SZ_PRE && 4096 + 32 * 32 * 4 <= WL_PRE * 2 && OFF_PRE == OFF_BIG + SZ_BIG, "work list + big-cluster list / rank prefix + warp scratch do not fit");
static_assert(WL_S16 <= SZ_BIG, "lane-stage positions do not fit");

constexpr uint32_t PADDED = MAXB + (MAXB >> 6) * 4;          // 69632
#define PADX(p) ((p) + (((p) >> 6) << 2))

struct Misc {
    uint32_t scan[34];
    uint32_t p1_next, fallback, ce, pad0;
    uint32_t ccnt[7], cbase[8], cfill[7], next_t, next_w, next_l;   // work list by size class: 17..64 | 9..16 | 5..8 | 3..4 | 2
    uint32_t nbig, r0, r1, rdone;                                   // big-cluster list; the round of the final stage
    uint32_t cut0, top_start, nsp, sp_pad;                          // the clusters on slot 0 / the table end: raw bounds, their entries
    uint32_t tpick[8];
    uint32_t ncls[4];            // final stage: listed clusters of the round by size class (teams | one warp, 129+ | one warp | lane batches)
    uint32_t tscr[8][12];        // per four-warp team: [2][4] cross-warp scan scratch, words 8 - 10 = progress / hand-over of the pipelined phases
    uint8_t  sexit[32][32];
    uint8_t  sentry[36];
};
static_assert(sizeof(Misc) <= SZ_MISC, "misc region too small");

__device__ __forceinline__ uint32_t sm_word(const uint8_t* data, uint32_t p) {
    const uint32_t* a = reinterpret_cast<const uint32_t*>(data + (p & ~3u));
    return __funnelshift_r(a[0], a[1], (p & 3u) * 8);
}

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

// position of the n-th set bit of x (n = 0 .. popc(x) - 1): five popcount steps (__fns walks bit by bit)
__device__ __forceinline__ uint32_t nth_set(uint32_t x, uint32_t n) {
    uint32_t pos = 0, c;
    c = __popc(x & 0xFFFFu); if (n >= c) { n -= c; pos = 16; x >>= 16; }
    c = __popc(x & 0xFFu);   if (n >= c) { n -= c; pos += 8; x >>= 8; }
    c = __popc(x & 0xFu);    if (n >= c) { n -= c; pos += 4; x >>= 4; }
    c = __popc(x & 0x3u);    if (n >= c) { n -= c; pos += 2; x >>= 2; }
    if (n >= (x & 1u)) pos += 1;
    return pos;
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

// ---- clusters of 2 .. LMAX entries: ONE LANE per cluster, general FIFO expiry (any number of phases).
// S[0 .. m): entry = position | compact home << 16, sorted by position; cstart = compact slot of the cluster's first slot.
// PACKED (m <= 16): occupant of every slot / slot of every entry as 4-bit fields of two 64-bit registers;
// otherwise (m <= 64) in the cluster's own E1 / E2 words.
template <bool PACKED, typename ST>
__device__ __forceinline__ void lane_cluster(const ST* S, uint16_t* E1x, uint16_t* E2x, uint32_t m, uint32_t cstart,
                                             const uint8_t* data, uint16_t* fres) {
    // PACKED: S holds bare positions (u16) and cstart is the RAW slot of the cluster's first slot (lane_sort16), the home
    // offset of an entry is its hash minus that; otherwise entry = position | compact home << 16, cstart = compact start.
    unsigned long long live = 0, slotpack = 0, occpack = 0;
    uint32_t ex = 0;
    for (uint32_t i = 0; i < m; ++i) {
        const uint32_t e = S[i], p = e & 0xFFFFu;
        const uint32_t w = sm_word(data, p);
        const uint32_t o = PACKED ? lz_hash(w) - cstart : (e >> 16) - cstart;
        while (ex < i) {                                     // entries expire in the order they came
            const uint32_t tex = S[ex] & 0xFFFFu;
            if (tex + W >= p) break;
            const uint32_t sl = PACKED ? (uint32_t)(slotpack >> (4 * ex)) & 15u : (uint32_t)E2x[ex];
            live &= ~(1ull << sl);
            ++ex;
        }
        const uint32_t dead = o + (uint32_t)(__ffsll((long long)(~live >> o)) - 1);   // first dead slot at/after the home: p's own slot
        uint32_t f = NONE16;
        for (uint32_t j = o; j < dead; ++j) {
            const uint32_t y = PACKED ? (uint32_t)(occpack >> (4 * j)) & 15u : (uint32_t)E1x[j];
            const uint32_t ty = S[y] & 0xFFFFu;
            if (sm_word(data, ty) == w) { f = ty; break; }
        }
        fres[p] = (uint16_t)f;
        live |= 1ull << dead;
        if (PACKED) {
            occpack = (occpack & ~(15ull << (4 * dead))) | ((unsigned long long)i << (4 * dead));
            slotpack |= (unsigned long long)dead << (4 * i);
        } else { E1x[dead] = (uint16_t)i; E2x[i] = (uint16_t)dead; }
    }
}

// insertion sort of bare positions (m <= 16); returns the raw slot the cluster starts on = the hash of the entry that sits in
// its first slot (nothing is occupied below it, so that entry cannot have probed its way there), read before the sort
__device__ __forceinline__ uint32_t lane_sort16(uint16_t* S, uint32_t m, const uint8_t* data) {
    const uint32_t hmin = lz_hash(sm_word(data, S[0]));
    for (uint32_t i = 1; i < m; ++i) {
        const uint32_t e = S[i];
        uint32_t j = i;
        while (j > 0) { const uint32_t f = S[j - 1]; if (f <= e) break; S[j] = (uint16_t)f; --j; }
        S[j] = (uint16_t)e;
    }
    return hmin;
}

// one warp sorts up to 64 entries by position: bitonic network over two registers per lane
// (the compact homes become offsets inside the cluster, so that no key equals the padding)
__device__ __forceinline__ void warp_sort64(uint32_t* Sx, uint32_t m, uint32_t cstart) {
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t x0 = 0xFFFFFFFFu, x1 = 0xFFFFFFFFu;                                      // position in the high half
    if (lane < m) { const uint32_t e = Sx[lane]; x0 = (e << 16) | ((e >> 16) - cstart); }
    if (lane + 32 < m) { const uint32_t e = Sx[lane + 32]; x1 = (e << 16) | ((e >> 16) - cstart); }
#pragma unroll
    for (uint32_t k = 2; k <= 64; k <<= 1) {
        if (k == 64) { const uint32_t lo = min(x0, x1), hi = max(x0, x1); x0 = lo; x1 = hi; }
#pragma unroll
        for (uint32_t j = (k == 64 ? 16 : k >> 1); j > 0; j >>= 1) {
            const uint32_t y0 = __shfl_xor_sync(0xffffffffu, x0, j), y1 = __shfl_xor_sync(0xffffffffu, x1, j);
            const bool lower = (lane & j) == 0;
            const bool up0 = k < 32 ? (lane & k) == 0 : true;            // k == 32: register 0 ascending, register 1 descending
            const bool up1 = k < 32 ? (lane & k) == 0 : (k == 64);
            x0 = (lower == up0) ? min(x0, y0) : max(x0, y0);
            x1 = (lower == up1) ? min(x1, y1) : max(x1, y1);
        }
    }
    if (lane < m) Sx[lane] = __byte_perm(x0, 0, 0x1032);
    if (lane + 32 < m) Sx[lane + 32] = __byte_perm(x1, 0, 0x1032);
}

// ---- team primitives (TEAM = 32: one warp; TEAM = 128: four warps on a named barrier)
template <int TEAM> struct Team {
    uint32_t tt;        // thread in team
    uint32_t bar;       // named barrier (TEAM == 128)
    uint32_t* scr;      // [2][4] cross-warp scratch (TEAM == 128)
    uint32_t par;
    __device__ __forceinline__ void sync() {
        if (TEAM == 32) __syncwarp();
        else asm volatile("bar.sync %0, %1;" ::"r"(bar), "r"(TEAM) : "memory");
    }
    __device__ __forceinline__ uint32_t exscan(uint32_t v, uint32_t& tot) {
        const uint32_t incl = warp_incl_scan_u32(v);
        if (TEAM == 32) { tot = __shfl_sync(0xffffffffu, incl, 31); return incl - v; }
        uint32_t* s = scr + par * 4; par ^= 1u;
        if ((tt & 31u) == 31u) s[tt >> 5] = incl;
        sync();
        uint32_t off = 0, t = 0;
#pragma unroll
        for (uint32_t w = 0; w < 4; ++w) { const uint32_t x = s[w]; if (w < (tt >> 5)) off += x; t += x; }
        tot = t;
        return off + incl - v;
    }
    // inclusive running maximum; last = maximum over the whole team
    __device__ __forceinline__ int maxscan(int v, int& last) {
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, v, d); if ((int)(tt & 31u) >= d) v = max(v, t); }
        if (TEAM == 32) { last = __shfl_sync(0xffffffffu, v, 31); return v; }
        int* s = reinterpret_cast<int*>(scr + par * 4); par ^= 1u;
        if ((tt & 31u) == 31u) s[tt >> 5] = v;
        sync();
        int pre = -(1 << 30), t = -(1 << 30);
#pragma unroll
        for (uint32_t w = 0; w < 4; ++w) { const int x = s[w]; if (w < (tt >> 5)) pre = max(pre, x); t = max(t, x); }
        last = t;
        return max(v, pre);
    }
    __device__ __forceinline__ uint32_t tmin(uint32_t v) {
        v = __reduce_min_sync(0xffffffffu, v);
        if (TEAM == 32) return v;
        uint32_t* s = scr + par * 4; par ^= 1u;
        if ((tt & 31u) == 0u) s[tt >> 5] = v;
        sync();
        return min(min(s[0], s[1]), min(s[2], s[3]));
    }
};

// ---- clusters above LMAX entries: slot sweeps, two expiry phases.
// Sx[i]: entry i = position | home offset << 16 (14 bits) | PLACED; bit ISHOME of Sx[j]: slot j is some entry's home.
// E1x[j] / E2x[j]: index of the phase-1 / phase-2 occupant of slot j (NONE16 = none). E2x[lo ..] also holds the
// candidate pool of the sub-segment being swept (a slot's final value overwrites a pool entry no later slot needs).
template <int TEAM>
__device__ void big_cluster(Team<TEAM>& T, uint32_t* Sx, uint16_t* E1x, uint16_t* E2x, uint32_t m, uint32_t cstart,
                            const uint8_t* data, uint16_t* fres, uint32_t* scr, uint32_t* tdbg = nullptr) {
    const uint32_t tt = T.tt;
    long long t_m = tdbg ? clock64() : 0;
    const long long t_in = t_m;
#define BC_MARK(k, who) do { if (tdbg && tt == (who)) atomicAdd(&tdbg[k], (uint32_t)((clock64() - t_in) >> 6)); } while (0)
#define BC_STAMP(k) do { if (tdbg && tt == 0) { const long long t_n = clock64(); atomicAdd(&tdbg[k], (uint32_t)((t_n - t_m) >> 6)); t_m = t_n; } } while (0)
    for (uint32_t i = tt; i < m; i += TEAM) { const uint32_t e = Sx[i]; Sx[i] = (e & 0xFFFFu) | (((e >> 16) - cstart) << 16); }
    T.sync();
    // bitonic sort by position (all comparators ascending, so the virtual +inf padding behind m never moves)
    uint32_t P = 2; while (P < m) P <<= 1;
    if (TEAM == 32 && m <= 64) { warp_sort64(Sx, m, 0u); __syncwarp(); P = 1; }   // two registers per lane instead of shared memory
    if (TEAM == 128) {
        // a team sorts by counting: 128 buckets of 512 positions (thread b owns bucket b), entries scattered into the cluster's
        // still unused occupant arrays (position -> E1x, home -> E2x), then every thread sorts its own bucket back into Sx.
        // Positions are distinct, so a bucket holds a handful of entries unless the cluster is a run of consecutive
        // positions: above 24 entries in a bucket the bitonic network below runs instead.
        uint32_t* C = scr - 32u * (tt >> 5);                  // the team's four scratch rows: 128 words, zero on entry and exit
        for (uint32_t i = tt; i < m; i += TEAM) atomicAdd(&C[(Sx[i] & 0xFFFFu) >> 9], 1u);
        T.sync();
        const uint32_t mycount = C[tt];
        uint32_t tot;
        const uint32_t mystart = T.exscan(mycount, tot);
        const uint32_t big = T.tmin(mycount <= 24u ? 1u : 0u);   // 0 if any bucket is too full (tmin ends with every thread past its reads)
        T.sync();
        C[tt] = big ? mystart : 0u;
        T.sync();
        if (big) {
            for (uint32_t i = tt; i < m; i += TEAM) {
                const uint32_t e = Sx[i];
                const uint32_t pos = atomicAdd(&C[(e & 0xFFFFu) >> 9], 1u);
                E1x[pos] = (uint16_t)e; E2x[pos] = (uint16_t)(e >> 16);
            }
            T.sync();
            for (uint32_t a = 0; a < mycount; ++a) {          // insertion sort of my bucket, written back in order
                const uint32_t t0 = E1x[mystart + a], h0 = E2x[mystart + a];
                uint32_t j = a;
                while (j > 0 && (uint32_t)E1x[mystart + j - 1] > t0) { E1x[mystart + j] = E1x[mystart + j - 1]; E2x[mystart + j] = E2x[mystart + j - 1]; --j; }
                E1x[mystart + j] = (uint16_t)t0; E2x[mystart + j] = (uint16_t)h0;
            }
            for (uint32_t a = 0; a < mycount; ++a) Sx[mystart + a] = (uint32_t)E1x[mystart + a] | ((uint32_t)E2x[mystart + a] << 16);
            P = 1;
        }
        C[tt] = 0;
        T.sync();
    }
    for (uint32_t k = 2; k <= P; k <<= 1) {
        const uint32_t half = k >> 1;
        for (uint32_t idx = tt; idx < (P >> 1); idx += TEAM) {
            const uint32_t l = ((idx & ~(half - 1u)) << 1) | (idx & (half - 1u));
            const uint32_t r = l ^ (k - 1u);
            if (r < m) { const uint32_t a = Sx[l], b = Sx[r]; if ((a & 0xFFFFu) > (b & 0xFFFFu)) { Sx[l] = b; Sx[r] = a; } }
        }
        T.sync();
        for (uint32_t j = k >> 2; j >= 1; j >>= 1) {
            for (uint32_t idx = tt; idx < (P >> 1); idx += TEAM) {
                const uint32_t l = ((idx & ~(j - 1u)) << 1) | (idx & (j - 1u));
                const uint32_t r = l + j;
                if (r < m) { const uint32_t a = Sx[l], b = Sx[r]; if ((a & 0xFFFFu) > (b & 0xFFFFu)) { Sx[l] = b; Sx[r] = a; } }
            }
            T.sync();
        }
    }
    BC_STAMP(0);
    for (uint32_t i = tt; i < m; i += TEAM) atomicOr(&Sx[(Sx[i] >> 16) & 0x3FFFu], ISHOME);
    T.sync();
    uint32_t n1;                                             // phase-1 entries: positions 0 .. W
    { uint32_t a = 0, b = m; while (a < b) { const uint32_t mid = (a + b) >> 1; if ((Sx[mid] & 0xFFFFu) <= W) a = mid + 1; else b = mid; } n1 = a; }
    auto next_home = [&](uint32_t x) -> uint32_t {
        for (uint32_t base = x + 1; base < m; base += TEAM) {
            const uint32_t j = base + tt;
            const uint32_t v = (j < m && (Sx[j] & ISHOME)) ? j : m;
            const uint32_t r = T.tmin(v);
            if (r < m) return r;
        }
        return m;
    };
    // find of the entry at position p (pattern of p, home o) that got slot j: first occupant of [o, j) with p's pattern.
    // Every slot of that range is live at time p (phase 1: by an earlier phase-1 entry; phase 2: by e1 if it has not
    // expired, else by e2). A few slots by the lane itself, the rest of a long range 32 slots per step by its warp.
    auto occ_pos = [&](uint32_t jj, uint32_t p, bool ph2) -> uint32_t {
        const uint32_t y1 = E1x[jj];
        if (!ph2) return Sx[y1] & 0xFFFFu;
        if (y1 != NONE16) { const uint32_t ty = Sx[y1] & 0xFFFFu; if (ty + W >= p) return ty; }
        const uint32_t y2 = E2x[jj];
        return y2 == NONE16 ? NONE16 : (Sx[y2] & 0xFFFFu);   // (a live slot always has an occupant)
    };
    auto find = [&](bool active, uint32_t o, uint32_t j, uint32_t p, bool ph2) -> uint32_t {
        uint32_t f = NONE16, cur = o, w = 0;
        if (active && o < j) {
            w = sm_word(data, p);
            const uint32_t lim = o + 6 < j ? o + 6 : j;
            for (; cur < lim; ++cur) {
                const uint32_t ty = occ_pos(cur, p, ph2);
                if (ty != NONE16 && sm_word(data, ty) == w) { f = ty; break; }
            }
        }
        uint32_t lm = __ballot_sync(0xffffffffu, active && f == NONE16 && cur < j);
        while (lm) {
            const int srcl = __ffs(lm) - 1;
            lm &= lm - 1;
            const uint32_t bc = __shfl_sync(0xffffffffu, cur, srcl), bj = __shfl_sync(0xffffffffu, j, srcl);
            const uint32_t bp = __shfl_sync(0xffffffffu, p, srcl), bw = __shfl_sync(0xffffffffu, w, srcl);
            uint32_t res = NONE16;
            for (uint32_t base = bc; base < bj; base += 32) {
                const uint32_t sl = base + (tt & 31u);
                uint32_t ty = NONE16;
                if (sl < bj) ty = occ_pos(sl, bp, ph2);
                const uint32_t hm = __ballot_sync(0xffffffffu, ty != NONE16 && sm_word(data, ty) == bw);
                if (hm) { res = __shfl_sync(0xffffffffu, ty, __ffs(hm) - 1); break; }
            }
            if ((int)(tt & 31u) == srcl) f = res;
        }
        return f;
    };
    // ---- clusters whose entries of either phase fit 32 x 32 bits: ONE warp, the candidate pool as a bit mask per lane
    // (lane l: entries 32 l .. 32 l + 31 of the phase, in time order). "eligible" = unplaced and home at/before the segment;
    // a sub-segment ranks / selects in the mask as it was at its start (the pool of the recurrence is fixed), the slots it
    // hands out are cleared through 32 words of scratch. No compaction, no barriers.
    if (n1 <= 1024u && m - n1 <= 1024u) {
        const uint32_t lane = tt & 31u;
        // TEAM == 128: warp 0 sweeps phase 1 and publishes how far its slots are final, warp 1 sweeps phase 2 right behind it
        // (a sub-segment of phase 2 needs the phase-1 occupants of its own slots only); the whole team does the finds.
        volatile uint32_t* prog = (TEAM == 32) ? nullptr : reinterpret_cast<volatile uint32_t*>(T.scr + 8);
        volatile uint32_t* progg = (TEAM == 32) ? nullptr : reinterpret_cast<volatile uint32_t*>(T.scr + 9);
        if (TEAM != 32) { if (tt == 0) { *prog = 0; *progg = 0; } T.sync(); }
        // ---- phase 1 in TIME order for a team cluster (checked against the sweep on the CPU, tools/proto_v4.c). Nothing dies in
        // phase 1, so the table is: T = the lowest free slot, plus the few slots taken AHEAD of it. An entry whose home is at or
        // before T takes T ("chain-like": every home-0 entry, and most others once the chain has grown past their home); the k-th
        // chain-like entry therefore sits in the k-th slot that is not an ahead slot. An entry whose home is beyond T takes the
        // first slot from its home that no earlier ahead entry holds. Only the entries with a home other than the cluster's first
        // slot need a look (one warp, in time order, <= 63 of them or the sweep below runs instead); everything else is a scatter.
        bool p1_closed = false;
        if (TEAM != 32 && n1 > 0u) {
            uint16_t* IL = E2x;             // entries with home != 0, in time order
            uint16_t* AH = E2x + 64;        // ahead slots, ascending
            uint16_t* BI = E2x + 128;       // the entries that took them (ascending index = time order)
            uint16_t* BS = E2x + 192;       // and which slot each took
            volatile uint32_t* p1info = reinterpret_cast<volatile uint32_t*>(T.scr + 10);   // na | nB << 8 | ok << 31
            if (tt < 32u) {
                uint32_t q = 0;
                for (uint32_t base = 0; base < n1; base += 32) {
                    const uint32_t i = base + lane;
                    const bool isI = i < n1 && ((Sx[i] >> 16) & 0x3FFFu) != 0u;
                    const uint32_t mk = __ballot_sync(0xffffffffu, isI);
                    if (isI) { const uint32_t r = q + __popc(mk & ((1u << lane) - 1u)); if (r < 64u) IL[r] = (uint16_t)i; }
                    q += __popc(mk);
                }
                __syncwarp();
                uint32_t na = 0, nB = 0;
                const bool ok = q <= 63u;
                for (uint32_t r = 0; ok && r < q; ++r) {
                    const uint32_t idx = IL[r], o = (Sx[idx] >> 16) & 0x3FFFu;
                    const uint32_t k = idx - nB;
                    const uint32_t a0 = lane < na ? (uint32_t)AH[lane] : 0xFFFFFu, a1 = lane + 32 < na ? (uint32_t)AH[lane + 32] : 0xFFFFFu;
                    const uint32_t cnt = __popc(__ballot_sync(0xffffffffu, lane < na && a0 - lane <= k)) +
                                         __popc(__ballot_sync(0xffffffffu, lane + 32 < na && a1 - (lane + 32) <= k));
                    if (o > k + cnt) {                                    // lands ahead of the lowest free slot
                        const uint32_t p0 = __popc(__ballot_sync(0xffffffffu, lane < na && a0 < o)) + __popc(__ballot_sync(0xffffffffu, lane + 32 < na && a1 < o));
                        const unsigned long long run = (unsigned long long)__ballot_sync(0xffffffffu, lane >= p0 && lane < na && a0 == o + (lane - p0)) |
                                                       ((unsigned long long)__ballot_sync(0xffffffffu, lane + 32 >= p0 && lane + 32 < na && a1 == o + (lane + 32 - p0)) << 32);
                        const uint32_t rl = (uint32_t)(__ffsll((long long)~(run >> p0)) - 1);
                        const uint32_t sl = o + rl, pos = p0 + rl;
                        __syncwarp();
                        if (lane >= pos && lane < na) AH[lane + 1] = (uint16_t)a0;
                        if (lane + 32 >= pos && lane + 32 < na) AH[lane + 33] = (uint16_t)a1;
                        if (lane == 0) { AH[pos] = (uint16_t)sl; BI[nB] = (uint16_t)idx; BS[nB] = (uint16_t)sl; }
                        __syncwarp();
                        ++na; ++nB;
                    }
                }
                if (lane == 0) *p1info = na | (nB << 8) | (ok ? 0x80000000u : 0u);
            }
            T.sync();
            const uint32_t info = *p1info;
            if (info >> 31) {
                p1_closed = true;
                const uint32_t na = info & 0xFFu, nB = (info >> 8) & 0xFFu;
                for (uint32_t j = tt; j < m; j += TEAM) E1x[j] = (uint16_t)NONE16;
                T.sync();
                for (uint32_t i = tt; i < n1; i += TEAM) {
                    uint32_t lo2 = 0, hi2 = nB;                           // B entries with a smaller index
                    while (lo2 < hi2) { const uint32_t mid = (lo2 + hi2) >> 1; if ((uint32_t)BI[mid] < i) lo2 = mid + 1; else hi2 = mid; }
                    if (lo2 < nB && (uint32_t)BI[lo2] == i) E1x[BS[lo2]] = (uint16_t)i;
                    else {
                        const uint32_t k = i - lo2;
                        uint32_t a = 0, b2 = na;                          // ahead slots with (slot - rank) <= k: they lie below the k-th other slot
                        while (a < b2) { const uint32_t mid = (a + b2) >> 1; if ((uint32_t)AH[mid] - mid <= k) a = mid + 1; else b2 = mid; }
                        E1x[k + a] = (uint16_t)i;
                    }
                }
                if (tt == 0) { T.scr[10] = 0u; T.scr[11] = 0u; }     // (p1info has been read by everyone: the helpers' progress words)
                T.sync();
                if (tt == 0) { __threadfence_block(); *prog = m; }
            }
        }
        if (TEAM != 32 && m - n1 > 0u && (p1_closed ? (tt >> 5) != 1u : (tt >> 5) == 2u)) {
            // release ranks: for every slot phase 1 has finished, how many phase-2 entries arrive before the slot is dead (lower
            // bound of its release time in the time-ordered phase-2 entries); parked in E2x[j] until phase 2 reaches the slot.
            // Chunks of 32 slots; behind the closed form warps 0, 2 and 3 take them in turn (each publishes the start of its next
            // chunk), behind a sweeping warp 0 it is warp 2 alone.
            const uint32_t cnt2 = m - n1;
            const uint32_t nh = p1_closed ? 3u : 1u, hidx = p1_closed ? ((tt >> 5) == 0u ? 0u : (tt >> 5) - 1u) : 0u;
            volatile uint32_t* mine = reinterpret_cast<volatile uint32_t*>(T.scr + 9 + hidx);
            for (uint32_t j0 = hidx * 32u; j0 < m; j0 += nh * 32u) {
                const uint32_t need = j0 + 32u < m ? j0 + 32u : m;
                while (*prog < need) { __nanosleep(40); }
                __threadfence_block();
                const uint32_t j = j0 + lane;
                if (j < m) {
                    const uint32_t y1 = E1x[j];
                    uint32_t a = 0;
                    if (y1 != NONE16) {
                        const uint32_t rel = (Sx[y1] & 0xFFFFu) + W + 1u;
                        uint32_t b = cnt2;
                        while (a < b) { const uint32_t mid = (a + b) >> 1; if ((Sx[n1 + mid] & 0xFFFFu) >= rel) b = mid; else a = mid + 1; }
                    }
                    E2x[j] = (uint16_t)a;
                }
                __syncwarp();
                if (lane == 0) { __threadfence_block(); *mine = j0 + nh * 32u; }
            }
            if (lane == 0) { __threadfence_block(); *mine = 0xFFFFFFFFu; }
            BC_MARK(5, 64u);
        }
        auto next_home_w = [&](uint32_t x) -> uint32_t {
            for (uint32_t base = x + 1; base < m; base += 32) {
                const uint32_t j = base + lane;
                const uint32_t r = __reduce_min_sync(0xffffffffu, (j < m && (Sx[j] & ISHOME)) ? j : m);
                if (r < m) return r;
            }
            return m;
        };
        for (uint32_t ph = 0; ph < 2; ++ph) {
            const uint32_t ebase = ph ? n1 : 0u, cntp = ph ? m - n1 : n1;
            uint16_t* Ex = ph ? E2x : E1x;
            if (ph && TEAM == 32) BC_STAMP(1);
            const bool sweeper = (TEAM == 32) ? true : ((tt >> 5) == ph);
            if (!sweeper || (!ph && p1_closed)) continue;
            if (cntp == 0) {
                for (uint32_t j = lane; j < m; j += 32) Ex[j] = (uint16_t)NONE16;
                __syncwarp();
                if (!ph && TEAM != 32 && lane == 0) { __threadfence_block(); *prog = m; }
                continue;
            }
            uint32_t elig = 0, pend = 0, pmin = 0xFFFFFFFFu;              // pmin: the earliest home among my pending entries
            for (uint32_t r = 0; r * 32u < cntp; ++r) {                   // a row of 32 entries per step, lane r keeps the row's masks
                const uint32_t g = r * 32u + lane;
                const uint32_t hm = g < cntp ? (Sx[ebase + g] >> 16) & 0x3FFFu : 0xFFFFFFFFu;
                const uint32_t be = __ballot_sync(0xffffffffu, hm == 0u), bp = __ballot_sync(0xffffffffu, hm != 0u && hm != 0xFFFFFFFFu);
                const uint32_t rmin = __reduce_min_sync(0xffffffffu, hm == 0u ? 0xFFFFFFFFu : hm);
                if (lane == r) { elig = be; pend = bp; pmin = rmin; }
            }
            if (ph && TEAM != 32) BC_MARK(6, 32u);
            uint32_t x = 0;
            while (x < m) {
                const uint32_t xe = next_home_w(x);
                if (x > 0 && __any_sync(0xffffffffu, pmin <= x)) {        // entries whose home the sweep has reached join the pool
                    if (pmin <= x) {
                        uint32_t pb = pend, nmin = 0xFFFFFFFFu;
                        while (pb) {
                            const uint32_t bq = (uint32_t)(__ffs(pb) - 1);
                            pb &= pb - 1;
                            const uint32_t hm = (Sx[ebase + 32u * lane + bq] >> 16) & 0x3FFFu;
                            if (hm <= x) { elig |= 1u << bq; pend &= ~(1u << bq); } else nmin = min(nmin, hm);
                        }
                        pmin = nmin;
                    }
                }
                if (ph && TEAM != 32) {                                   // phase 1 and the release ranks are past this segment
                    volatile uint32_t* pw = reinterpret_cast<volatile uint32_t*>(T.scr + 9);
                    for (;;) {
                        uint32_t f = pw[0];
                        if (p1_closed) f = min(f, min(pw[1], pw[2]));
                        if (f >= xe) break;
                        __nanosleep(20);
                    }
                    __threadfence_block();
                }
                uint32_t used = xe;
                if (ph) {                                                 // the phase-1 slots of a segment are a prefix of it
                    if (xe - x <= 96u) {
                        used = x;
                        for (uint32_t b0 = x; b0 < xe; b0 += 32) {
                            const uint32_t fm = __ballot_sync(0xffffffffu, b0 + lane < xe && E1x[b0 + lane] != NONE16);
                            used = b0 + (uint32_t)__popc(fm);
                            if (fm != 0xFFFFFFFFu) break;
                        }
                    } else { uint32_t a = x, b = xe; while (a < b) { const uint32_t mid = (a + b) >> 1; if (E1x[mid] != NONE16) a = mid + 1; else b = mid; } used = a; }
                }
                for (uint32_t part = 0; part < (ph ? 2u : 1u); ++part) {
                    const uint32_t lo = ph ? (part ? used : x) : x, hi = ph ? (part ? xe : used) : xe;
                    if (lo >= hi) continue;
                    const bool use_rel = ph && part == 0;
                    const uint32_t elig0 = elig, pc = __popc(elig0);
                    const uint32_t incl = warp_incl_scan_u32(pc), pre = incl - pc, np = __shfl_sync(0xffffffffu, incl, 31);
                    if (np == 0) {
                        for (uint32_t j = lo + lane; j < hi; j += 32) Ex[j] = (uint16_t)NONE16;
                        __syncwarp();
                        continue;
                    }
                    int carry = 0;
                    for (uint32_t i0 = 0; i0 < hi - lo; i0 += 32) {
                        const uint32_t i = i0 + lane, j = lo + i;
                        const bool valid = j < hi;
                        uint32_t G = 0;
                        if (valid && use_rel) {
                            if (TEAM != 32) G = E2x[j];                                  // (warp 2)
                            else {
                                const uint32_t rel = (Sx[E1x[j]] & 0xFFFFu) + W + 1u;
                                uint32_t a = 0, b = cntp;
                                while (a < b) { const uint32_t mid = (a + b) >> 1; if ((Sx[ebase + mid] & 0xFFFFu) >= rel) b = mid; else a = mid + 1; }
                                G = a;
                            }
                        }
                        int u = carry;
                        if (use_rel) {                       // (warp-uniform) pool rank of the release bound, running maximum
                            const uint32_t wd = G >> 5;
                            const uint32_t pw = __shfl_sync(0xffffffffu, pre, wd & 31u), ew = __shfl_sync(0xffffffffu, elig0, wd & 31u);
                            const uint32_t lbp = wd >= 32u ? np : pw + __popc(ew & ((1u << (G & 31u)) - 1u));
                            int v = valid ? (int)lbp - (int)i : -(1 << 28);
#pragma unroll
                            for (int d = 1; d < 32; d <<= 1) { const int t2 = __shfl_up_sync(0xffffffffu, v, d); if ((int)lane >= d) v = max(v, t2); }
                            const int last = __shfl_sync(0xffffffffu, v, 31);
                            u = max(v, carry);
                            carry = max(carry, last);
                        }
                        const uint32_t idx = i + (uint32_t)u;
                        const bool take = valid && idx < np;
                        uint32_t t = 0;
#pragma unroll
                        for (uint32_t st = 16; st > 0; st >>= 1) {
                            const uint32_t c = t + st;
                            const uint32_t pcand = __shfl_sync(0xffffffffu, pre, c & 31u);
                            if (c < 32u && pcand <= idx) t = c;
                        }
                        const uint32_t pt = __shfl_sync(0xffffffffu, pre, t), et = __shfl_sync(0xffffffffu, elig0, t);
                        const uint32_t g = take ? 32u * t + nth_set(et, idx - pt) : 0u;
                        const uint32_t ent = take ? ebase + g : NONE16;
                        if (valid) Ex[j] = (uint16_t)ent;
                        if (take) atomicOr(&scr[g >> 5], 1u << (g & 31u));    // clear the entries handed out (through the warp's scratch words)
                        __syncwarp();
                        elig &= ~scr[lane];
                        __syncwarp();
                        scr[lane] = 0;
                        __syncwarp();
                        if ((int)(i0 + 32u) + carry >= (int)np) {
                            for (uint32_t j2 = lo + i0 + 32u + lane; j2 < hi; j2 += 32) Ex[j2] = (uint16_t)NONE16;
                            break;
                        }
                    }
                    __syncwarp();
                }
                if (!ph && TEAM != 32) { __syncwarp(); if (lane == 0) { __threadfence_block(); *prog = xe; } }
                x = xe;
            }
            if (ph && TEAM != 32) BC_MARK(7, 32u);
        }
        BC_STAMP(3);
        // post-pass (whole team): the finds, slot by slot (every occupant of [home, own slot) is final now; an e1 that had
        // expired at p's time was followed by an e2 placed before p, or the slot would have been dead and p would sit there)
        T.sync();
        BC_STAMP(2);
        for (uint32_t base = 0; base < m; base += TEAM) {
            const uint32_t j = base + tt;
#pragma unroll
            for (uint32_t ph = 0; ph < 2; ++ph) {
                const uint32_t i = j < m ? (uint32_t)(ph ? E2x[j] : E1x[j]) : NONE16;
                const bool act = i != NONE16;
                uint32_t p = 0, o = 0;
                if (act) { const uint32_t e = Sx[i]; p = e & 0xFFFFu; o = (e >> 16) & 0x3FFFu; }
                const uint32_t f = find(act, o, j, p, ph != 0u);
                if (act) fres[p] = (uint16_t)f;
            }
        }
        BC_STAMP(4);
        return;
    }
    // ---- phase 1: between two homes the next slots go to the next unplaced entries whose home is at/before the segment
    {
        uint32_t x = 0, fu = 0;
        while (x < m) {
            const uint32_t xe = next_home(x);
            const uint32_t seglen = xe - x;
            uint32_t filled = 0;
            for (uint32_t base = fu; base < n1 && filled < seglen; base += TEAM) {
                const uint32_t i = base + tt;
                const uint32_t e = i < n1 ? Sx[i] : 0xFFFFFFFFu;
                const bool elig = i < n1 && !(e & PLACED) && ((e >> 16) & 0x3FFFu) <= x;
                uint32_t tot;
                const uint32_t r = T.exscan(elig ? 1u : 0u, tot);
                if (elig && filled + r < seglen) { E1x[x + filled + r] = (uint16_t)i; Sx[i] = e | PLACED; }
                filled += tot;
            }
            if (filled > seglen) filled = seglen;
            for (uint32_t j = x + filled + tt; j < xe; j += TEAM) E1x[j] = (uint16_t)NONE16;
            T.sync();
            for (;;) {                                       // first entry that is still unplaced
                const uint32_t i = fu + tt;
                const uint32_t r = T.tmin((i < n1 && !(Sx[i] & PLACED)) ? i : n1);
                if (r < n1 || fu + TEAM >= n1) { fu = r; break; }
                fu += TEAM;
            }
            x = xe;
        }
    }
    for (uint32_t base = 0; base < m; base += TEAM) {
        const uint32_t j = base + tt;
        const uint32_t i = j < m ? (uint32_t)E1x[j] : NONE16;
        const bool act = i != NONE16;
        uint32_t p = 0, o = 0;
        if (act) { const uint32_t e = Sx[i]; p = e & 0xFFFFu; o = (e >> 16) & 0x3FFFu; }
        const uint32_t f = find(act, o, j, p, false);
        if (act) fres[p] = (uint16_t)f;
    }
    if (n1 == m) return;
    // ---- phase 2
    uint32_t x = 0, fu2 = n1;
    while (x < m) {
        const uint32_t xe = next_home(x);
        uint32_t used;                                       // the phase-1 slots of a segment are a prefix of it
        { uint32_t a = x, b = xe; while (a < b) { const uint32_t mid = (a + b) >> 1; if (E1x[mid] != NONE16) a = mid + 1; else b = mid; } used = a; }
        for (uint32_t part = 0; part < 2; ++part) {
            const uint32_t lo = part ? used : x, hi = part ? xe : used;
            if (lo >= hi) continue;
            uint32_t np = 0;                                 // pool: unplaced phase-2 entries with home at/before the segment, in time order
            for (uint32_t base = fu2; base < m; base += TEAM) {
                const uint32_t i = base + tt;
                const uint32_t e = i < m ? Sx[i] : 0xFFFFFFFFu;
                const bool elig = i < m && !(e & PLACED) && ((e >> 16) & 0x3FFFu) <= x;
                uint32_t tot;
                const uint32_t r = T.exscan(elig ? 1u : 0u, tot);
                if (elig) E2x[lo + np + r] = (uint16_t)i;
                np += tot;
            }
            T.sync();
            if (np == 0) {
                for (uint32_t j = lo + tt; j < hi; j += TEAM) E2x[j] = (uint16_t)NONE16;
                T.sync();
                continue;
            }
            int carry = 0;
            for (uint32_t i0 = 0; i0 < hi - lo; i0 += TEAM) {
                const uint32_t i = i0 + tt, j = lo + i;
                const bool valid = j < hi;
                int v = -(1 << 28);
                if (valid) {
                    uint32_t a = i0 < np ? i0 : np, b = np;
                    if (part == 0) {
                        const uint32_t rel = (Sx[E1x[j]] & 0xFFFFu) + W + 1u;     // first arrival time for which the slot is dead
                        while (a < b) { const uint32_t mid = (a + b) >> 1; if ((Sx[E2x[lo + mid]] & 0xFFFFu) >= rel) b = mid; else a = mid + 1; }
                    }
                    v = (int)a - (int)i;
                }
                int last;
                int u = T.maxscan(v, last);
                u = max(u, carry);
                carry = max(carry, last);
                const uint32_t idx = i + (uint32_t)u;
                const uint32_t ent = (valid && idx < np) ? E2x[lo + idx] : NONE16;
                T.sync();                                    // every pool read of the group is done
                if (valid) { E2x[j] = (uint16_t)ent; if (ent != NONE16) Sx[ent] |= PLACED; }
                T.sync();
                {                                            // find of the entry that just got slot j
                    const bool act = valid && ent != NONE16;
                    uint32_t p = 0, o = 0;
                    if (act) { const uint32_t e = Sx[ent]; p = e & 0xFFFFu; o = (e >> 16) & 0x3FFFu; }
                    const uint32_t f = find(act, o, j, p, true);
                    if (act) fres[p] = (uint16_t)f;
                }
                if ((int)(i0 + TEAM) + carry >= (int)np) {   // the pool is used up: the remaining slots stay dead
                    for (uint32_t j2 = lo + i0 + TEAM + tt; j2 < hi; j2 += TEAM) E2x[j2] = (uint16_t)NONE16;
                    break;
                }
            }
            T.sync();
            for (;;) {
                const uint32_t i = fu2 + tt;
                const uint32_t r = T.tmin((i < m && !(Sx[i] & PLACED)) ? i : m);
                if (r < m || fu2 + TEAM >= m) { fu2 = r; break; }
                fu2 += TEAM;
            }
        }
        x = xe;
    }
}

__device__ __forceinline__ long long clk_ordered() { long long t; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)::"memory"); return t; }
#define CLK() (DBG ? clk_ordered() : 0ll)
template <bool DBG>
__global__ void __launch_bounds__(NTHREADS, 1) lz77_v4_kernel(const uint8_t* __restrict__ in, uint64_t n, uint32_t bs, uint32_t nblocks,
                                                             uint16_t* __restrict__ fres_all, uint32_t* __restrict__ tok_all,
                                                             uint8_t* __restrict__ scratch, uint64_t stride,
                                                             uint64_t* __restrict__ block_sizes, uint64_t* __restrict__ block_bytes,
                                                             uint32_t* __restrict__ fb_list, uint32_t* __restrict__ fb_cnt,
                                                             const uint32_t* __restrict__ use_flag,
                                                             uint32_t* __restrict__ dbg_tok) {
    if (use_flag && *use_flag == 0u) {       // the sample says this input is not text-like: every block goes to lz77_v2_kernel
        if (threadIdx.x == 0) for (uint32_t b = blockIdx.x; b < nblocks; b += gridDim.x) fb_list[atomicAdd(fb_cnt, 1u)] = b;
        return;
    }
    extern __shared__ __align__(16) uint8_t smem[];
    uint8_t* data = smem + OFF_DATA;
    uint8_t* big = smem + OFF_BIG;
    uint32_t* bm = reinterpret_cast<uint32_t*>(big);
    uint32_t* pre = reinterpret_cast<uint32_t*>(smem + OFF_PRE);
    uint16_t* pre16 = reinterpret_cast<uint16_t*>(smem + OFF_PRE);
    uint32_t* flags = reinterpret_cast<uint32_t*>(smem + OFF_FLAGS);
    Misc* ms = reinterpret_cast<Misc*>(smem + OFF_MISC);
    uint32_t* S = reinterpret_cast<uint32_t*>(big);
    uint16_t* E1 = reinterpret_cast<uint16_t*>(big + CH * 4);
    uint16_t* E2 = E1 + CH;
    uint16_t* wl = reinterpret_cast<uint16_t*>(smem + OFF_BIG + WL_S16);   // work list: compact start of every cluster of 2 .. LMAX entries, largest classes first (tail of BIG + head of PRE)
    uint32_t* bigl = reinterpret_cast<uint32_t*>(smem + OFF_PRE + WL_PRE * 2);   // clusters above LMAX entries: compact start | size << 16
    uint16_t* clist = reinterpret_cast<uint16_t*>(smem + OFF_PRE + 8192);         // final stage only: [4][NBIG] indices into bigl by size class
    uint32_t* wscr = reinterpret_cast<uint32_t*>(smem + OFF_PRE + 4096);          // final stage only (behind its rank prefix): 32 words per warp, zero between uses
    uint8_t* adv = big;
    uint8_t* exitof = big + PADDED;

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint16_t* fres = fres_all + (uint64_t)blockIdx.x * MAXB;
    uint16_t* wlg = fres_all + ((uint64_t)gridDim.x + blockIdx.x) * MAXB;   // work-list spill (second half of the launch's F(p) allocation)
    uint32_t* tokb = tok_all + (uint64_t)blockIdx.x * MAXB;
    uint32_t* dbg_stats = dbg_tok ? dbg_tok + (uint64_t)nblocks * MAXB : nullptr;

    for (uint32_t b = blockIdx.x; b < nblocks; b += gridDim.x) {
        const long long t_begin = CLK();
        if (DBG && dbg_stats && tid < 128) dbg_stats[(uint64_t)b * 136 + 8 + tid] = 0;
#define PHASE_STAMP(k) do { if (DBG && dbg_stats && tid == 0) dbg_stats[(uint64_t)b * 136 + (k)] = (uint32_t)(clk_ordered() - t_begin); } while (0)
        const uint8_t* src = in + (uint64_t)b * bs;
        const uint32_t len = (uint32_t)(n - (uint64_t)b * bs < bs ? n - (uint64_t)b * bs : bs);

        // ---------------- P0: block -> shared memory (zero pad behind it, U1), zero bitmap / summary / flags
        {
            const bool al = (reinterpret_cast<uintptr_t>(src) & 15) == 0;
            for (uint32_t i = tid * 16; i < len + 128 && i < SZ_DATA; i += NTHREADS * 16) {
                if (al && i + 16 <= len) *reinterpret_cast<uint4*>(data + i) = __ldg(reinterpret_cast<const uint4*>(src + i));
                else for (uint32_t k = 0; k < 16 && i + k < SZ_DATA; ++k) data[i + k] = (i + k < len) ? __ldg(src + i + k) : 0;
            }
            for (uint32_t i = tid; i < BM_WORDS; i += NTHREADS) bm[i] = 0;
            for (uint32_t i = tid; i < BM_WORDS / 32; i += NTHREADS) pre[i] = 0;
            for (uint32_t i = tid; i < SZ_FLAGS / 4; i += NTHREADS) flags[i] = 0;
            if (tid == 0) { ms->p1_next = 0; ms->fallback = 0; ms->nbig = 0; }
        }
        __syncthreads();
        PHASE_STAMP(0);
        // ---------------- P1: no-expiry occupancy by atomic linear probing (as v2); the slot an entry wins is ITS slot
        {
            uint32_t* summ = pre;
            for (;;) {
                uint32_t row = 0;
                if (lane == 0) row = atomicAdd(&ms->p1_next, 32u);
                row = __shfl_sync(0xffffffffu, row, 0);
                if (row >= len) break;
                const uint32_t i = row + lane;
                if (i >= len) continue;
                const uint32_t h = lz_hash(sm_word(data, i));
                uint32_t s = h, cl;
                for (;;) {
                    const uint32_t wi = s >> 5;
                    const uint32_t free_bits = ~bm[wi] & (0xFFFFFFFFu << (s & 31));
                    if (free_bits) {
                        const uint32_t bp = (uint32_t)(__ffs(free_bits) - 1), bit = 1u << bp;
                        const uint32_t old = atomicOr(&bm[wi], bit);
                        if (!(old & bit)) {
                            if ((old | bit) == 0xFFFFFFFFu) atomicOr(&summ[wi >> 5], 1u << (wi & 31));
                            cl = (wi << 5) + bp;
                            break;
                        }
                        continue;
                    }
                    uint32_t nw = wi + 1;
                    uint32_t open = ~summ[nw >> 5] & (0xFFFFFFFFu << (nw & 31));
                    while (!open) { nw = ((nw >> 5) + 1) << 5; open = ~summ[nw >> 5]; }
                    s = ((nw & ~31u) + (uint32_t)(__ffs(open) - 1)) << 5;
                    if (s >= SLOTS) s = 0;                            // the insert wraps (deflate/lz77.c:99-101)
                }
                const uint32_t d = cl - h;
                tokb[i] = cl | ((d < 4095u ? d : 4095u) << 20);
            }
        }
        __syncthreads();
        PHASE_STAMP(1);
        // ---------------- P2: rank prefix per 4 words, cluster-start flag per compact slot
        {
            const uint32_t c0 = tid * 5;
            uint32_t part[5], half[5], mine = 0;
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                const uint32_t ch = c0 + k;
                uint32_t s = 0, s4 = 0;
                if (ch < PRE_N) {
                    const uint4 qa = *reinterpret_cast<const uint4*>(bm + ch * PRE_CHUNK), qb = *reinterpret_cast<const uint4*>(bm + ch * PRE_CHUNK + 4);
                    s4 = __popc(qa.x) + __popc(qa.y) + __popc(qa.z) + __popc(qa.w);
                    s = s4 + __popc(qb.x) + __popc(qb.y) + __popc(qb.z) + __popc(qb.w);
                }
                part[k] = s; half[k] = s4; mine += s;
            }
            const uint32_t incl = warp_incl_scan_u32(mine);
            if (lane == 31) ms->scan[warp] = incl;
            __syncthreads();
            if (warp == 0) {
                const uint32_t t = ms->scan[lane];
                const uint32_t ti = warp_incl_scan_u32(t);
                ms->scan[lane] = ti - t;
            }
            if (tid == 32) {
                // The cluster on slot 0 (the reference clears that slot early, U10) and the one that ends on the last slot (the
                // insert wraps from there to slot 0, find does not) are simulated by one thread with the reference's own rules
                // (special_clusters below); here: their raw bounds. More than 60 slots between them: the block is handed back.
                uint32_t c0s = 0;
                while (c0s < 64u && ((bm[c0s >> 5] >> (c0s & 31u)) & 1u)) ++c0s;
                uint32_t ts = SLOTS;
                while (ts > SLOTS - 64u && ((bm[(ts - 1) >> 5] >> ((ts - 1) & 31u)) & 1u)) --ts;
                ms->cut0 = c0s; ms->top_start = ts; ms->nsp = 0;
                if (c0s + (SLOTS - ts) > 60u) ms->fallback = 1u;
            }
            __syncthreads();
            uint32_t run = ms->scan[warp] + incl - mine;
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                const uint32_t ch = c0 + k;
                if (ch < PRE_N) { pre16[2 * ch] = (uint16_t)run; pre16[2 * ch + 1] = (uint16_t)(run + half[k]); }
                run += part[k];
            }
            __syncthreads();
            // cluster-start flags, four bitmap words per thread and step, consecutive threads on consecutive groups (a walk along a
            // thread's own 40 words has eight lanes of a warp on every bank): the rank before a group comes from the prefix, the bit
            // before it from the neighbouring lane
            for (uint32_t g0 = 0; g0 < BM_WORDS / 4; g0 += NTHREADS) {
                const uint32_t g = g0 + tid;
                const bool in = g < BM_WORDS / 4;
                const uint4 q = in ? *reinterpret_cast<const uint4*>(bm + 4 * g) : make_uint4(0u, 0u, 0u, 0u);
                uint32_t prev = __shfl_up_sync(0xffffffffu, q.w, 1);
                if (lane == 0) prev = (in && g) ? bm[4 * g - 1] : 0u;
                prev >>= 31;
                if (q.x | q.y | q.z | q.w) {
                    const uint32_t wd[4] = {q.x, q.y, q.z, q.w};
                    uint32_t rw = pre16[g], fw = 0xFFFFFFFFu, fbits = 0;
#pragma unroll
                    for (uint32_t wq_ = 0; wq_ < 4; ++wq_) {
                        const uint32_t xw = wd[wq_];
                        uint32_t st = xw & ~((xw << 1) | prev);
                        while (st) {
                            const uint32_t bp = (uint32_t)(__ffs(st) - 1);
                            st &= st - 1;
                            const uint32_t u = rw + __popc(xw & ((1u << bp) - 1u));
                            if ((u >> 5) != fw) { if (fbits) atomicOr(&flags[fw], fbits); fw = u >> 5; fbits = 0; }
                            fbits |= 1u << (u & 31);
                        }
                        rw += __popc(xw);
                        prev = xw >> 31;
                    }
                    if (fbits) atomicOr(&flags[fw], fbits);
                }
            }
            if (tid == 0) atomicOr(&flags[len >> 5], 1u << (len & 31));   // sentinel behind the last compact slot
        }
        __syncthreads();
        const bool fb0 = ms->fallback != 0;
        PHASE_STAMP(2);
        // ---------------- P3: compact slot of the claimed slot, displacement from the home; loners
        const uint32_t sp_cut0 = ms->cut0, sp_top = ms->top_start;
        uint16_t* spl = reinterpret_cast<uint16_t*>(&ms->sexit[0][0]);          // (the parse scratch is free until P5) [64] positions
        if (!fb0) {
            for (uint32_t i0 = tid; i0 < len; i0 += 8 * NTHREADS) {
                uint32_t tv[8];
#pragma unroll
                for (uint32_t k = 0; k < 8; ++k) { const uint32_t i = i0 + k * NTHREADS; tv[k] = i < len ? tokb[i] : 0u; }
                // (branch-free per entry, so that the eight entries' shared-memory reads overlap)
                uint32_t hv[8], rk[8];
                bool ln[8];
#pragma unroll
                for (uint32_t k = 0; k < 8; ++k) {
                    const uint32_t i = i0 + k * NTHREADS;
                    const uint32_t cl = tv[k] & 0xFFFFFu, dd = tv[k] >> 20;
                    uint32_t h = cl - dd;
                    if (dd >= 4095u) h = lz_hash(sm_word(data, i < len ? i : 0u));
                    hv[k] = h;
                    const uint32_t hb = h ? h - 1u : 0u, ha = h + 1u;
                    const uint32_t below = h ? (bm[hb >> 5] >> (hb & 31u)) & 1u : 0u;
                    const uint32_t above = (bm[ha >> 5] >> (ha & 31u)) & 1u;
                    ln[k] = cl == h && !below && !above;
                    rk[k] = bm_rank(bm, pre16, cl) | ((cl - h) << 16);
                }
#pragma unroll
                for (uint32_t k = 0; k < 8; ++k) {
                    const uint32_t i = i0 + k * NTHREADS;
                    if (i >= len) break;
                    const uint32_t h = hv[k];
                    if (h < sp_cut0 || h >= sp_top) {                     // entry of a special cluster: listed, kept out of everything else
                        const uint32_t si = atomicAdd(&ms->nsp, 1u);
                        if (si < 60u) spl[si] = (uint16_t)i; else ms->fallback = 1u;
                        tokb[i] = LONER; fres[i] = (uint16_t)NONE16;
                        continue;
                    }
                    tokb[i] = ln[k] ? LONER : rk[k];
                    fres[i] = (uint16_t)NONE16;
                }
            }
        }
        __syncthreads();
        if (!fb0 && tid == 0 && !ms->fallback && ms->nsp) {
            // ---- the special clusters, serially, exactly as the reference runs them (deflate/lz77.c:77-174 with lazy expiry):
            // local slots [0, mt) = the run that ends on the table's last slot, [mt, mt + ml) = the run that starts on slot 0, so
            // the wrapping insert just walks on; find stops at the table end; slot 0 (local mt) is wiped W inserts after every
            // placement into it and once at insert W - 1 (the ring's zero-initialised entries, U10).
            uint16_t* Tsp = spl + 64;                                            // [64] position + 1 of the occupant, 0 = none
            uint32_t* clr = reinterpret_cast<uint32_t*>(spl + 128);              // [<= 64] pending wipe times of slot 0
            const uint32_t ns = ms->nsp, mt = SLOTS - sp_top, ml = sp_cut0, tot = mt + ml;
            for (uint32_t a = 1; a < ns; ++a) { const uint16_t v = spl[a]; uint32_t j = a; while (j > 0 && spl[j - 1] > v) { spl[j] = spl[j - 1]; --j; } spl[j] = v; }
            for (uint32_t k = 0; k < 64; ++k) Tsp[k] = 0;
            uint32_t qh = 0, qt = 0;
            clr[qt++] = W - 1;
            for (uint32_t a = 0; a < ns; ++a) {
                const uint32_t q = spl[a], w = sm_word(data, q), h = lz_hash(w);
                const uint32_t loc = h >= sp_top ? h - sp_top : mt + h;
                while (qh < qt && clr[qh] < q) { if (ml) Tsp[mt] = 0; ++qh; }
                const uint32_t dthr = q > W ? q - W : 0u;
                uint32_t k = loc, m = 0xFFFFFFFFu;
                bool ran_off = false;
                for (;;) {
                    const uint32_t v = Tsp[k];
                    if (v <= dthr) break;
                    if (sm_word(data, v - 1) == w) { m = v - 1; break; }
                    if (loc < mt && k + 1 == mt) { ran_off = true; break; }   // find does not wrap (deflate/lz77.c:168)
                    if (k + 1 >= tot) break;
                    ++k;
                }
                uint32_t e = ran_off ? mt : k;                                // the wrapping insert continues at slot 0
                while (e + 1 < tot && Tsp[e] > dthr) ++e;
                if (q != 65535u) Tsp[e] = (uint16_t)(q + 1);
                if (e == mt && ml && qt < 62u) clr[qt++] = q + W;
                if (qh < qt && clr[qh] == q) { if (ml) Tsp[mt] = 0; ++qh; }
                fres[q] = (uint16_t)(m == 0xFFFFFFFFu ? NONE16 : m);
            }
        }
        __syncthreads();
        PHASE_STAMP(3);
        // ---------------- clusters, one chunk of the compact slot space at a time
        long long dt_sc = 0, dt_c2 = 0, dt_w = 0, dt_t = 0, dt_wait = 0, t_mark = CLK();
        uint32_t n_chunks = 0, n_wq = 0;
#define SUBSTAMP(acc) do { if (DBG) { const long long t_now = clk_ordered(); acc += t_now - t_mark; t_mark = t_now; } } while (0)
        if (!fb0 && !ms->fallback) {
            // ---- lane stage: every entry's bare position at S16[compact slot] (131072 bytes: the whole block at once; the home
            // offsets come back from the hashes, lane_cluster), clusters of 2 .. LMAX entries one lane each
            uint16_t* S16 = reinterpret_cast<uint16_t*>(big);
            if (tid == 0) { for (int c = 0; c < 7; ++c) { ms->ccnt[c] = 0; ms->cfill[c] = 0; } }
            for (uint32_t i0 = tid; i0 < len; i0 += 8 * NTHREADS) {
                uint32_t tv[8];
#pragma unroll
                for (uint32_t k = 0; k < 8; ++k) { const uint32_t i = i0 + k * NTHREADS; tv[k] = i < len ? tokb[i] : LONER; }
#pragma unroll
                for (uint32_t k = 0; k < 8; ++k)
                    if (tv[k] != LONER) S16[tv[k] & 0xFFFFu] = (uint16_t)(i0 + k * NTHREADS);
            }
            __syncthreads();
            SUBSTAMP(dt_sc);
            // work list: every thread looks at the 64 compact slots of two flag words, counts its clusters per size class, then
            // files them (the size of a listed cluster is read from the flags again when it is worked)
            auto cls_of = [](uint32_t m) -> uint32_t { return m > 8 ? 3u : m > 4 ? 4u : m > 2 ? 5u : 6u; };
            auto csize = [&](uint32_t start) -> uint32_t {
                const uint32_t q = start + 1;
                uint32_t wi = q >> 5;
                uint32_t xw = flags[wi] & (0xFFFFFFFFu << (q & 31u));
                while (!xw) { ++wi; xw = flags[wi]; }
                return (wi << 5) + (uint32_t)(__ffs(xw) - 1) - start;
            };
            auto wl_put = [&](uint32_t i, uint32_t v) { if (i < WL_CAP) wl[i] = (uint16_t)v; else wlg[i - WL_CAP] = (uint16_t)v; };
            auto wl_get = [&](uint32_t i) -> uint32_t { return i < WL_CAP ? wl[i] : wlg[i - WL_CAP]; };
            const uint32_t sp_hi = len - (SLOTS - sp_top);
            uint32_t nlb[2] = {0, 0};                         // starts of clusters of 2 .. LMAX entries among my slots
            unsigned long long ccl[2] = {0, 0};               // ... and their size classes, two bits per slot (for the second pass)
            uint32_t cA = 0, cB = 0;                          // my clusters per size class: 9..16 | 5..8 << 16, 3..4 | 2 << 16
#pragma unroll
            for (uint32_t h = 0; h < 2; ++h) {
                const uint32_t a = (2 * tid + h) * 32;
                if (a < len) {
                    const uint32_t fb = flags[a >> 5], nx = flags[(a >> 5) + 1];
                    uint32_t bits = fb & ~((fb >> 1) | (nx << 31));              // a start directly followed by a start is a loner
                    if (len - a < 32u) bits &= (1u << (len - a)) - 1u;           // (the sentinel behind the last slot is no cluster)
                    nlb[h] = bits;
                    while (bits) {
                        const uint32_t k = (uint32_t)(__ffs(bits) - 1);
                        bits &= bits - 1;
                        const uint32_t start = a + k;
                        if (start < sp_cut0 || start >= sp_hi) { nlb[h] &= ~(1u << k); continue; }   // a special cluster: done above
                        const uint32_t m = csize(start);
                        if (m > LMAX) {                          // kept for the final stage, where all of them run at once
                            nlb[h] &= ~(1u << k);
                            const uint32_t bi = m <= CL_MAX ? atomicAdd(&ms->nbig, 1u) : NBIG;
                            if (bi < NBIG) bigl[bi] = start | (m << 16); else ms->fallback = 1u;
                            continue;
                        }
                        const uint32_t c = cls_of(m);
                        ccl[h] |= (unsigned long long)(c - 3u) << (2u * k);
                        if (c < 5u) cA += c == 3u ? 1u : 0x10000u; else cB += c == 5u ? 1u : 0x10000u;
                    }
                }
            }
            // places inside the list without an atomic per cluster (four counters shared by 9 000 clusters serialise): a warp scan of
            // the per-thread counts, one atomicAdd per class and warp
            const uint32_t iA = warp_incl_scan_u32(cA), iB = warp_incl_scan_u32(cB);
            uint32_t wA = 0, wB = 0;                          // the warp's first place inside each class
            if (lane == 31) {
                wA = atomicAdd(&ms->ccnt[3], iA & 0xFFFFu) | (atomicAdd(&ms->ccnt[4], iA >> 16) << 16);
                wB = atomicAdd(&ms->ccnt[5], iB & 0xFFFFu) | (atomicAdd(&ms->ccnt[6], iB >> 16) << 16);
            }
            wA = __shfl_sync(0xffffffffu, wA, 31); wB = __shfl_sync(0xffffffffu, wB, 31);
            uint32_t myA = wA + iA - cA, myB = wB + iB - cB;  // my first place inside each class (16-bit halves; a class holds < 32 768 clusters)
            __syncthreads();
            if (tid == 0) {
                uint32_t run = 0;
                for (int c = 0; c < 7; ++c) { ms->cbase[c] = run; run += ms->ccnt[c]; }
                ms->cbase[7] = run; ms->next_l = 0;
            }
            __syncthreads();
#pragma unroll
            for (uint32_t h = 0; h < 2; ++h) {
                const uint32_t a = (2 * tid + h) * 32;
                uint32_t nlbits = nlb[h];
                while (nlbits && !ms->fallback) {
                    const uint32_t k = (uint32_t)(__ffs(nlbits) - 1);
                    nlbits &= nlbits - 1;
                    const uint32_t c = 3u + ((uint32_t)(ccl[h] >> (2u * k)) & 3u);
                    uint32_t place;
                    if (c == 3u) { place = myA & 0xFFFFu; myA += 1u; } else if (c == 4u) { place = myA >> 16; myA += 0x10000u; }
                    else if (c == 5u) { place = myB & 0xFFFFu; myB += 1u; } else { place = myB >> 16; myB += 0x10000u; }
                    wl_put(ms->cbase[c] + place, a + k);
                }
            }
            __syncthreads();
            SUBSTAMP(dt_c2);
            if (!ms->fallback) {
                const uint32_t total = ms->cbase[7];
                if (DBG) { ++n_chunks; n_wq += total; }
                for (uint32_t rr = 0; rr * 1024u < total; ++rr) {             // 32 clusters of similar size per warp, one lane each; batches dealt
                    const uint32_t g = rr * 1024u + ((rr & 1u) ? 31u - warp : warp) * 32u;   // in snake order (the list runs from the costly to the cheap)
                    if (g >= total) continue;
                    const uint32_t it = g + lane;
                    if (it < total) {
                        const uint32_t kl = wl_get(it), m = csize(kl);
                        if (m == 2u) {
                            // two entries (half of all listed clusters): the later one finds the earlier one iff that is still live
                            // and has its 4-gram (equal 4-grams share the home; with different homes the later entry's own home
                            // is dead or never taken, and find scans nothing)
                            const uint32_t a = S16[kl], c = S16[kl + 1];
                            const uint32_t p0 = min(a, c), p1 = max(a, c);
                            fres[p0] = (uint16_t)NONE16;
                            fres[p1] = (uint16_t)((p0 + W >= p1 && sm_word(data, p0) == sm_word(data, p1)) ? p0 : NONE16);
                        } else {
                            const uint32_t hmin = lane_sort16(S16 + kl, m, data);
                            lane_cluster<true>(S16 + kl, nullptr, nullptr, m, hmin, data, fres);
                        }
                    }
                }
                if (DBG) SUBSTAMP(dt_w);
            }
            __syncthreads();
            SUBSTAMP(dt_wait);
        }
        // ---------------- final stage: the clusters above LMAX entries, all at once (the largest decides, not their sum).
        // Rounds of as many listed clusters as fit the chunk area; a bit per compact slot says "in a cluster of this round",
        // the rank in that bit map is the slot's place in S.
        long long dt_f[4] = {0, 0, 0, 0};
        while (!fb0 && !ms->fallback && ms->nbig) {
            const uint32_t nbig = ms->nbig;
            __syncthreads();
            if (tid == 0) { ms->rdone = 0; }
            uint32_t r0 = 0;
            while (r0 < nbig) {
                if (tid == 0) ms->scan[33] = 0;
                __syncthreads();
                {                                                // entries of all listed clusters that are left: one round if they fit
                    uint32_t mine = (r0 + tid < nbig) ? bigl[r0 + tid] >> 16 : 0u;
                    mine = warp_sum_u32(mine);
                    if (lane == 0 && mine) atomicAdd(&ms->scan[33], mine);
                }
                __syncthreads();
                if (tid == 0) {                                  // clusters r0 .. r1 of this round
                    uint32_t r1 = nbig;
                    if (ms->scan[33] > CH) { uint32_t tot = 0; r1 = r0; while (r1 < nbig && tot + (bigl[r1] >> 16) <= CH) { tot += bigl[r1] >> 16; ++r1; } }
                    ms->r0 = r0; ms->r1 = r1; ms->next_t = 0; ms->rdone = 0; ms->next_w = 0;
                    ms->ncls[0] = 0; ms->ncls[1] = 0; ms->ncls[2] = 0; ms->ncls[3] = 0;
                }
                for (uint32_t i = tid; i < SZ_FLAGS / 4; i += NTHREADS) flags[i] = 0;
                wscr[tid] = 0;
                __syncthreads();
                const uint32_t r1 = ms->r1;
                {   // the round's clusters by size class (one pick = one atomic, the larger ones of a tier first)
                    const uint32_t k = r0 + tid;
                    const uint32_t mk = k < r1 ? bigl[k] >> 16 : 0u;
                    const uint32_t c = mk > TMIN ? 0u : mk > 128u ? 1u : mk > L2MAX ? 2u : 3u;
#pragma unroll
                    for (uint32_t cc = 0; cc < 4; ++cc) {
                        const uint32_t bal = __ballot_sync(0xffffffffu, mk != 0u && c == cc);
                        if (bal) {
                            uint32_t base = 0;
                            if (lane == 0) base = atomicAdd(&ms->ncls[cc], (uint32_t)__popc(bal));
                            base = __shfl_sync(0xffffffffu, base, 0);
                            if (mk != 0u && c == cc) clist[cc * NBIG + base + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)k;
                        }
                    }
                }
                for (uint32_t k = r0 + warp; k < r1; k += 32) {      // bits of the cluster's slots
                    const uint32_t cs0 = bigl[k] & 0xFFFFu, m = bigl[k] >> 16;
                    for (uint32_t w = (cs0 >> 5) + lane; w <= ((cs0 + m - 1) >> 5); w += 32) {
                        uint32_t mask = 0xFFFFFFFFu;
                        if (w == (cs0 >> 5)) mask &= 0xFFFFFFFFu << (cs0 & 31u);
                        if (w == ((cs0 + m - 1) >> 5)) mask &= 0xFFFFFFFFu >> (31u - ((cs0 + m - 1) & 31u));
                        atomicOr(&flags[w], mask);
                    }
                }
                __syncthreads();
                if (tid == 992 && ms->ncls[0] <= 32u) {            // the team tier starts with its largest cluster
                    const uint32_t nt = ms->ncls[0];
                    for (uint32_t a = 1; a < nt; ++a) {
                        const uint16_t v = clist[a]; const uint32_t mv = bigl[v] >> 16;
                        uint32_t j = a;
                        while (j > 0 && (bigl[clist[j - 1]] >> 16) < mv) { clist[j] = clist[j - 1]; --j; }
                        clist[j] = v;
                    }
                }
                {   // rank prefix per word (u16 at pre16[0 .. 2048]: the work-list area is free now)
                    const uint32_t w0 = flags[2 * tid], w1 = flags[2 * tid + 1];
                    const uint32_t s2 = __popc(w0) + __popc(w1);
                    const uint32_t incl = warp_incl_scan_u32(s2);
                    if (lane == 31) ms->scan[warp] = incl;
                    __syncthreads();
                    if (warp == 0) { const uint32_t t = ms->scan[lane]; const uint32_t ti = warp_incl_scan_u32(t); ms->scan[lane] = ti - t; }
                    __syncthreads();
                    const uint32_t ex = ms->scan[warp] + incl - s2;
                    pre16[2 * tid] = (uint16_t)ex; pre16[2 * tid + 1] = (uint16_t)(ex + __popc(w0));
                }
                __syncthreads();
                SUBSTAMP(dt_f[0]);
                auto brank = [&](uint32_t u) -> uint32_t { return pre16[u >> 5] + __popc(flags[u >> 5] & ((1u << (u & 31u)) - 1u)); };
                for (uint32_t i0 = tid; i0 < len; i0 += 8 * NTHREADS) {
                    uint32_t tv[8];
#pragma unroll
                    for (uint32_t k = 0; k < 8; ++k) { const uint32_t i = i0 + k * NTHREADS; tv[k] = i < len ? tokb[i] : LONER; }
#pragma unroll
                    for (uint32_t k = 0; k < 8; ++k) {
                        if (tv[k] != LONER) {
                            const uint32_t u = tv[k] & 0xFFFFu;
                            if ((flags[u >> 5] >> (u & 31u)) & 1u) S[brank(u)] = (i0 + k * NTHREADS) | ((u - (tv[k] >> 16)) << 16);
                        }
                    }
                }
                __syncthreads();
                SUBSTAMP(dt_f[1]);
                {
                    Team<128> TT; const uint32_t team = warp >> 2;
                    TT.tt = tid & 127u; TT.bar = 1u + team; TT.scr = &ms->tscr[team][0]; TT.par = 0;
                    Team<32> TW; TW.tt = lane; TW.bar = 0; TW.scr = nullptr; TW.par = 0;
                    // clusters above TMIN entries by four-warp teams, largest first would be better still: list order
                    for (;;) {
                        if (TT.tt == 0) {
                            const uint32_t q = atomicAdd(&ms->next_t, 1u);
                            ms->tpick[team] = q < ms->ncls[0] ? (uint32_t)clist[q] : r1;
                        }
                        TT.sync();
                        const uint32_t qi = ms->tpick[team];
                        TT.sync();
                        if (qi >= r1) break;
                        const uint32_t cs0 = bigl[qi] & 0xFFFFu, m = bigl[qi] >> 16, kl = brank(cs0);
                        const long long t_c0 = CLK();
                        big_cluster<128>(TT, S + kl, E1 + kl, E2 + kl, m, cs0, data, fres, wscr + warp * 32, (DBG && dbg_stats) ? dbg_stats + (uint64_t)b * 136 + 40 : nullptr);
                        if (DBG && dbg_stats && TT.tt == 0) {
                            const uint32_t dtc = (uint32_t)((clock64() - t_c0) >> 6);
                            atomicMax(&dbg_stats[(uint64_t)b * 136 + 16], (dtc << 14) | m);
                            atomicAdd(&dbg_stats[(uint64_t)b * 136 + 24 + team], dtc);
                            atomicAdd(&dbg_stats[(uint64_t)b * 136 + 17], m);
                            atomicAdd(&dbg_stats[(uint64_t)b * 136 + 14], 1u);
                        }
                    }
                    for (;;) {                                   // L2MAX+1 .. TMIN entries: one warp each
                        uint32_t qi = 0;
                        if (lane == 0) { const uint32_t q = atomicAdd(&ms->rdone, 1u), na = ms->ncls[1]; qi = q < na ? (uint32_t)clist[NBIG + q] : q - na < ms->ncls[2] ? (uint32_t)clist[2 * NBIG + q - na] : r1; }
                        qi = __shfl_sync(0xffffffffu, qi, 0);
                        if (qi >= r1) break;
                        const uint32_t cs0 = bigl[qi] & 0xFFFFu, m = bigl[qi] >> 16, kl = brank(cs0);
                        big_cluster<32>(TW, S + kl, E1 + kl, E2 + kl, m, cs0, data, fres, wscr + warp * 32);
                    }
                    for (;;) {                                   // LMAX+1 .. L2MAX entries: 32 listed clusters per warp, the warp sorts, a lane simulates
                        uint32_t g = 0;
                        if (lane == 0) g = atomicAdd(&ms->next_w, 32u);
                        g = __shfl_sync(0xffffffffu, g, 0);
                        if (g >= ms->ncls[3]) break;
                        const uint32_t ent = g + lane < ms->ncls[3] ? bigl[clist[3 * NBIG + g + lane]] : 0u;
                        const uint32_t m = ent >> 16, cs0 = ent & 0xFFFFu;
                        const bool mine = m != 0u;
                        const uint32_t kl = mine ? brank(cs0) : 0u;
                        uint32_t coop = __ballot_sync(0xffffffffu, mine);
                        while (coop) {
                            const int srcl = __ffs(coop) - 1;
                            coop &= coop - 1;
                            warp_sort64(S + __shfl_sync(0xffffffffu, kl, srcl), __shfl_sync(0xffffffffu, m, srcl), __shfl_sync(0xffffffffu, cs0, srcl));
                        }
                        __syncwarp();
                        if (mine) lane_cluster<false>(S + kl, E1 + kl, E2 + kl, m, 0u, data, fres);
                    }
                }
                SUBSTAMP(dt_f[2]);
                __syncthreads();
                SUBSTAMP(dt_f[3]);
                r0 = r1;
            }
            break;
        }
        SUBSTAMP(dt_t);
        if (DBG && dbg_stats && tid == 0) {
            uint32_t* o = dbg_stats + (uint64_t)b * 136 + 8;
            o[0] = (uint32_t)dt_sc; o[1] = (uint32_t)dt_c2; o[2] = (uint32_t)dt_w; o[3] = (uint32_t)dt_t; o[4] = n_chunks; o[5] = n_wq; o[7] = (uint32_t)dt_wait;
            for (int k = 0; k < 4; ++k) dbg_stats[(uint64_t)b * 136 + 50 + k] = (uint32_t)dt_f[k];
        }
        __syncthreads();
        if (ms->fallback) {                                  // hand the block to lz77_v2_kernel
            if (tid == 0) fb_list[atomicAdd(fb_cnt, 1u)] = b;
            __syncthreads();
            continue;
        }
        PHASE_STAMP(4);
        // ---------------- P5: token candidates (reject rule deflate/lz77.c:222, extension :238-247), greedy parse as v2
        for (uint32_t i0 = tid; i0 < len; i0 += 8 * NTHREADS) {
            uint32_t fv[8];
#pragma unroll
            for (uint32_t k = 0; k < 8; ++k) { const uint32_t i = i0 + k * NTHREADS; fv[k] = i < len ? fres[i] : NONE16; }
#pragma unroll
            for (uint32_t k = 0; k < 8; ++k) {
                const uint32_t i = i0 + k * NTHREADS;
                if (i < len) {
                    uint32_t tk = 0;
                    if (fv[k] != NONE16 && i - fv[k] < W - 1u) tk = (i - fv[k]) | (match_len(data, fv[k], i) << 16);
                    tokb[i] = tk;
                    if (dbg_tok) dbg_tok[(uint64_t)b * MAXB + i] = tk;
                    adv[PADX(i)] = (uint8_t)(tk ? (tk >> 16) : 1u);
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
            uint32_t e = lane;
            for (uint32_t k = 0; k < 32; ++k) {
                const uint32_t ch = warp * 32 + k;
                const uint32_t p = (ch << 6) + e;
                if (ch < nchunks && p < len && e < 31) e = exitof[PADX(p)];
            }
            ms->sexit[warp][lane] = (uint8_t)e;
        }
        __syncthreads();
        if (tid == 0) {
            uint32_t e = 0;
            for (uint32_t s = 0; s < 32; ++s) { ms->sentry[s] = (uint8_t)e; e = ms->sexit[s][e & 31u]; }
        }
        __syncthreads();
        uint8_t* centry = reinterpret_cast<uint8_t*>(pre) + 8192;
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
        uint8_t* orel = exitof;
        for (uint32_t i = tid; i < (PADDED >> 2); i += NTHREADS) reinterpret_cast<uint32_t*>(orel)[i] = 0xFFFFFFFFu;
        __syncthreads();
        uint32_t my_units = 0;
        if (tid < nchunks) {
            const uint32_t cend = (tid << 6) + 64;
            const uint32_t hi = cend < len ? cend : len;
            for (uint32_t p = (tid << 6) + centry[tid]; p < hi; p += adv[PADX(p)]) {
                orel[PADX(p)] = (uint8_t)(my_units >> 1);
                my_units += adv[PADX(p)] == 1 ? 2u : 4u;
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
        PHASE_STAMP(5);
        // ---------------- P6: emission, one position per thread (deflate/lz77.c:176-197)
        {
            uint32_t* coff = pre + 3072;
            if (tid < nchunks) coff[tid] = my_off;
            __syncthreads();
            for (uint32_t p0 = tid; p0 < len; p0 += 4 * NTHREADS) {
                uint32_t tk[4];
#pragma unroll
                for (uint32_t k = 0; k < 4; ++k) { const uint32_t p = p0 + k * NTHREADS; tk[k] = p < len ? tokb[p] : 0u; }
#pragma unroll
                for (uint32_t k = 0; k < 4; ++k) {
                    const uint32_t p = p0 + k * NTHREADS;
                    if (p >= len) break;
                    const uint32_t r = orel[PADX(p)];
                    if (r == 0xFFu) continue;
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
        }
        __syncthreads();
        PHASE_STAMP(6);
    }
}

// Which match finder? lz77_v4_kernel is 15 % faster on text, 4 % on near-random input and slower on few-symbol input (a block that
// is one long chain is handed back after its occupancy pass). One CTA estimates the byte entropy of 32 768 bytes spread over the
// buffer: from 1.5 bits per byte on the input goes to lz77_v4_kernel.
__global__ void __launch_bounds__(1024) lz77_pick_kernel(const uint8_t* __restrict__ in, uint64_t n, uint32_t* __restrict__ flag) {
    __shared__ uint32_t hist[256];
    __shared__ float part[32];
    const uint32_t tid = threadIdx.x;
    if (tid < 256) hist[tid] = 0;
    __syncthreads();
    const uint64_t S = n < 32768 ? n : 32768;
    const uint64_t stride = n / S;
    for (uint64_t i = tid; i < S; i += 1024) atomicAdd(&hist[__ldg(in + i * stride)], 1u);
    __syncthreads();
    float h = 0.f;
    if (tid < 256) { const float c = (float)hist[tid]; if (c > 0.f) h = c * __log2f(c); }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) h += __shfl_xor_sync(0xffffffffu, h, d);
    if ((tid & 31) == 0) part[tid >> 5] = h;
    __syncthreads();
    if (tid == 0) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += part[w];
        const float H = __log2f((float)S) - t / (float)S;
        *flag = H >= 1.5f ? 1u : 0u;
    }
}

}  // namespace

int lz77_v2_launch(b200_ctx* ctx, int variant, const uint8_t* d_in, uint64_t n, uint64_t bs, uint64_t nblocks,
                   uint8_t* scratch, uint64_t stride, uint64_t* d_block_sizes, uint64_t* block_bytes, uint32_t* dbg_tok,
                   const uint32_t* blist, const uint32_t* bcount);

// scratch slots: 13 = F (u16 per position; v2's lists when it runs the handed-back blocks), 14 = tok, 18 = hand-back list
// mode: 1 = lz77_v4_kernel for every block it can take, 2 = decided by a sample of the input (lz77_pick_kernel)
int lz77_v4_launch(b200_ctx* ctx, const uint8_t* d_in, uint64_t n, uint64_t bs, uint64_t nblocks,
                   uint8_t* scratch, uint64_t stride, uint64_t* d_block_sizes, uint64_t* block_bytes, uint32_t* dbg_tok, int mode) {
    static bool attr_done_dev[64] = {};
    bool& attr_done = attr_done_dev[ctx->device >= 0 && ctx->device < 64 ? ctx->device : 0];
    if (!attr_done) {
        CUDA_TRY(cudaFuncSetAttribute(lz77_v4_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        CUDA_TRY(cudaFuncSetAttribute(lz77_v4_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        attr_done = true;
    }
    uint64_t grid = (uint64_t)ctx->sm_count;
    if (grid > nblocks) grid = nblocks;
    uint16_t* fres; uint32_t* tok; uint32_t* fb;
    B200_TRY(b200_scratch(ctx, B200_SLOT(ctx, 13), (size_t)grid * MAXB * 4 + 64, reinterpret_cast<void**>(&fres)));
    B200_TRY(b200_scratch(ctx, B200_SLOT(ctx, 14), (size_t)grid * MAXB * 4 + 64, reinterpret_cast<void**>(&tok)));
    B200_TRY(b200_scratch(ctx, B200_SLOT(ctx, 18), (size_t)(nblocks + 16) * 4, reinterpret_cast<void**>(&fb)));
    CUDA_TRY(cudaMemsetAsync(fb, 0, 16, ctx->stream));
    uint32_t* fb_cnt = fb; uint32_t* fb_list = fb + 4;
    uint32_t* use_flag = nullptr;
    if (mode == 2) {
        use_flag = fb + 1;
        lz77_pick_kernel<<<1, 1024, 0, ctx->stream>>>(d_in, n, use_flag);
        ctx->launches += 1;
    }
    if (dbg_tok) lz77_v4_kernel<true><<<(unsigned)grid, NTHREADS, SMEM_BYTES, ctx->stream>>>(d_in, n, (uint32_t)bs, (uint32_t)nblocks, fres, tok, scratch, stride, d_block_sizes, block_bytes, fb_list, fb_cnt, use_flag, dbg_tok);
    else lz77_v4_kernel<false><<<(unsigned)grid, NTHREADS, SMEM_BYTES, ctx->stream>>>(d_in, n, (uint32_t)bs, (uint32_t)nblocks, fres, tok, scratch, stride, d_block_sizes, block_bytes, fb_list, fb_cnt, use_flag, dbg_tok);
    CUDA_TRY(cudaGetLastError());
    ctx->launches += 1;
    if (getenv("B200_LZ_V4_SYNC")) {           // diagnostic: separate a fault of this kernel from one of the hand-back launch
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        uint32_t hcnt = 0;
        CUDA_TRY(cudaMemcpy(&hcnt, fb_cnt, 4, cudaMemcpyDeviceToHost));
        fprintf(stderr, "lz77_v4_kernel done: %u of %llu blocks handed back\n", hcnt, (unsigned long long)nblocks);
    }
    // the blocks handed back (empty list: the kernel's CTAs return at once)
    return lz77_v2_launch(ctx, 1, d_in, n, bs, nblocks, scratch, stride, d_block_sizes, block_bytes, dbg_tok, fb_list, fb_cnt);
}
