// Huffman pieces shared by the byte codec (huffman.cu) and the deflate token entropy stage
// (deflate_huff.cu): the exact heap-order table build and the scans that place chunk streams.
#pragma once
#include "common.cuh"

// (static: one private copy per translation unit)

// ---------------------------------------------------------------- K2 table build
// One warp per block. Lane 0 replays the reference's binary min-heap exactly
// (enqueue in symbol order, strict '<' sift-up, left-then-right sift-down, dequeue
// moves the last element to the root); afterwards the 32 lanes walk leaf->root in
// parallel to produce codes and lengths.
// NSYM symbols per table (256: algorithms/huffman; 286: NUM_CODES of algorithms/deflate/huffman.h:6),
// rows of STRIDE entries in freq/codes/lens, 2*NSYM-1 nodes per tree. SINGLE_OK: a table with one
// distinct symbol gets the 1-bit code 0 (an inner root over that leaf) instead of status 1.
// ZIGPQ: the sift-down of Zig's std.PriorityQueue instead of the reference C heap's (algorithms/huffman/zig_huffman/
// src/main.zig:133-156 builds its tree with it): the lesser child is the right one only when strictly smaller than the
// left, and the moved element stops only when STRICTLY smaller than that child (so it keeps sinking through equal
// frequencies, where heapify_down of huffman.c:112-131 stops). Sift-up is the same strict '<' in both.
template <int NSYM, int STRIDE, bool SINGLE_OK, bool ZIGPQ = false>
static __global__ void __launch_bounds__(32) huff_build_kernel(const uint32_t* __restrict__ freq, uint32_t* __restrict__ codes,
                                                       uint8_t* __restrict__ lens, int16_t* __restrict__ tree,
                                                       uint32_t* __restrict__ meta) {
    constexpr int NODES = 2 * NSYM - 1;
    __shared__ uint32_t fr[NSYM];
    __shared__ uint32_t nfreq[NODES];
    __shared__ int16_t  left[NODES], right[NODES], parent[NODES];
    __shared__ uint8_t  isright[NODES];
    __shared__ uint64_t hq[NSYM];   // heap entry = frequency << 32 | node: one load / store per element; only the frequency is compared
    __shared__ int s_nn, s_root, s_distinct;
    const uint64_t b = blockIdx.x;
    const unsigned lane = threadIdx.x;
    for (int s = lane; s < NSYM; s += 32) fr[s] = freq[b * STRIDE + s];
    __syncwarp();
    if (lane == 0) {
#define HQ_F(x) ((uint32_t)((x) >> 32))
        auto sift_up = [&](int idx) {
            const uint64_t x = hq[idx];
            while (idx > 0) {
                const int p = (idx - 1) >> 1;
                const uint64_t y = hq[p];
                if (!(HQ_F(x) < HQ_F(y))) break;
                hq[idx] = y;              // the swaps of heapify_up move x up and each parent down
                idx = p;
            }
            hq[idx] = x;
        };
        int size = 0, nn = 0;
        for (int s = 0; s < NSYM; ++s) {
            const uint32_t f = fr[s];
            if (!f) continue;
            nfreq[nn] = f; left[nn] = -1; right[nn] = (int16_t)s; parent[nn] = -1; isright[nn] = 0;
            hq[size] = ((uint64_t)f << 32) | (uint32_t)nn; ++nn;
            sift_up(size++);
        }
        const int distinct = nn;
        while (size > 1) {
            int pick[2];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                pick[q] = (int)(uint32_t)hq[0];
                --size;
                const uint64_t x = hq[size];   // dequeue: the last element goes to the root, then sifts down
                int idx = 0;
                for (;;) {
                    const int l = 2 * idx + 1, r = l + 1;
                    if (l >= size) break;
                    const uint64_t yl = hq[l], yr = r < size ? hq[r] : 0xFFFFFFFFFFFFFFFFull;
                    int sm = idx; uint64_t ys = x;
                    if (ZIGPQ) {
                        sm = l; ys = yl;
                        if (r < size && HQ_F(yr) < HQ_F(yl)) { sm = r; ys = yr; }
                        if (HQ_F(x) < HQ_F(ys)) break;
                    } else {
                        // smallest of (x, left, right) with the reference's order of comparisons (strict <, left first)
                        if (HQ_F(yl) < HQ_F(ys)) { sm = l; ys = yl; }
                        if (r < size && HQ_F(yr) < HQ_F(ys)) { sm = r; ys = yr; }
                        if (sm == idx) break;
                    }
                    hq[idx] = ys;
                    idx = sm;
                }
                if (size > 0) hq[idx] = x;
            }
            const uint32_t f = nfreq[pick[0]] + nfreq[pick[1]];  // u32 wrap like init_node(uint32_t)
            nfreq[nn] = f; left[nn] = (int16_t)pick[0]; right[nn] = (int16_t)pick[1]; parent[nn] = -1; isright[nn] = 0;
            parent[pick[0]] = (int16_t)nn; isright[pick[0]] = 0;
            parent[pick[1]] = (int16_t)nn; isright[pick[1]] = 1;
            hq[size] = ((uint64_t)f << 32) | (uint32_t)nn; ++nn;
            sift_up(size++);
        }
        int root = distinct ? (int)(uint32_t)hq[0] : 0;
        if (SINGLE_OK && distinct == 1) {   // one leaf: an inner root with that leaf on both sides -> code 0, length 1
            left[1] = 0; right[1] = 0; parent[1] = -1; isright[1] = 0; parent[0] = 1; isright[0] = 0;
            nn = 2; root = 1;
        }
        s_nn = nn; s_distinct = distinct; s_root = root;
    }
    __syncwarp();
    const int nn = s_nn, distinct = s_distinct, root = s_root;
    uint32_t maxlen = 0;
    for (int v = lane; v < distinct; v += 32) {
        uint32_t code = 0, len = 0;
        int u = v;
        while (u != root) {
            if (len < 32) code |= (uint32_t)isright[u] << len;
            ++len;
            u = parent[u];
        }
        const int sym = right[v];
        codes[b * STRIDE + sym] = code;
        lens[b * STRIDE + sym] = (uint8_t)(len > 255 ? 255 : len);
        maxlen = max(maxlen, len);
    }
    for (int v = lane; v < nn; v += 32) {
        tree[(b * NODES + v) * 2 + 0] = left[v];
        tree[(b * NODES + v) * 2 + 1] = right[v];
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) maxlen = max(maxlen, __shfl_xor_sync(0xffffffffu, maxlen, d));
    if (lane == 0) {
        uint32_t status = 0;
        if (distinct < (SINGLE_OK ? 1 : 2)) status = 1;       // reference: exit(1), huffman.c:278-281 / :149-152
        else if (maxlen > 32) status = 2;   // reference: silent corruption (uint32_t code, BIT_MASK[33])
        meta[b * 4 + 0] = status; meta[b * 4 + 1] = (uint32_t)distinct; meta[b * 4 + 2] = (uint32_t)root; meta[b * 4 + 3] = maxlen;
    }
}

// ---------------------------------------------------------------- K4 offsets (single CTA)
// Exclusive scans that place every chunk: P = scan(chunk_bits); per block
// bits -> ceil/32 words -> scan -> first word; absolute bit offset per chunk.
// Also zeroes the words shared by two chunks so the encoder can OR into them.
static __global__ void __launch_bounds__(1024) huff_offsets_kernel(const uint32_t* __restrict__ chunk_bits, uint64_t nchunks,
                                                            uint32_t cpb, uint64_t nblocks, uint64_t* __restrict__ P,
                                                            uint64_t* __restrict__ block_bits, uint64_t* __restrict__ block_word,
                                                            uint64_t words_capacity, uint64_t* __restrict__ info) {
    __shared__ uint64_t warp_tot[33];
    const uint64_t total_bits = cta_exscan_1024(nchunks, warp_tot,
        [&](uint64_t i) { return (uint64_t)chunk_bits[i]; }, [&](uint64_t i, uint64_t ex) { P[i] = ex; });
    if (threadIdx.x == 0) P[nchunks] = total_bits;
    __syncthreads();
    const uint64_t total_words = cta_exscan_1024(nblocks, warp_tot,
        [&](uint64_t b) {
            const uint64_t first = b * cpb;
            uint64_t last = first + cpb; if (last > nchunks) last = nchunks;
            const uint64_t bits = P[last] - P[first];
            block_bits[b] = bits;
            return (bits + 31) >> 5;
        },
        [&](uint64_t b, uint64_t ex) { block_word[b] = ex; });
    if (threadIdx.x == 0) {
        block_word[nblocks] = total_words;
        info[0] = total_words;
        info[1] = total_words > words_capacity ? 1 : 0;
    }
}

// absolute bit offset of every chunk; zeroes the words shared by two chunks so the encoder can OR into them
static __global__ void __launch_bounds__(256) huff_chunkoff_kernel(uint64_t nchunks, uint32_t cpb, const uint64_t* __restrict__ P,
                                                            const uint64_t* __restrict__ block_word, uint64_t* __restrict__ chunk_off,
                                                            uint32_t* __restrict__ words, const uint64_t* __restrict__ info) {
    const uint64_t c = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    if (c >= nchunks) return;
    const uint64_t b = c / cpb;
    const uint64_t off = block_word[b] * 32 + (P[c] - P[b * cpb]);
    chunk_off[c] = off;
    if (!info[1] && (off & 31)) words[off >> 5] = 0;
}

