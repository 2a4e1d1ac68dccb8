// Token-parallel LZ77 decoder for streams of ANY block size, including one block = the whole buffer (what the
// drop-in lz77_decompress of algorithms/lz77/lz77.c:347-377 is handed): no warp owns a block, every kernel is
// parallel over the whole stream.
//
//  K1 chunk tables   The bit stream of a block is cut into chunks of 2048 bits. A token is 9 or 19 bits (flag bit 0:
//                    literal, 1: 14-bit offset + 4-bit length, lz77.c:359-372), so a token sequence can enter a chunk at
//                    bit offset 0..18. One thread per chunk runs a backward pass over its bit positions: exit(p) = where
//                    the token sequence that starts at p leaves the chunk, bytes(p) = output bytes it produces on the
//                    way; kept for the 19 possible entries.
//  K2 composition    One warp per block walks its chunks in order (tables staged 32 at a time in shared memory):
//                    entry offset and output position of every chunk.
//  K3 emission       One thread per chunk walks its tokens: a literal stores its byte and marks the position LIT; byte k
//                    of a match stores the POINTER (position - offset) of the byte it copies (the byte-serial copy of
//                    lz77.c:364-366 reads what it has just written when length > offset: the pointer chain does the same).
//  K4 resolution     One thread per output byte follows its pointer chain to a LIT byte (at most 16 hops per round; a
//                    longer chain leaves the ancestor it reached, so the depth shrinks 16 x per round) and copies that byte.
//
// Deflate variant (2- or 4-byte tokens, deflate/lz77.c:176-197): same K2 / K4, the chunk table has two entries
// (does the chunk start inside a match token or not).
#include "common.cuh"
#include "../../include/b200comp.h"

namespace {

constexpr uint32_t CB = 2048;                  // bits per chunk (standalone variant)
constexpr uint32_t NENT = 19;                  // entry offsets 0..18
constexpr uint32_t TAB_WORDS = 20;             // per chunk: 19 x (exit << 24 | bytes), padded
constexpr uint32_t END = 31;                   // exit marker: the token sequence ends inside the chunk
constexpr uint32_t LIT = 0x80000000u;          // P[pos]: the byte at pos is final and a root
constexpr uint32_t RESOLVED = 0x40000000u;     // P[pos]: out[pos] has been written (pointer = its root)
constexpr uint32_t PMASK = 0x3FFFFFFFu;
constexpr uint32_t CU = 1024;                  // 2-byte units per chunk (deflate variant)

__device__ __forceinline__ uint32_t ldw(const uint32_t* w, uint64_t i, uint64_t nwords) { return i < nwords ? __ldg(w + i) : 0u; }

// A thread emits the pointer words of consecutive output positions: four of them (one aligned 16-byte group) are gathered
// in registers and stored at once; a group it shares with the neighbouring chunk's thread goes out word by word.
struct PWriter {
    uint32_t* P; uint64_t grp; uint32_t v0, v1, v2, v3, mask;
    __device__ __forceinline__ void init(uint32_t* p) { P = p; grp = ~0ull; mask = 0; v0 = v1 = v2 = v3 = 0; }
    __device__ __forceinline__ void flush() {
        if (mask == 15u) *reinterpret_cast<uint4*>(P + (grp << 2)) = make_uint4(v0, v1, v2, v3);
        else {
            if (mask & 1u) P[(grp << 2) + 0] = v0;
            if (mask & 2u) P[(grp << 2) + 1] = v1;
            if (mask & 4u) P[(grp << 2) + 2] = v2;
            if (mask & 8u) P[(grp << 2) + 3] = v3;
        }
        mask = 0;
    }
    __device__ __forceinline__ void put(uint64_t pos, uint32_t val) {
        const uint64_t g = pos >> 2;
        if (g != grp) { if (mask) flush(); grp = g; }
        const uint32_t k = (uint32_t)pos & 3u;
        if (k == 0) v0 = val; else if (k == 1) v1 = val; else if (k == 2) v2 = val; else v3 = val;
        mask |= 1u << k;
    }
    __device__ __forceinline__ void done() { if (mask) flush(); }
};

// ---------------------------------------------------------------- K1 (standalone variant)
__global__ void __launch_bounds__(128) pdec_tables_v0_kernel(const uint8_t* __restrict__ stream, const uint64_t* __restrict__ block_off,
                                                            const uint64_t* __restrict__ block_bits, uint64_t nblocks, uint32_t cpb,
                                                            uint32_t* __restrict__ tables) {
    __shared__ uint8_t  ex[32][128];
    __shared__ uint16_t os[32][128];
    const uint64_t c = (uint64_t)blockIdx.x * 128 + threadIdx.x;
    const uint64_t b = c / cpb;
    if (b >= nblocks) return;
    const uint32_t ci = (uint32_t)(c % cpb);
    const uint64_t T = block_bits[b];
    const uint64_t cstart = (uint64_t)ci * CB;
    if (cstart >= T) return;
    const uint8_t* tk = stream + block_off[b];
    const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(tk) & 3);
    const uint32_t* wbase = reinterpret_cast<const uint32_t*>(tk - mis);
    const uint64_t nwords = ((uint64_t)mis * 8 + T + 31) >> 5;
    const uint32_t t = threadIdx.x;
    const uint32_t plim = (uint32_t)(T - cstart < CB ? T - cstart : CB);      // positions >= plim start no token
    // every slot of the ring that can be read before it is written: END
#pragma unroll
    for (uint32_t k = 0; k < 32; ++k) { ex[k][t] = (uint8_t)END; os[k][t] = 0; }
    const uint64_t q0 = (uint64_t)mis * 8 + cstart;                          // bit index of chunk position 0 in the aligned words
    uint64_t wi = (q0 + plim - 1) >> 5;
    uint32_t w0 = ldw(wbase, wi, nwords), w1 = ldw(wbase, wi + 1, nwords);
    for (uint32_t p = plim; p-- > 0;) {
        const uint64_t q = q0 + p;
        if ((q >> 5) != wi) { wi = q >> 5; w1 = w0; w0 = ldw(wbase, wi, nwords); }
        const uint32_t v = __funnelshift_r(w0, w1, (uint32_t)(q & 31));
        const uint32_t flag = v & 1u, len = flag ? 19u : 9u;
        uint32_t e = END, o = 0;
        if (cstart + p + len <= T) {
            const uint32_t ob = flag ? (v >> 15) & 15u : 1u;
            const uint32_t nx = p + len;
            if (nx >= CB) { e = nx - CB; o = ob; }
            else if (nx >= plim) { e = END; o = ob; }                        // the stream ends right behind this token
            else { e = ex[nx & 31][t]; o = ob + os[nx & 31][t]; }
        }
        ex[p & 31][t] = (uint8_t)e; os[p & 31][t] = (uint16_t)o;
    }
    uint32_t* tab = tables + c * TAB_WORDS;
#pragma unroll
    for (uint32_t k = 0; k < NENT; ++k) tab[k] = k < plim ? ((uint32_t)ex[k][t] << 24) | os[k][t] : (END << 24);
}

// ---------------------------------------------------------------- K1 (deflate variant)
// unit u = bytes 2u, 2u+1 of the block's token stream. state 0: u starts a token, state 1: u is the tail of a match token.
// table[s] = exit state << 24 | output bytes of the tokens that START in the chunk, for entry state s.
__global__ void __launch_bounds__(128) pdec_tables_v1_kernel(const uint8_t* __restrict__ stream, const uint64_t* __restrict__ block_off,
                                                            const uint64_t* __restrict__ block_bytes, uint64_t nblocks, uint32_t cpb,
                                                            uint32_t* __restrict__ tables) {
    const uint64_t c = (uint64_t)blockIdx.x * 128 + threadIdx.x;
    const uint64_t b = c / cpb;
    if (b >= nblocks) return;
    const uint32_t ci = (uint32_t)(c % cpb);
    const uint64_t U = block_bytes[b] >> 1;                                   // units of the block
    const uint64_t u0 = (uint64_t)ci * CU;
    if (u0 >= U) return;
    const uint8_t* tk = stream + block_off[b];
    const uint32_t ulim = (uint32_t)(U - u0 < CU ? U - u0 : CU);
    // both entry states in one pass over the units (they read the same bytes; after the first unit both walks see a token
    // start they are the same walk)
    uint32_t res[2];
    uint32_t sa = 0, sb = 1, oa = 0, ob = 0;
    for (uint32_t u = 0; u < ulim; ++u) {
        const uint8_t f = __ldg(tk + 2 * (u0 + u));
        uint32_t ml = 0;
        const bool head_ok = f != 0 && u0 + u + 1 < U;
        if (head_ok && (!sa || !sb)) ml = __ldg(tk + 2 * (u0 + u) + 3);
        if (sa) sa = 0; else if (f) { if (head_ok) { oa += ml; sa = 1; } } else oa += 1;
        if (sb) sb = 0; else if (f) { if (head_ok) { ob += ml; sb = 1; } } else ob += 1;
    }
    res[0] = (sa << 24) | oa; res[1] = (sb << 24) | ob;
    tables[c * TAB_WORDS + 0] = res[0]; tables[c * TAB_WORDS + 1] = res[1];
}

// ---------------------------------------------------------------- K2: entry state and output position of every chunk
__global__ void __launch_bounds__(128) pdec_compose_kernel(const uint32_t* __restrict__ tables, uint64_t nblocks, uint32_t cpb,
                                                          uint32_t nent, uint64_t* __restrict__ centry) {
    __shared__ uint32_t st[4][32][TAB_WORDS];
    const uint32_t lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    const uint64_t b = (uint64_t)blockIdx.x * 4 + wp;
    if (b >= nblocks) return;
    uint32_t e = 0; uint64_t opos = 0;                                       // carried by lane 0
    for (uint32_t g = 0; g < cpb; g += 32) {
        const uint32_t nc = cpb - g < 32 ? cpb - g : 32;
        for (uint32_t i = lane; i < nc * TAB_WORDS; i += 32) st[wp][i / TAB_WORDS][i % TAB_WORDS] = tables[(b * cpb + g) * TAB_WORDS + i];
        __syncwarp();
        if (lane == 0) {
            for (uint32_t k = 0; k < nc; ++k) {
                centry[b * cpb + g + k] = ((uint64_t)e << 56) | opos;
                if (e != END) {
                    const uint32_t x = e < nent ? st[wp][k][e] : (END << 24);
                    opos += x & 0xFFFFFFu;
                    e = x >> 24;
                }
            }
        }
        __syncwarp();
        e = __shfl_sync(0xffffffffu, e, 0);
        if (e == END) {                                                      // nothing starts in the remaining chunks
            for (uint32_t k = g + 32 + lane; k < cpb; k += 32) centry[b * cpb + k] = (uint64_t)END << 56;
            break;
        }
    }
}

// ---------------------------------------------------------------- K2 for blocks of many chunks: two levels.
// Segments of SEG chunks: (a) one thread per (segment, entry state) composes the segment's chunk tables; (b) one thread per
// block walks the segment tables; (c) one thread per segment walks its chunks from the entry (b) found.
constexpr uint32_t SEG = 256;
__global__ void __launch_bounds__(128) pdec_seg_tables_kernel(const uint32_t* __restrict__ tables, uint64_t nblocks, uint32_t cpb,
                                                             uint32_t spb, uint32_t nent, uint32_t* __restrict__ seg_tables, uint64_t* __restrict__ seg_bytes) {
    const uint64_t t = (uint64_t)blockIdx.x * 128 + threadIdx.x;
    const uint64_t sg = t / nent;
    const uint32_t e0 = (uint32_t)(t % nent);
    if (sg >= nblocks * spb) return;
    const uint64_t b = sg / spb;
    const uint32_t si = (uint32_t)(sg % spb);
    const uint32_t c0 = si * SEG, c1 = c0 + SEG < cpb ? c0 + SEG : cpb;
    uint32_t e = e0; uint64_t o = 0;
    for (uint32_t c = c0; c < c1 && e != END; ++c) {
        const uint32_t x = e < nent ? __ldg(tables + (b * cpb + c) * TAB_WORDS + e) : (END << 24);
        o += x & 0xFFFFFFu;
        e = x >> 24;
    }
    seg_tables[sg * TAB_WORDS + e0] = e;
    seg_bytes[sg * TAB_WORDS + e0] = o;
}
__global__ void __launch_bounds__(128) pdec_seg_walk_kernel(const uint32_t* __restrict__ seg_tables, const uint64_t* __restrict__ seg_bytes,
                                                           uint64_t nblocks, uint32_t spb, uint32_t nent, uint64_t* __restrict__ seg_entry) {
    const uint64_t b = (uint64_t)blockIdx.x * 128 + threadIdx.x;
    if (b >= nblocks) return;
    uint32_t e = 0; uint64_t o = 0;
    for (uint32_t si = 0; si < spb; ++si) {
        seg_entry[b * spb + si] = ((uint64_t)e << 56) | o;
        if (e != END) {
            const uint64_t sg = b * spb + si;
            const uint32_t e2 = e < nent ? seg_tables[sg * TAB_WORDS + e] : END;
            o += e < nent ? seg_bytes[sg * TAB_WORDS + e] : 0;
            e = e2;
        }
    }
}
__global__ void __launch_bounds__(128) pdec_seg_fill_kernel(const uint32_t* __restrict__ tables, const uint64_t* __restrict__ seg_entry,
                                                           uint64_t nblocks, uint32_t cpb, uint32_t spb, uint32_t nent, uint64_t* __restrict__ centry) {
    const uint64_t sg = (uint64_t)blockIdx.x * 128 + threadIdx.x;
    if (sg >= nblocks * spb) return;
    const uint64_t b = sg / spb;
    const uint32_t si = (uint32_t)(sg % spb);
    const uint32_t c0 = si * SEG, c1 = c0 + SEG < cpb ? c0 + SEG : cpb;
    const uint64_t se = seg_entry[sg];
    uint32_t e = (uint32_t)(se >> 56); uint64_t o = se & 0xFFFFFFFFFFFFFFull;
    for (uint32_t c = c0; c < c1; ++c) {
        centry[b * cpb + c] = ((uint64_t)e << 56) | o;
        if (e != END) {
            const uint32_t x = e < nent ? __ldg(tables + (b * cpb + c) * TAB_WORDS + e) : (END << 24);
            o += x & 0xFFFFFFu;
            e = x >> 24;
        }
    }
}

// ---------------------------------------------------------------- K3 (standalone variant)
__global__ void __launch_bounds__(128) pdec_emit_v0_kernel(const uint8_t* __restrict__ stream, const uint64_t* __restrict__ block_off,
                                                          const uint64_t* __restrict__ block_bits, uint64_t n, uint64_t bs, uint64_t nblocks,
                                                          uint32_t cpb, const uint64_t* __restrict__ centry, uint8_t* __restrict__ out,
                                                          uint32_t* __restrict__ P) {
    const uint64_t c = (uint64_t)blockIdx.x * 128 + threadIdx.x;
    const uint64_t b = c / cpb;
    if (b >= nblocks) return;
    const uint32_t ci = (uint32_t)(c % cpb);
    const uint64_t T = block_bits[b];
    const uint64_t cstart = (uint64_t)ci * CB;
    if (cstart >= T) return;
    const uint64_t ce = centry[c];
    uint32_t p = (uint32_t)(ce >> 56);
    if (p == END) return;
    uint64_t o = ce & 0xFFFFFFFFFFFFFFull;
    const uint64_t base = b * bs;
    const uint64_t len = n - base < bs ? n - base : bs;
    const uint8_t* tk = stream + block_off[b];
    const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(tk) & 3);
    const uint32_t* wbase = reinterpret_cast<const uint32_t*>(tk - mis);
    const uint64_t nwords = ((uint64_t)mis * 8 + T + 31) >> 5;
    const uint64_t q0 = (uint64_t)mis * 8 + cstart;
    PWriter pw; pw.init(P);
    while (p < CB && o < len) {
        const uint64_t q = q0 + p;
        const uint32_t v = __funnelshift_r(ldw(wbase, q >> 5, nwords), ldw(wbase, (q >> 5) + 1, nwords), (uint32_t)(q & 31));
        const uint32_t flag = v & 1u, tl = flag ? 19u : 9u;
        if (cstart + p + tl > T) break;
        if (!flag) { out[base + o] = (uint8_t)(v >> 1); pw.put(base + o, LIT); ++o; }
        else {
            const uint32_t off = (v >> 1) & 0x3FFFu, ml = (v >> 15) & 15u;
            const bool ok = off != 0u && off <= o;                           // otherwise nothing is copied (as the serial decoders)
            for (uint32_t k = 0; k < ml && o + k < len; ++k) pw.put(base + o + k, ok ? (uint32_t)(o + k - off) : LIT);
            o += ml;
        }
        p += tl;
    }
    pw.done();
}

// ---------------------------------------------------------------- K3 (deflate variant)
__global__ void __launch_bounds__(128) pdec_emit_v1_kernel(const uint8_t* __restrict__ stream, const uint64_t* __restrict__ block_off,
                                                          const uint64_t* __restrict__ block_bytes, uint64_t n, uint64_t bs, uint64_t nblocks,
                                                          uint32_t cpb, const uint64_t* __restrict__ centry, uint8_t* __restrict__ out,
                                                          uint32_t* __restrict__ P) {
    const uint64_t c = (uint64_t)blockIdx.x * 128 + threadIdx.x;
    const uint64_t b = c / cpb;
    if (b >= nblocks) return;
    const uint32_t ci = (uint32_t)(c % cpb);
    const uint64_t U = block_bytes[b] >> 1;
    const uint64_t u0 = (uint64_t)ci * CU;
    if (u0 >= U) return;
    const uint64_t ce = centry[c];
    uint32_t s = (uint32_t)(ce >> 56);
    if (s == END) return;
    uint64_t o = ce & 0xFFFFFFFFFFFFFFull;
    const uint64_t base = b * bs;
    const uint64_t len = n - base < bs ? n - base : bs;
    const uint8_t* tk = stream + block_off[b];
    const uint32_t ulim = (uint32_t)(U - u0 < CU ? U - u0 : CU);
    PWriter pw; pw.init(P);
    for (uint32_t u = s; u < ulim && o < len; ++u) {
        const uint8_t* t = tk + 2 * (u0 + u);
        if (__ldg(t)) {
            if (u0 + u + 1 >= U) break;
            const uint32_t off = (uint32_t)__ldg(t + 1) | ((uint32_t)__ldg(t + 2) << 8), ml = __ldg(t + 3);
            const bool ok = off != 0u && off <= o;
            for (uint32_t k = 0; k < ml && o + k < len; ++k) pw.put(base + o + k, ok ? (uint32_t)(o + k - off) : LIT);
            o += ml;
            ++u;
        } else { out[base + o] = __ldg(t + 1); pw.put(base + o, LIT); ++o; }
    }
    pw.done();
}

// ---------------------------------------------------------------- K4: pointer chains
// Grid-stride, four consecutive bytes per thread and step (the bytes of one match point at consecutive sources: the same
// sectors). Reads of P may be stale: any value it ever held is an ancestor of the position, so the chain stays valid.
__global__ void __launch_bounds__(256) pdec_resolve_kernel(uint8_t* __restrict__ out, uint32_t* __restrict__ P, uint64_t n, uint64_t bs,
                                                          const uint32_t* __restrict__ pending_in, uint32_t* __restrict__ pending_out) {
    if (pending_in && *pending_in == 0) return;
    bool left = false;
    const uint64_t nq = (n + 3) >> 2;
    for (uint64_t qd = (uint64_t)blockIdx.x * 256 + threadIdx.x; qd < nq; qd += (uint64_t)gridDim.x * 256) {
        const uint64_t pos0 = qd << 2;
        uint32_t x[4];
        if (pos0 + 4 <= n) { const uint4 v = *reinterpret_cast<const uint4*>(P + pos0); x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w; }
        else { for (int k = 0; k < 4; ++k) x[k] = pos0 + k < n ? P[pos0 + k] : LIT; }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (x[k] & (LIT | RESOLVED)) continue;
            const uint64_t pos = pos0 + k;
            const uint64_t base = pos / bs * bs;
            uint32_t p = x[k] & PMASK;
            bool done = false;
#pragma unroll 1
            for (int step = 0; step < 16; ++step) {
                const uint32_t y = P[base + p];
                if (y & LIT) { out[pos] = out[base + p]; P[pos] = p | RESOLVED; done = true; break; }
                p = y & PMASK;
            }
            if (!done) { P[pos] = p; left = true; }
        }
    }
    if (__syncthreads_or(left) && threadIdx.x == 0) atomicAdd(pending_out, 1u);
}

}  // namespace

// scratch slots (of the bank): 40 = P (one u32 per output byte), 41 = chunk tables, 42 = chunk entries + round counters, 43 = segment tables
int lz77_pdec_launch(b200_ctx* ctx, int variant, const uint8_t* d_stream, const uint64_t* d_block_off, const uint64_t* d_block_sizes,
                     uint64_t n, uint64_t bs, uint64_t nblocks, uint8_t* d_out) {
    if (bs >= (1ull << 30)) { B200_SET_ERR("lz77 decode: blocks of 1 GiB and more are not supported by the token-parallel decoder"); return B200_ERR_ARG; }
    const uint64_t cpb64 = variant ? (bs + 1 + CU - 1) / CU + 1 : (9 * bs + 19 + CB - 1) / CB + 1;   // deflate: <= 2 bs + 2 token bytes
    const uint32_t cpb = (uint32_t)cpb64;
    const uint64_t nchunks = nblocks * cpb;
    uint32_t* P; uint32_t* tables; uint64_t* centry;
    B200_TRY(b200_scratch(ctx, B200_SLOT(ctx, 40), (size_t)n * 4 + 64, reinterpret_cast<void**>(&P)));
    B200_TRY(b200_scratch(ctx, B200_SLOT(ctx, 41), (size_t)nchunks * TAB_WORDS * 4 + 64, reinterpret_cast<void**>(&tables)));
    B200_TRY(b200_scratch(ctx, B200_SLOT(ctx, 42), (size_t)nchunks * 8 + 256, reinterpret_cast<void**>(&centry)));
    uint32_t* cnt = reinterpret_cast<uint32_t*>(centry + nchunks);          // [32] round counters
    CUDA_TRY(cudaMemsetAsync(cnt, 0, 128, ctx->stream));
    CUDA_TRY(cudaMemsetAsync(P, 0x80, (size_t)n * 4, ctx->stream));         // bytes no token writes (truncated stream) count as final
    CUDA_TRY(cudaMemsetAsync(tables, 0x1F, (size_t)nchunks * TAB_WORDS * 4, ctx->stream));   // END entries for chunks behind the stream's end
    const unsigned gc = (unsigned)((nchunks + 127) / 128);
    if (variant == 0) pdec_tables_v0_kernel<<<gc, 128, 0, ctx->stream>>>(d_stream, d_block_off, d_block_sizes, nblocks, cpb, tables);
    else pdec_tables_v1_kernel<<<gc, 128, 0, ctx->stream>>>(d_stream, d_block_off, d_block_sizes, nblocks, cpb, tables);
    const uint32_t nent = variant ? 2u : NENT;
    if (cpb <= 4 * SEG) pdec_compose_kernel<<<(unsigned)((nblocks + 3) / 4), 128, 0, ctx->stream>>>(tables, nblocks, cpb, nent, centry);
    else {
        const uint32_t spb = (cpb + SEG - 1) / SEG;
        const uint64_t nseg = nblocks * spb;
        uint8_t* sbuf;
        B200_TRY(b200_scratch(ctx, B200_SLOT(ctx, 43), (size_t)nseg * (TAB_WORDS * 12 + 8) + 64, reinterpret_cast<void**>(&sbuf)));
        uint64_t* seg_bytes = reinterpret_cast<uint64_t*>(sbuf);
        uint64_t* seg_entry = seg_bytes + nseg * TAB_WORDS;
        uint32_t* seg_tables = reinterpret_cast<uint32_t*>(seg_entry + nseg);
        pdec_seg_tables_kernel<<<(unsigned)((nseg * nent + 127) / 128), 128, 0, ctx->stream>>>(tables, nblocks, cpb, spb, nent, seg_tables, seg_bytes);
        pdec_seg_walk_kernel<<<(unsigned)((nblocks + 127) / 128), 128, 0, ctx->stream>>>(seg_tables, seg_bytes, nblocks, spb, nent, seg_entry);
        pdec_seg_fill_kernel<<<(unsigned)((nseg + 127) / 128), 128, 0, ctx->stream>>>(tables, seg_entry, nblocks, cpb, spb, nent, centry);
        ctx->launches += 2;
    }
    if (variant == 0) pdec_emit_v0_kernel<<<gc, 128, 0, ctx->stream>>>(d_stream, d_block_off, d_block_sizes, n, bs, nblocks, cpb, centry, d_out, P);
    else pdec_emit_v1_kernel<<<gc, 128, 0, ctx->stream>>>(d_stream, d_block_off, d_block_sizes, n, bs, nblocks, cpb, centry, d_out, P);
    uint32_t rounds = 2;                                                     // 16^(rounds - 1) >= longest possible chain
    for (uint64_t d = 16; d < bs; d *= 16) ++rounds;
    if (rounds > 30) rounds = 30;
    uint64_t rg = ((n + 3) / 4 + 255) / 256;
    if (rg > (uint64_t)ctx->sm_count * 16) rg = (uint64_t)ctx->sm_count * 16;
    const unsigned rgrid = (unsigned)(rg ? rg : 1);
    for (uint32_t r = 0; r < rounds; ++r)
        pdec_resolve_kernel<<<rgrid, 256, 0, ctx->stream>>>(d_out, P, n, bs, r ? cnt + r - 1 : nullptr, cnt + r);
    ctx->launches += 3 + rounds;
    CUDA_TRY(cudaGetLastError());
    return B200_OK;
}
