/*
 * TEST INFRASTRUCTURE ONLY (oracle/_build/liboracle_port.so).
 *
 * CPU restatement of the reference's two hash-table LZ77 codecs. It is the
 * checker for the CUDA path and never part of the product. Parity is PINNED: it
 * is compared bit-for-bit against the compiled reference (oracle/_ref) in
 * tests/test_oracle.py and against the golden vectors in tests/golden/.
 *
 * Follows:
 *   hash                 /root/reference/algorithms/lz77/lz77.c:13-41
 *                        (identical copy: algorithms/deflate/lz77.c:14-42)
 *   table insert / find  algorithms/lz77/lz77.c:55-108, deflate/lz77.c:77-174
 *   greedy parse         algorithms/lz77/lz77.c:264-345, deflate/lz77.c:199-280
 *   token formats        algorithms/lz77/lz77.c:139-184,293-294,328-330 (LSB-first bits)
 *                        algorithms/deflate/lz77.c:176-197 (byte tokens)
 *   decoder              algorithms/lz77/lz77.c:347-377
 *
 * The FIFO ring of the reference is replaced by the equivalent "lazy expiry" rule
 * (SURVEY.md §7.4): the entry placed for position i is live at time P iff
 * i >= P - W, plus the slot-0 exception that reproduces the reference's early
 * is_full flip (lz77.c:70-85): after insert number i, slot 0 is cleared whenever
 * i == W-1 or position i-W had been placed in slot 0.
 *
 * Out-of-table probing (U9) is undefined in the reference; here the table has a
 * guard tail of GUARD slots that behaves like more table (standalone insert and
 * both finds do not wrap, lz77.c:61,102, deflate/lz77.c:168; the deflate insert
 * wraps, deflate/lz77.c:99-101).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <omp.h>
#include "port.h"

#define TABLE_SLOTS (1u << 20)
#define GUARD 65536u
#define NONE 0xFFFFFFFFu

uint32_t port_lz77_hash(uint32_t k) {
    k *= 0xcc9e2d51u;
    k = (k << 15) | (k >> 17);
    k *= 0x1b873593u;
    uint32_t h = k; /* seed 0 */
    h = ((h << 13) | (h >> 19)) * 5u + 0xe6546b64u;
    h ^= h >> 16;
    h *= 0x85ebca6bu;
    h ^= h >> 13;
    h *= 0xc2b2ae35u;
    h ^= h >> 16;
    return h & (TABLE_SLOTS - 1);
}

typedef struct {
    uint32_t* pat;   /* pattern stored in the slot */
    uint32_t* idx;   /* position that placed it, NONE = never used / cleared */
    uint32_t* touched;
    uint32_t  ntouched;
    uint32_t  cap_touched;
} table_t;

static void table_alloc(table_t* t, uint64_t max_n) {
    t->pat = (uint32_t*)malloc(sizeof(uint32_t) * (TABLE_SLOTS + GUARD));
    t->idx = (uint32_t*)malloc(sizeof(uint32_t) * (TABLE_SLOTS + GUARD));
    memset(t->idx, 0xFF, sizeof(uint32_t) * (TABLE_SLOTS + GUARD));
    t->cap_touched = (uint32_t)max_n + 64;
    t->touched = (uint32_t*)malloc(sizeof(uint32_t) * t->cap_touched);
    t->ntouched = 0;
}
static void table_scrub(table_t* t) {
    for (uint32_t i = 0; i < t->ntouched; ++i) t->idx[t->touched[i]] = NONE;
    t->ntouched = 0;
}
static void table_free(table_t* t) { free(t->pat); free(t->idx); free(t->touched); }

static inline uint32_t word_at(const uint8_t* d, uint64_t n, uint64_t p) {
    /* little-endian 4-byte load, bytes past the block read as 0 (U1) */
    uint32_t w = 0;
    for (int k = 0; k < 4; ++k) if (p + k < n) w |= (uint32_t)d[p + k] << (8 * k);
    return w;
}
static inline uint8_t byte_at(const uint8_t* d, uint64_t n, uint64_t p) { return p < n ? d[p] : 0; }

/*
 * Core: greedy parse over one block. variant 0 = algorithms/lz77 (W=2^14,
 * MAX_LEN 15, reject distance == W), variant 1 = algorithms/deflate (W=2^15,
 * MAX_LEN 31, reject distance >= W-1, wrapping insert).
 * Emits a token list: tok[k] = literal byte (len 0) or (offset,len).
 * Optionally records F[p] (find result at p if p is a token start, else NONE-1
 * "not evaluated") for kernel debugging.
 */
typedef struct { uint32_t off; uint32_t len; uint8_t lit; } token_t;

static uint64_t parse_block(const uint8_t* d, uint64_t n, int variant, table_t* t,
                            token_t* tok, uint32_t* F) {
    const uint32_t W = variant ? (1u << 15) : (1u << 14);
    const uint32_t MAX_LEN = variant ? 31u : 15u;
    uint64_t ntok = 0;
    uint64_t p = 0;
    /* slot-0 exception bookkeeping: placed0[i & (W-1)] != 0 iff the position that
     * was inserted W steps ago went to slot 0 (ring of one window). */
    uint8_t* placed0 = (uint8_t*)calloc(W, 1);
    uint64_t inserted = 0; /* number of inserts done == next position to insert */

#define LIVE(s) (t->idx[s] != NONE && (uint64_t)t->idx[s] + W >= inserted)
#define DO_INSERT(pos) do { \
        uint32_t pat_ = word_at(d, n, (pos)); \
        uint32_t s_ = port_lz77_hash(pat_); \
        while (LIVE(s_)) { ++s_; if (variant && s_ == TABLE_SLOTS) s_ = 0; } \
        t->pat[s_] = pat_; t->idx[s_] = (uint32_t)(pos); \
        if (t->ntouched < t->cap_touched) t->touched[t->ntouched++] = s_; \
        uint32_t ring_ = (uint32_t)(pos) & (W - 1); \
        int clear0_ = ((pos) == W - 1) || ((pos) >= W && placed0[ring_]); \
        placed0[ring_] = (s_ == 0); \
        ++inserted; \
        if (clear0_) t->idx[0] = NONE; \
    } while (0)

    while (p < n) {
        uint32_t pat = word_at(d, n, p);
        uint32_t s = port_lz77_hash(pat);
        while (LIVE(s) && t->pat[s] != pat) ++s;
        uint32_t m = LIVE(s) ? t->idx[s] : NONE;
        if (F) F[p] = m;
        int reject = (m == NONE);
        if (!reject) {
            uint64_t dist = p - m;
            reject = variant ? (dist >= W - 1) : (dist == W);
        }
        if (reject) {
            tok[ntok].len = 0; tok[ntok].off = 0; tok[ntok].lit = d[p]; ++ntok;
            DO_INSERT(p);
            ++p;
        } else {
            uint64_t mi = (uint64_t)m + 4, bi = p + 4;
            while (byte_at(d, n, mi) == byte_at(d, n, bi) && mi - m < MAX_LEN) { ++mi; ++bi; }
            uint32_t len = (uint32_t)(mi - m);
            tok[ntok].len = len; tok[ntok].off = (uint32_t)(bi - mi); tok[ntok].lit = 0; ++ntok;
            for (uint32_t k = 0; k < len; ++k) DO_INSERT(p + k);
            p += len;
        }
    }
#undef LIVE
#undef DO_INSERT
    free(placed0);
    return ntok;
}

/* LSB-first bit packer of algorithms/lz77/lz77.c:144-174 (values LSB first). */
static inline void put_bits(uint8_t* out, uint64_t* bitpos, uint32_t value, uint32_t nbits) {
    for (uint32_t b = 0; b < nbits; ++b) {
        if ((value >> b) & 1u) out[*bitpos >> 3] |= (uint8_t)(1u << (*bitpos & 7));
        ++*bitpos;
    }
}

static uint64_t emit_bits(const token_t* tok, uint64_t ntok, uint8_t* out) {
    uint64_t bp = 0;
    for (uint64_t k = 0; k < ntok; ++k) {
        if (tok[k].len == 0) { put_bits(out, &bp, 0, 1); put_bits(out, &bp, tok[k].lit, 8); }
        else { put_bits(out, &bp, 1, 1); put_bits(out, &bp, tok[k].off, 14); put_bits(out, &bp, tok[k].len, 4); }
    }
    return bp;
}

static uint64_t emit_bytes(const token_t* tok, uint64_t ntok, uint8_t* out) {
    uint64_t o = 0;
    for (uint64_t k = 0; k < ntok; ++k) {
        if (tok[k].len == 0) { out[o++] = 0; out[o++] = tok[k].lit; }
        else { out[o++] = 1; out[o++] = (uint8_t)(tok[k].off & 0xFF); out[o++] = (uint8_t)(tok[k].off >> 8); out[o++] = (uint8_t)tok[k].len; }
    }
    return o;
}

/* out: zero-initialised by this function, capacity 2*n+16. */
int port_lz77_compress(const uint8_t* in, uint64_t n, uint8_t* out, uint64_t* bit_index, uint32_t* F_or_null) {
    table_t t; table_alloc(&t, n);
    token_t* tok = (token_t*)malloc(sizeof(token_t) * (n + 1));
    uint64_t ntok = parse_block(in, n, 0, &t, tok, F_or_null);
    memset(out, 0, 2 * n + 16);
    *bit_index = emit_bits(tok, ntok, out);
    free(tok); table_free(&t);
    return 0;
}

int port_deflate_lz77_compress(const uint8_t* in, uint64_t n, uint8_t* out, uint64_t* out_n, uint32_t* F_or_null) {
    table_t t; table_alloc(&t, n);
    token_t* tok = (token_t*)malloc(sizeof(token_t) * (n + 1));
    uint64_t ntok = parse_block(in, n, 1, &t, tok, F_or_null);
    *out_n = emit_bytes(tok, ntok, out);
    free(tok); table_free(&t);
    return 0;
}

/* Block-segmented, OpenMP over blocks. variant as above. sizes[] = bit count
 * (variant 0) or byte count (variant 1). out_stride >= 2*block+16. */
int port_lz77_compress_blocks(const uint8_t* in, uint64_t n, uint64_t block, int variant,
                              uint8_t* out, uint64_t out_stride, uint64_t* sizes, int threads) {
    uint64_t nblocks = (n + block - 1) / block;
    if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel num_threads(threads)
    {
        table_t t; table_alloc(&t, block);
        token_t* tok = (token_t*)malloc(sizeof(token_t) * (block + 1));
#pragma omp for schedule(dynamic, 4)
        for (int64_t b = 0; b < (int64_t)nblocks; ++b) {
            uint64_t off = (uint64_t)b * block;
            uint64_t len = n - off < block ? n - off : block;
            uint64_t ntok = parse_block(in + off, len, variant, &t, tok, NULL);
            uint8_t* o = out + (uint64_t)b * out_stride;
            if (variant == 0) { memset(o, 0, 2 * len + 16); sizes[b] = emit_bits(tok, ntok, o); }
            else sizes[b] = emit_bytes(tok, ntok, o);
            table_scrub(&t);
        }
        free(tok); table_free(&t);
    }
    return 0;
}

/* Decoder for the LSB-first bit tokens, algorithms/lz77/lz77.c:347-377. out needs
 * size + 16 bytes (a final match may overshoot `size`). Returns bytes produced. */
static inline uint32_t get_bits(const uint8_t* s, uint64_t* bp, uint32_t nbits) {
    uint32_t v = 0;
    for (uint32_t b = 0; b < nbits; ++b) { v |= (uint32_t)((s[*bp >> 3] >> (*bp & 7)) & 1u) << b; ++*bp; }
    return v;
}
uint64_t port_lz77_decompress(const uint8_t* stream, uint64_t size, uint8_t* out) {
    uint64_t bp = 0, o = 0;
    while (o < size) {
        if (get_bits(stream, &bp, 1)) {
            uint32_t off = get_bits(stream, &bp, 14), len = get_bits(stream, &bp, 4);
            for (uint32_t k = 0; k < len; ++k) out[o + k] = out[o - off + k];
            o += len;
        } else {
            out[o++] = (uint8_t)get_bits(stream, &bp, 8);
        }
    }
    return o;
}

/* Decoder for the byte-token format of algorithms/deflate/lz77.c:176-197. The
 * reference's own lz77_decompress (deflate/lz77.c:282-311) is broken and
 * decompress() is empty (deflate/deflate.c:78-79), so this is pinned only by the
 * round trip: decode(reference tokens) == input. Returns bytes produced. */
uint64_t port_deflate_lz77_decompress(const uint8_t* tok, uint64_t ntokbytes, uint8_t* out) {
    uint64_t i = 0, o = 0;
    while (i < ntokbytes) {
        if (tok[i] == 0) { out[o++] = tok[i + 1]; i += 2; }
        else {
            uint32_t off = (uint32_t)tok[i + 1] | ((uint32_t)tok[i + 2] << 8), len = tok[i + 3];
            for (uint32_t k = 0; k < len; ++k) out[o + k] = out[o - off + k];
            o += len; i += 4;
        }
    }
    return o;
}

/* Block-parallel decode (OpenMP over blocks) for the timed CPU baseline: stream is
 * the concatenation of the per-block token streams, off[b] their byte offsets
 * (off[nblocks] = total). variant as in port_lz77_compress_blocks. Each block
 * decodes into out + b*block; scratch of block+64 bytes per thread absorbs the
 * overshoot of a final match (U1). Returns the number of mismatching block sizes. */
uint64_t port_lz77_decompress_blocks(const uint8_t* stream, const uint64_t* off, uint64_t nblocks, uint64_t block,
                                     uint64_t n, int variant, uint8_t* out, int threads) {
    uint64_t bad = 0;
    if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel num_threads(threads) reduction(+ : bad)
    {
        uint8_t* tmp = (uint8_t*)malloc(block + 64);
#pragma omp for schedule(dynamic, 8)
        for (int64_t b = 0; b < (int64_t)nblocks; ++b) {
            uint64_t o0 = (uint64_t)b * block;
            uint64_t len = n - o0 < block ? n - o0 : block;
            uint64_t got = variant ? port_deflate_lz77_decompress(stream + off[b], off[b + 1] - off[b], tmp)
                                   : port_lz77_decompress(stream + off[b], len, tmp);
            if (got < len) ++bad;
            memcpy(out + o0, tmp, len);
        }
        free(tmp);
    }
    return bad;
}
