/*
 * TEST INFRASTRUCTURE ONLY (oracle/_ref/liblz77_ref.so).
 *
 * Thin harness around the UNMODIFIED reference sources
 *   /root/reference/algorithms/lz77/lz77.c  (+ lz77.h)
 * compiled in place by oracle/Makefile. It neutralises the undefined behaviour
 * listed in SURVEY.md §4.2 so the reference becomes a deterministic oracle:
 *   U1  input is copied into a buffer followed by ORC_PAD zero bytes
 *       (lz77.c:285,305 read past `size`);
 *   U3  only `bit_index` bits are reported, the tail of the last byte is masked
 *       (lz77.c:276 mallocs the stream without clearing it);
 *   U11 allocations go to a per-thread arena (arena.h).
 */
#include <stdio.h>
#include <omp.h>
#include "arena.h"
#include "lz77.h" /* resolved with -I/root/reference/algorithms/lz77 */

#define ORC_PAD 64

uint32_t orc_ref_lz77_hash(uint32_t pattern) { return hash(pattern); }

/* One reference call on one buffer. out must hold 2*n + 16 bytes.
 * Returns 0, or the reference's exit code if it called exit(). */
int orc_ref_lz77_compress(const uint8_t* in, uint64_t n, uint8_t* out, uint64_t* bit_index) {
    orc_arena_reserve((size_t)n * 3 + ORC_PAD + sizeof(ArrayNode) * (size_t)TABLE_SIZE + (1u << 20));
    char* padded = (char*)orc_malloc(n + ORC_PAD);
    memcpy(padded, in, n);
    memset(padded + n, 0, ORC_PAD);
    orc_exit_armed = 1;
    int rc = setjmp(orc_exit_jmp);
    if (rc == 0) {
        BitStream* s = lz77_compress(padded, n);
        uint64_t bits = s->bit_index;
        uint64_t full = bits / 8;
        memcpy(out, s->data, full);
        if (bits % 8) out[full] = s->data[full] & (uint8_t)((1u << (bits % 8)) - 1);
        *bit_index = bits;
    }
    orc_exit_armed = 0;
    return rc;
}

/* Reference decoder on an arbitrary stream (e.g. one produced by the GPU path).
 * out must hold size + ORC_PAD bytes: the last match may overshoot `size`
 * (lz77.c:358-368). */
int orc_ref_lz77_decompress(const uint8_t* stream, uint64_t bit_index, uint64_t size,
                            uint8_t* out, uint64_t* out_size) {
    uint64_t nbytes = bit_index / 8 + 1;
    orc_arena_reserve((size_t)nbytes + size + 4 * ORC_PAD + 4096);
    BitStream bs;
    bs.data = (uint8_t*)orc_malloc(nbytes + ORC_PAD);
    memset(bs.data, 0, nbytes + ORC_PAD);
    memcpy(bs.data, stream, (bit_index + 7) / 8);
    bs.bit_index = bit_index;
    /* the reference mallocs exactly `size`; a pad allocation right behind it in the
     * arena absorbs the overshoot of the final match */
    orc_exit_armed = 1;
    int rc = setjmp(orc_exit_jmp);
    if (rc == 0) {
        uint64_t dsz = 0;
        size_t mark = orc_arena_off;
        char* dec = lz77_decompress(&bs, size, &dsz);
        (void)mark;
        memcpy(out, dec, dsz < size + ORC_PAD ? dsz : size + ORC_PAD);
        *out_size = dsz;
    }
    orc_exit_armed = 0;
    return rc;
}

/* Block-segmented oracle (SURVEY.md §8a "Parity contract for L-rows"): one fresh
 * reference call per block. out is nblocks * out_stride bytes (out_stride >=
 * 2*block+16); bit_index has nblocks entries. threads<=0 -> all cores. */
int orc_ref_lz77_compress_blocks(const uint8_t* in, uint64_t n, uint64_t block,
                                 uint8_t* out, uint64_t out_stride, uint64_t* bit_index,
                                 int threads) {
    uint64_t nblocks = (n + block - 1) / block;
    int bad = 0;
    if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
    for (int64_t b = 0; b < (int64_t)nblocks; ++b) {
        uint64_t off = (uint64_t)b * block;
        uint64_t len = n - off < block ? n - off : block;
        int rc = orc_ref_lz77_compress(in + off, len, out + (uint64_t)b * out_stride, &bit_index[b]);
        if (rc) {
#pragma omp atomic write
            bad = rc;
        }
    }
    return bad;
}

int orc_ref_lz77_threads(void) { return omp_get_max_threads(); }
