/*
 * TEST INFRASTRUCTURE ONLY (oracle/_ref/libdeflate_ref.so).
 *
 * Harness around the UNMODIFIED /root/reference/algorithms/deflate/{lz77,huffman,
 * deflate}.c. SURVEY.md §4.2 rules applied:
 *   U1  padded input copy (deflate/lz77.c:219,241 read past the block);
 *   U2  is_set[] and indices[] are zeroed after init_hash_table
 *       (deflate/lz77.c:47-49 leaves them uninitialised);
 *   U11 arena allocations.
 * The parity contract is lz77_compress (deflate/lz77.c:199)
 * called once per block on a FRESH table.
 */
#include <stdio.h>
#include <stdbool.h>
#include <omp.h>
#include "arena.h"
#include "lz77.h"    /* -I/root/reference/algorithms/deflate */
#include "huffman.h"

#define ORC_PAD 64

uint32_t orc_ref_deflate_hash(uint32_t pattern) { return hash(pattern); }

static void fresh_table(HashTableArray* t) {
    init_hash_table(t);
    memset(t->buckets.indices, 0, sizeof(uint64_t) * (size_t)TABLE_SIZE);
    memset(t->buckets.is_set, 0, sizeof(bool) * (size_t)TABLE_SIZE);
}

/* Bring a used table back to the freshly-initialised state without touching all
 * 13 MiB: the only slots that can still be set are the ones recorded in the FIFO
 * ring (every older one was cleared by insert_hash_table itself,
 * deflate/lz77.c:124-136). Used by the timed multi-block baseline only; the
 * parity entry point below always builds a fresh table. */
static void scrub_table(HashTableArray* t) {
    for (uint32_t i = 0; i < (uint32_t)WINDOW_SIZE; ++i) {
        uint32_t s = t->bucket_indices[i];
        t->buckets.patterns[s] = 0;
        t->buckets.indices[s] = 0;
        t->buckets.is_set[s] = false;
        t->bucket_indices[i] = 0;
    }
    t->current_idx = 0;
    t->is_full = false;
}

/* out must hold 2*n + 16 bytes. */
int orc_ref_deflate_lz77_compress(const uint8_t* in, uint64_t n, uint8_t* out, uint64_t* out_n) {
    orc_arena_reserve((size_t)n + ORC_PAD + sizeof(HashTableArray) + 14u * (size_t)TABLE_SIZE + (1u << 20));
    char* padded = (char*)orc_malloc(n + ORC_PAD);
    memcpy(padded, in, n);
    memset(padded + n, 0, ORC_PAD);
    HashTableArray* t = (HashTableArray*)orc_malloc(sizeof(HashTableArray));
    orc_exit_armed = 1;
    int rc = setjmp(orc_exit_jmp);
    if (rc == 0) {
        fresh_table(t);
        uint64_t cn = 0;
        lz77_compress(padded, n, (char*)out, &cn, t);
        *out_n = cn;
    }
    orc_exit_armed = 0;
    return rc;
}

/* Block-segmented run. persistent=0: fresh table per block (the parity contract);
 * persistent=1: one table shared by all blocks in order, as the shipped
 * compress() does (deflate/deflate.c:13-63; inherently sequential -> 1 thread).
 * out is nblocks*out_stride bytes, out_stride >= 2*block+16. */
int orc_ref_deflate_lz77_compress_blocks(const uint8_t* in, uint64_t n, uint64_t block,
                                         uint8_t* out, uint64_t out_stride, uint64_t* out_n,
                                         int persistent, int threads) {
    uint64_t nblocks = (n + block - 1) / block;
    int bad = 0;
    if (threads <= 0) threads = omp_get_max_threads();
    if (persistent) threads = 1;
#pragma omp parallel num_threads(threads)
    {
        orc_arena_reserve((size_t)block + ORC_PAD + sizeof(HashTableArray) + 14u * (size_t)TABLE_SIZE + (1u << 20));
        char* padded = (char*)orc_malloc(block + ORC_PAD);
        HashTableArray* t = (HashTableArray*)orc_malloc(sizeof(HashTableArray));
        fresh_table(t);
        orc_exit_armed = 1;
        int rc = setjmp(orc_exit_jmp);
        if (rc == 0) {
#pragma omp for schedule(dynamic, 4)
            for (int64_t b = 0; b < (int64_t)nblocks; ++b) {
                uint64_t off = (uint64_t)b * block;
                uint64_t len = n - off < block ? n - off : block;
                memcpy(padded, in + off, len);
                memset(padded + len, 0, ORC_PAD);
                uint64_t cn = 0;
                lz77_compress(padded, len, (char*)(out + (uint64_t)b * out_stride), &cn, t);
                out_n[b] = cn;
                if (!persistent) scrub_table(t);
            }
        } else {
#pragma omp atomic write
            bad = rc;
        }
        orc_exit_armed = 0;
    }
    return bad;
}

int orc_ref_deflate_threads(void) { return omp_get_max_threads(); }

/* frequencies[286] as lz77_compress counts them (deflate/lz77.c:206,231,273): the harness walks a
 * token stream and calls the reference's own append_huffman_tree_literal / _pair
 * (deflate/huffman.c:49-62) for every token. freq: 286 entries, zeroed here. */
void orc_ref_deflate_token_frequencies(const uint8_t* tok, uint64_t nbytes, uint32_t* freq) {
    memset(freq, 0, NUM_CODES * sizeof(uint32_t));
    uint64_t i = 0;
    while (i + 1 < nbytes) {
        if (tok[i] == 1 && i + 3 < nbytes) {
            append_huffman_tree_pair(freq, (uint16_t)(tok[i + 1] | (tok[i + 2] << 8)));
            i += 4;
        } else {
            append_huffman_tree_literal(freq, (char)tok[i + 1]);
            i += 2;
        }
    }
}

/* the reference's bit writer (deflate/huffman.c:9-48) fed with a list of (value, length) pairs;
 * values are pre-masked to their length as the code tables of gather_codes are. Returns the
 * number of bits; words must hold nbits/32 + 2 entries. */
uint64_t orc_ref_deflate_write_bits(const uint32_t* values, const uint8_t* lengths, uint64_t count,
                                    uint32_t* words, uint64_t words_cap) {
    orc_arena_reserve((size_t)words_cap * 4 + 4096);
    BitWriter w;
    init_bitwriter(&w, words_cap * 4);
    for (uint64_t i = 0; i < count; ++i) if (lengths[i]) write_bits(&w, values[i], lengths[i]);
    const uint64_t nw = w.word_idx + (w.bit_idx ? 1 : 0);
    memcpy(words, w.buffer, (size_t)(nw < words_cap ? nw : words_cap) * 4);
    return w.word_idx * 32 + w.bit_idx;
}
