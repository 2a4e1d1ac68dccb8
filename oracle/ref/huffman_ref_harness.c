/*
 * TEST INFRASTRUCTURE ONLY (oracle/_ref/libhuffman_ref.so).
 *
 * Harness around the UNMODIFIED /root/reference/algorithms/huffman/huffman.c.
 * SURVEY.md §4.2 rules applied:
 *   U4  the compressed stream is captured as word_idx + (bit_idx>0) whole u32
 *       words (huffman.c:318-325 shrinks the buffer below the last word);
 *       orc_realloc shrinks in place so those bytes stay intact;
 *   U5  the decoder is given a zero-padded, word-rounded buffer and an output
 *       buffer with slack; the symbol count it reports is returned as is;
 *   U6  exit(1) on a single-symbol input comes back as a non-zero return code;
 *   U11 arena allocations.
 */
#include <stdio.h>
#include <omp.h>
#include "arena.h"
#include "huffman.h" /* -I/root/reference/algorithms/huffman */

/* codes/lens exactly as gather_codes (huffman.c:217-250) leaves them. */
int orc_ref_huffman_tables(const uint8_t* in, uint64_t n, uint32_t* codes, uint8_t* lens) {
    orc_arena_reserve(1u << 20);
    memset(codes, 0, 256 * sizeof(uint32_t));
    memset(lens, 0, 256);
    orc_exit_armed = 1;
    int rc = setjmp(orc_exit_jmp);
    if (rc == 0) {
        Node* root = NULL;
        build_huffman_tree((char*)in, n, &root);
        gather_codes(root, 0, 0, codes, lens);
    }
    orc_exit_armed = 0;
    return rc;
}

/* Whole-buffer huffman_compress (huffman.c:288-328). out_words needs
 * n/4 + 2 words. */
int orc_ref_huffman_compress(const uint8_t* in, uint64_t n, uint32_t* out_words,
                             uint64_t* word_idx, uint64_t* bit_idx, uint64_t* buffer_size,
                             uint32_t* codes, uint8_t* lens) {
    orc_arena_reserve((size_t)n + (1u << 20));
    orc_exit_armed = 1;
    int rc = setjmp(orc_exit_jmp);
    if (rc == 0) {
        BitWriter w;
        Node root = huffman_compress((char*)in, n, &w);
        uint64_t nwords = w.word_idx + (w.bit_idx > 0);
        memcpy(out_words, w.buffer, nwords * 4);
        *word_idx = w.word_idx;
        *bit_idx = w.bit_idx;
        *buffer_size = w.buffer_size;
        if (codes && lens) {
            memset(codes, 0, 256 * sizeof(uint32_t));
            memset(lens, 0, 256);
            gather_codes(&root, 0, 0, codes, lens);
        }
    }
    orc_exit_armed = 0;
    return rc;
}

static Node* tree_from_table(const uint32_t* codes, const uint8_t* lens) {
    Node* root = init_node(0, 0);
    for (int s = 0; s < 256; ++s) {
        if (!lens[s]) continue;
        Node* cur = root;
        for (int b = lens[s] - 1; b >= 0; --b) {
            int bit = (codes[s] >> b) & 1;
            Node** next = bit ? &cur->right : &cur->left;
            if (!*next) *next = init_node(0, 0);
            cur = *next;
        }
        cur->value = (uint8_t)s;
    }
    return root;
}

/* Reference tree-walk decoder (huffman.c:330-364) on an arbitrary stream with
 * the tree rebuilt from a (codes, lens) table. out must hold out_cap bytes,
 * out_cap >= expected symbols + 64 (U5 over-run). */
int orc_ref_huffman_decompress(const uint32_t* words, uint64_t nwords, uint64_t buffer_size,
                               const uint32_t* codes, const uint8_t* lens,
                               uint8_t* out, uint64_t out_cap, uint64_t* out_size) {
    orc_arena_reserve((size_t)nwords * 4 + 4096 + (1u << 20));
    BitWriter w;
    w.buffer = (uint32_t*)orc_malloc((nwords + 4) * 4);
    memset(w.buffer, 0, (nwords + 4) * 4);
    memcpy(w.buffer, words, nwords * 4);
    w.word_idx = 0;
    w.bit_idx = 0;
    w.buffer_size = buffer_size;
    orc_exit_armed = 1;
    int rc = setjmp(orc_exit_jmp);
    if (rc == 0) {
        Node* root = tree_from_table(codes, lens);
        uint64_t sz = out_cap;
        huffman_decompress(&w, root, (char*)out, &sz);
        *out_size = sz;
    }
    orc_exit_armed = 0;
    return rc;
}

/* Per-block oracle: one huffman_compress per block. words is nblocks*stride_words. */
int orc_ref_huffman_compress_blocks(const uint8_t* in, uint64_t n, uint64_t block,
                                    uint32_t* words, uint64_t stride_words,
                                    uint64_t* word_idx, uint64_t* bit_idx,
                                    uint32_t* codes /*nblocks*256*/, uint8_t* lens /*nblocks*256*/,
                                    int threads) {
    uint64_t nblocks = (n + block - 1) / block;
    int bad = 0;
    if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
    for (int64_t b = 0; b < (int64_t)nblocks; ++b) {
        uint64_t off = (uint64_t)b * block;
        uint64_t len = n - off < block ? n - off : block;
        uint64_t bs;
        int rc = orc_ref_huffman_compress(in + off, len, words + (uint64_t)b * stride_words,
                                          &word_idx[b], &bit_idx[b], &bs,
                                          codes + (uint64_t)b * 256, lens + (uint64_t)b * 256);
        if (rc) {
#pragma omp atomic write
            bad = rc;
        }
    }
    return bad;
}

/* Timed legs for bench.py's cpu_baseline / --impl reference. Returns seconds of
 * wall time for compress and decompress of the whole buffer (the reference as
 * written: one tree, one thread). */
int orc_ref_huffman_time(const uint8_t* in, uint64_t n, double* t_comp, double* t_decomp,
                         uint64_t* comp_bytes, uint64_t* mismatches) {
    orc_arena_reserve((size_t)n * 2 + (1u << 20));
    orc_exit_armed = 1;
    int rc = setjmp(orc_exit_jmp);
    if (rc == 0) {
        BitWriter w;
        double t0 = omp_get_wtime();
        Node root = huffman_compress((char*)in, n, &w);
        double t1 = omp_get_wtime();
        uint64_t sz = n + 64;
        char* dec = (char*)orc_malloc(sz);
        double t2 = omp_get_wtime();
        huffman_decompress(&w, &root, dec, &sz);
        double t3 = omp_get_wtime();
        uint64_t bad = 0;
        for (uint64_t i = 0; i < n; ++i) bad += dec[i] != (char)in[i];
        *t_comp = t1 - t0;
        *t_decomp = t3 - t2;
        *comp_bytes = w.buffer_size;
        *mismatches = bad;
    }
    orc_exit_armed = 0;
    return rc;
}

int orc_ref_huffman_threads(void) { return omp_get_max_threads(); }
