/*
 * TEST INFRASTRUCTURE ONLY (oracle/_build/liboracle_port.so).
 *
 * Entropy stage of the deflate token stream: the step the reference leaves as
 * "TODO: Build huffman tree and encode compressed buffer"
 * (/root/reference/algorithms/deflate/lz77.c:279). What the reference does define is
 * restated exactly; what it does not define is chosen here and this file is its written spec.
 *
 * PINNED (tests/test_oracle.py checks them against the compiled reference):
 *   frequencies[286]   lz77.c:206,231,273 with append_huffman_tree_literal / _pair
 *                      (deflate/huffman.c:49-62): a literal token counts its byte, a match token
 *                      counts symbol 256 + clz16(offset)  (a distance bit-width class; lengths
 *                      are not counted)
 *   bit packing        write_bits, deflate/huffman.c:18-48 = algorithms/huffman/huffman.c:18-48:
 *                      MSB-first into host-endian u32 words
 * CHOSEN HERE (parity unpinned: the reference has declarations only, deflate/huffman.h:16-32,84-92):
 *   code construction  the min-heap rule of algorithms/huffman/huffman.c:100-211 over the 286
 *                      symbols in symbol order (compare_nodes, deflate/huffman.h:18-20, is the same
 *                      strict order), codes as gather_codes assigns them (deflate/huffman.c:64-98:
 *                      left appends 0, right appends 1), kept in 32 bits instead of the u16 of the
 *                      scaffolding; a block with ONE distinct symbol gets the 1-bit code 0
 *   token layout       literal: code[byte]
 *                      match  : code[256 + k], k = clz16(offset); then the 15-k bits of the offset
 *                               below its leading one; then the length in MAX_LENGTH_BITS = 5 bits
 *                               (deflate/lz77.h:7)
 *   an offset of 0 (never produced by lz77_compress) is class k = 16 with no extra bits.
 */
#include <stdint.h>
#include <string.h>
#include "port.h"

static int clz16(uint32_t x) { return x ? __builtin_clz(x) - 16 : 16; }

/* tok: the byte tokens of ONE block (write_literal / write_length_distance, lz77.c:176-197) */
void port_dfl_frequencies(const uint8_t* tok, uint64_t nbytes, uint64_t* freq) {
    memset(freq, 0, PORT_DFL_NSYM * sizeof(uint64_t));
    uint64_t i = 0;
    while (i + 1 < nbytes) {
        if (tok[i] == 1 && i + 3 < nbytes) {
            const uint32_t off = tok[i + 1] | ((uint32_t)tok[i + 2] << 8);
            ++freq[256 + clz16(off)];
            i += 4;
        } else {
            ++freq[tok[i + 1]];
            i += 2;
        }
    }
}

/* returns the number of distinct symbols */
int port_dfl_build(const uint64_t* freq, uint32_t* codes, uint8_t* lens) {
    const int distinct = port_huffman_build_n(freq, PORT_DFL_NSYM, codes, lens, 0, 0);
    if (distinct == 1)
        for (int s = 0; s < PORT_DFL_NSYM; ++s) if (freq[s]) { codes[s] = 0; lens[s] = 1; }
    return distinct;
}

static void put_bits(uint32_t* words, uint64_t* bitpos, uint32_t v, uint32_t len) {
    if (!len) return;
    const uint64_t w = *bitpos >> 5;
    const uint32_t used = (uint32_t)(*bitpos & 31), room = 32 - used;
    if (len < 32) v &= (1u << len) - 1;
    if (len <= room) words[w] |= v << (room - len);
    else { words[w] |= v >> (len - room); words[w + 1] |= v << (32 - (len - room)); }
    *bitpos += len;
}

/* words must be zeroed; returns the bits written */
uint64_t port_dfl_encode(const uint8_t* tok, uint64_t nbytes, const uint32_t* codes, const uint8_t* lens, uint32_t* words) {
    uint64_t bitpos = 0, i = 0;
    while (i + 1 < nbytes) {
        if (tok[i] == 1 && i + 3 < nbytes) {
            const uint32_t off = tok[i + 1] | ((uint32_t)tok[i + 2] << 8);
            const int k = clz16(off);
            put_bits(words, &bitpos, codes[256 + k], lens[256 + k]);
            if (k < 15) put_bits(words, &bitpos, off, 15 - k);
            put_bits(words, &bitpos, tok[i + 3], 5);
            i += 4;
        } else {
            put_bits(words, &bitpos, codes[tok[i + 1]], lens[tok[i + 1]]);
            i += 2;
        }
    }
    return bitpos;
}

static uint32_t get_bits(const uint32_t* words, uint64_t nwords, uint64_t* bitpos, uint32_t len) {
    uint32_t v = 0;
    for (uint32_t q = 0; q < len; ++q) {
        const uint64_t w = *bitpos >> 5;
        const uint32_t bit = w < nwords ? (words[w] >> (31 - (*bitpos & 31))) & 1 : 0;
        v = (v << 1) | bit;
        ++*bitpos;
    }
    return v;
}

/* decodes tokens until nbytes token bytes are produced; returns 0, or 1 for a corrupt stream */
int port_dfl_decode(const uint32_t* words, uint64_t nwords, const uint32_t* codes, const uint8_t* lens,
                    uint64_t nbytes, uint8_t* tok_out, uint64_t* bits_used) {
    int left[2 * PORT_HUFF_MAX_SYMS], right[2 * PORT_HUFF_MAX_SYMS], sym[2 * PORT_HUFF_MAX_SYMS], nn = 1;
    left[0] = right[0] = -1; sym[0] = -1;
    for (int s = 0; s < PORT_DFL_NSYM; ++s) {
        if (!lens[s]) continue;
        int cur = 0;
        for (int b = lens[s] - 1; b >= 0; --b) {
            int* nx = ((codes[s] >> b) & 1) ? &right[cur] : &left[cur];
            if (*nx < 0) { *nx = nn; left[nn] = right[nn] = -1; sym[nn] = -1; ++nn; }
            cur = *nx;
        }
        sym[cur] = s;
    }
    uint64_t bitpos = 0, o = 0;
    while (o + 1 < nbytes) {
        int v = 0;
        while (sym[v] < 0) {
            const int bit = (int)get_bits(words, nwords, &bitpos, 1);
            v = bit ? right[v] : left[v];
            if (v < 0) return 1;
        }
        const int s = sym[v];
        if (s < 256) { tok_out[o++] = 0; tok_out[o++] = (uint8_t)s; }
        else {
            const int k = s - 256;
            uint32_t off = 0;
            if (k < 15) off = (1u << (15 - k)) | get_bits(words, nwords, &bitpos, 15 - k);
            else if (k == 15) off = 1;
            const uint32_t len = get_bits(words, nwords, &bitpos, 5);
            if (o + 3 >= nbytes) return 1;
            tok_out[o++] = 1; tok_out[o++] = (uint8_t)(off & 0xFF); tok_out[o++] = (uint8_t)(off >> 8); tok_out[o++] = (uint8_t)len;
        }
    }
    if (bits_used) *bits_used = bitpos;
    return 0;
}
