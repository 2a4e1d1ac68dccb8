/*
 * TEST INFRASTRUCTURE ONLY (oracle/_build/liboracle_port.so).
 *
 * CPU restatement of the chunked Huffman file format of
 * /root/reference/algorithms/huffman/zig_huffman/src/main.zig. PARITY UNPINNED: there is no Zig toolchain in
 * this image, so nothing here was ever compared with the reference program's output; it is the written spec the
 * CUDA path (b200_zig_huffman_*_host) has to match byte for byte.
 *
 * Follows:
 *   BUFFER_SIZE 4 MiB chunks                         main.zig:5, 545-552
 *   histogram over the WHOLE read buffer             main.zig:100-121 (stale bytes behind a short read count too)
 *   tree by std.PriorityQueue(lessThan on freq)      main.zig:123-153; heap order restated from Zig's std (add = sift
 *                                                    up while strictly less than the parent; remove = last to the root,
 *                                                    sift down: lesser child = right only if strictly less than left,
 *                                                    stop only when the moved element is strictly less than that child)
 *   pre-order tree dump, -1 for a missing child      main.zig:155-176
 *   left-aligned codes, MSB-first into bytes         main.zig:202-238, 304-338
 *   CompressedSize{last_block:1, value:31} + bytes   main.zig:11-18, 513-530 (whole bytes only: the tail bits are lost)
 *   decoder                                          main.zig:401-447
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "port.h"

#define ZCHUNK (1u << 22)

typedef struct { uint32_t freq; int16_t left, right; uint8_t value; } znode;

static void z_sift_up(int* h, const znode* nd, int idx) {
    const int x = h[idx];
    while (idx > 0) {
        const int p = (idx - 1) >> 1;
        if (!(nd[x].freq < nd[h[p]].freq)) break;
        h[idx] = h[p]; idx = p;
    }
    h[idx] = x;
}
static int z_remove(int* h, int* size, const znode* nd) {
    const int top = h[0];
    const int x = h[--*size];
    int idx = 0;
    for (;;) {
        int c = 2 * idx + 1;
        if (c >= *size) break;
        if (c + 1 < *size && nd[h[c + 1]].freq < nd[h[c]].freq) ++c;
        if (nd[x].freq < nd[h[c]].freq) break;
        h[idx] = h[c]; idx = c;
    }
    if (*size > 0) h[idx] = x;
    return top;
}

/* returns the root, nodes in nd[0..*nn) */
static int z_build(const uint64_t* freqs, znode* nd, int* nn) {
    int h[256], size = 0; *nn = 0;
    for (int s = 0; s < 256; ++s) if (freqs[s]) {
        nd[*nn] = (znode){(uint32_t)freqs[s], -1, -1, (uint8_t)s};
        h[size] = (*nn)++; z_sift_up(h, nd, size++);
    }
    if (size == 0) return -1;
    while (size > 1) {
        const int l = z_remove(h, &size, nd), r = z_remove(h, &size, nd);
        nd[*nn] = (znode){nd[l].freq + nd[r].freq, (int16_t)l, (int16_t)r, 0};
        h[size] = (*nn)++; z_sift_up(h, nd, size++);
    }
    return h[0];
}
static uint64_t z_put_tree(const znode* nd, int v, uint8_t* o) {
    if (v < 0) { memset(o, 0xFF, 4); return 4; }
    uint64_t w = 0;
    o[w++] = nd[v].value; memcpy(o + w, &nd[v].freq, 4); w += 4;
    w += z_put_tree(nd, nd[v].left, o + w);
    w += z_put_tree(nd, nd[v].right, o + w);
    return w;
}
static void z_codes(const znode* nd, int v, uint32_t code, uint32_t len, uint32_t* codes, uint8_t* lens) {
    if (nd[v].left < 0 && nd[v].right < 0) { codes[nd[v].value] = len ? code << (32 - len) : 0; lens[nd[v].value] = (uint8_t)len; return; }
    if (nd[v].left >= 0) z_codes(nd, nd[v].left, code << 1, len + 1, codes, lens);
    if (nd[v].right >= 0) z_codes(nd, nd[v].right, (code << 1) | 1, len + 1, codes, lens);
}

/* out capacity: n + n/4 + 8 KiB per chunk. Returns bytes written, or 0 when the reference's behaviour is undefined
 * (a chunk with one distinct symbol, a code longer than 25 bits). */
uint64_t port_zig_huffman_compress(const uint8_t* in, uint64_t n, uint8_t* out) {
    uint8_t* buf = (uint8_t*)calloc(ZCHUNK, 1);        /* the read buffer: zero pages at first, then whatever was read last */
    uint8_t* comp = (uint8_t*)malloc(ZCHUNK * 4u + 16);
    uint64_t o = 0, pos = 0;
    int done = 0;
    while (!done) {
        const uint64_t len = n - pos < ZCHUNK ? n - pos : ZCHUNK;
        memcpy(buf, in + pos, len); pos += len;
        done = len < ZCHUNK;
        uint64_t freqs[256] = {0};
        for (uint32_t i = 0; i < ZCHUNK; ++i) ++freqs[buf[i]];
        znode nd[511]; int nn;
        const int root = z_build(freqs, nd, &nn);
        if (nn < 3) { o = 0; break; }
        o += z_put_tree(nd, root, out + o);
        uint32_t codes[256] = {0}; uint8_t lens[256] = {0};
        z_codes(nd, root, 0, 0, codes, lens);
        int bad = 0;
        for (int s = 0; s < 256; ++s) if (lens[s] > 25) bad = 1;
        if (bad) { o = 0; break; }
        memset(comp, 0, ZCHUNK * 4u + 16);
        uint64_t byte_idx = 0; uint32_t bit_idx = 0;
        for (uint64_t i = 0; i < len; ++i) {
            uint32_t code = codes[buf[i]] >> bit_idx;
            comp[byte_idx] |= (uint8_t)(code >> 24); comp[byte_idx + 1] |= (uint8_t)(code >> 16);
            comp[byte_idx + 2] |= (uint8_t)(code >> 8); comp[byte_idx + 3] |= (uint8_t)code;
            bit_idx += lens[buf[i]];
            byte_idx += bit_idx / 8; bit_idx %= 8;
        }
        const uint32_t hdr = (uint32_t)(len < ZCHUNK ? 1u : 0u) | (uint32_t)(byte_idx << 1);
        memcpy(out + o, &hdr, 4); o += 4;
        memcpy(out + o, comp, byte_idx); o += byte_idx;
    }
    free(buf); free(comp);
    return o;
}

/* returns bytes produced (out capacity: 4 MiB per chunk), or UINT64_MAX on a corrupt stream */
static int z_get_tree(const uint8_t* in, uint64_t bytes, uint64_t* i, znode* nd, int* nn) {
    if (*i + 4 > bytes) return -2;
    int32_t m; memcpy(&m, in + *i, 4);
    if (m == -1) { *i += 4; return -1; }
    if (*i + 5 > bytes || *nn >= 511) return -2;
    const int v = (*nn)++;
    nd[v].value = in[*i]; memcpy(&nd[v].freq, in + *i + 1, 4); *i += 5;
    const int l = z_get_tree(in, bytes, i, nd, nn); if (l == -2) return -2;
    const int r = z_get_tree(in, bytes, i, nd, nn); if (r == -2) return -2;
    nd[v].left = (int16_t)l; nd[v].right = (int16_t)r;
    return v;
}
uint64_t port_zig_huffman_decompress(const uint8_t* in, uint64_t bytes, uint8_t* out, uint64_t out_cap) {
    uint64_t i = 0, o = 0;
    int done = 0;
    while (!done) {
        znode nd[511]; int nn = 0;
        const int root = z_get_tree(in, bytes, &i, nd, &nn);
        if (root < 0 || i + 4 > bytes) return UINT64_MAX;
        uint32_t hdr; memcpy(&hdr, in + i, 4); i += 4;
        done = hdr & 1u;
        const uint32_t size = hdr >> 1;
        if (i + size > bytes) return UINT64_MAX;
        const uint8_t* p = in + i;
        uint64_t byte_idx = 0; uint32_t bit_idx = 0;
        while (byte_idx < size) {
            int v = root;
            while (nd[v].left >= 0 && nd[v].right >= 0) {
                const uint32_t cur = byte_idx < size ? p[byte_idx] : 0u;   /* past the payload: zero bits */
                v = (cur >> (7 - bit_idx)) & 1u ? nd[v].right : nd[v].left;
                ++bit_idx; byte_idx += bit_idx / 8; bit_idx %= 8;
            }
            if (o < out_cap) out[o] = nd[v].value;
            ++o;
            if (v == root) break;
        }
        i += size;
    }
    return o;
}
