/*
 * TEST INFRASTRUCTURE ONLY (oracle/_build/liboracle_port.so).
 *
 * CPU restatement of /root/reference/algorithms/huffman/huffman.c. Parity is
 * PINNED against the compiled reference (tests/test_oracle.py) and the
 * "nine times" fixture of algorithms/huffman/main.c:25-31.
 *
 *   histogram            huffman.c:184-187
 *   heap rules           huffman.c:100-159  (strict '<' sift-up; sift-down picks
 *                        left if left < cur, then right if right < that; dequeue
 *                        moves the last element to the root)
 *   tree build           huffman.c:189-211  (leaves enqueued in symbol order,
 *                        parent = (left = 1st dequeue, right = 2nd dequeue))
 *   code assignment      huffman.c:217-250  (left appends 0, right appends 1,
 *                        code value right-aligned, unlimited length)
 *   bit packing          huffman.c:18-48    (MSB-first into host-endian u32 words)
 *   sizes                huffman.c:318-320
 *   decoder              huffman.c:330-364  (including its termination rule, U5)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "port.h"

typedef struct { uint64_t freq; int left, right; int sym; } hnode_t;

static void sift_up(int* heap, const hnode_t* nd, int idx) {
    while (idx > 0) {
        int parent = (idx - 1) / 2;
        if (!(nd[heap[idx]].freq < nd[heap[parent]].freq)) break;
        int tmp = heap[idx]; heap[idx] = heap[parent]; heap[parent] = tmp;
        idx = parent;
    }
}
static void sift_down(int* heap, int size, const hnode_t* nd, int idx) {
    for (;;) {
        int left = 2 * idx + 1, right = 2 * idx + 2, smallest = idx;
        if (left < size && nd[heap[left]].freq < nd[heap[smallest]].freq) smallest = left;
        if (right < size && nd[heap[right]].freq < nd[heap[smallest]].freq) smallest = right;
        if (smallest == idx) return;
        int tmp = heap[idx]; heap[idx] = heap[smallest]; heap[smallest] = tmp;
        idx = smallest;
    }
}
static int heap_pop(int* heap, int* size, const hnode_t* nd) {
    int top = heap[0];
    heap[0] = heap[--*size];
    sift_down(heap, *size, nd, 0);
    return top;
}

/* freq[nsym] -> codes/lens (nsym <= PORT_HUFF_MAX_SYMS). The reference sums u32 frequencies
 * in u32 (init_node takes uint32_t, huffman.c:165-168,204); inputs < 4 GiB never wrap.
 * Returns the number of distinct symbols; 0 or 1 distinct symbols is the
 * reference's exit(1) case (U6) and yields all-zero lengths. nodes_out (optional,
 * 2*nsym-1 entries x 3 ints: left,right,sym) receives the tree, root index returned
 * in *root_out. nsym = 256 is algorithms/huffman; nsym = 286 is the alphabet of the
 * deflate scaffolding (NUM_CODES, algorithms/deflate/huffman.h:6). */
int port_huffman_build_n(const uint64_t* freq, int nsym, uint32_t* codes, uint8_t* lens, int* nodes_out, int* root_out) {
    hnode_t nd[2 * PORT_HUFF_MAX_SYMS - 1];
    int heap[PORT_HUFF_MAX_SYMS];
    int size = 0, nn = 0;
    memset(codes, 0, (size_t)nsym * sizeof(uint32_t));
    memset(lens, 0, (size_t)nsym);
    for (int s = 0; s < nsym; ++s) {
        if (!freq[s]) continue;
        nd[nn].freq = (uint32_t)freq[s]; nd[nn].left = nd[nn].right = -1; nd[nn].sym = s;
        heap[size++] = nn++;
        sift_up(heap, nd, size - 1);
    }
    int distinct = nn;
    if (distinct == 0) return 0;
    while (size > 1) {
        int l = heap_pop(heap, &size, nd);
        int r = heap_pop(heap, &size, nd);
        nd[nn].freq = (uint32_t)(nd[l].freq + nd[r].freq); nd[nn].left = l; nd[nn].right = r; nd[nn].sym = 0;
        heap[size++] = nn++;
        sift_up(heap, nd, size - 1);
    }
    int root = heap[0];
    if (root_out) *root_out = root;
    if (nodes_out) for (int i = 0; i < nn; ++i) { nodes_out[3 * i] = nd[i].left; nodes_out[3 * i + 1] = nd[i].right; nodes_out[3 * i + 2] = nd[i].sym; }
    /* iterative DFS, code is kept in 64 bits only to detect >32-bit depth (U7) */
    struct { int node; uint32_t code; int len; } st[2 * PORT_HUFF_MAX_SYMS];
    int sp = 0;
    st[sp].node = root; st[sp].code = 0; st[sp].len = 0; ++sp;
    while (sp) {
        --sp;
        int v = st[sp].node; uint32_t c = st[sp].code; int l = st[sp].len;
        if (nd[v].left < 0 && nd[v].right < 0) { codes[nd[v].sym] = c; lens[nd[v].sym] = (uint8_t)l; continue; }
        c <<= 1;
        st[sp].node = nd[v].right; st[sp].code = c + 1; st[sp].len = l + 1; ++sp;
        st[sp].node = nd[v].left;  st[sp].code = c;     st[sp].len = l + 1; ++sp;
    }
    return distinct;
}

int port_huffman_build(const uint64_t* freq, uint32_t* codes, uint8_t* lens, int* nodes_out, int* root_out) {
    return port_huffman_build_n(freq, 256, codes, lens, nodes_out, root_out);
}

void port_histogram(const uint8_t* in, uint64_t n, uint64_t* freq) {
    memset(freq, 0, 256 * sizeof(uint64_t));
    for (uint64_t i = 0; i < n; ++i) ++freq[in[i]];
}

/* MSB-first packing. words must be zeroed and hold n/4+2 entries. Returns total
 * bits; word_idx = bits/32, bit_idx = bits%32 as in BitWriter. */
uint64_t port_huffman_encode(const uint8_t* in, uint64_t n, const uint32_t* codes, const uint8_t* lens, uint32_t* words) {
    uint64_t bitpos = 0;
    for (uint64_t i = 0; i < n; ++i) {
        uint32_t code = codes[in[i]];
        uint32_t len = lens[in[i]];
        uint64_t w = bitpos >> 5;
        uint32_t used = (uint32_t)(bitpos & 31);
        uint32_t room = 32 - used;
        if (len <= room) {
            words[w] |= (len == 32 ? code : (code & ((1u << len) - 1))) << (room - len);
        } else {
            uint32_t spill = len - room;
            words[w] |= (code >> spill) & (room == 32 ? 0xFFFFFFFFu : ((1u << room) - 1));
            words[w + 1] |= code << (32 - spill);
        }
        bitpos += len;
    }
    return bitpos;
}

/* Whole-buffer compress = huffman_compress (huffman.c:288-328).
 * Returns 0, or 1 for the reference's exit(1) case (fewer than 2 distinct symbols). */
int port_huffman_compress(const uint8_t* in, uint64_t n, uint32_t* words, uint64_t* word_idx, uint64_t* bit_idx,
                          uint64_t* buffer_size, uint32_t* codes, uint8_t* lens) {
    uint64_t freq[256];
    port_histogram(in, n, freq);
    int distinct = port_huffman_build(freq, codes, lens, NULL, NULL);
    if (distinct < 2) return 1;
    memset(words, 0, (n / 4 + 2) * 4);
    uint64_t bits = port_huffman_encode(in, n, codes, lens, words);
    *word_idx = bits >> 5;
    *bit_idx = bits & 31;
    *buffer_size = (bits >> 5) * 4 + ((bits & 31) / 8) + (((bits & 31) % 8) > 0);
    return 0;
}

/* Tree-walk decoder with the reference's termination rule: decode whole symbols
 * until the number of consumed bits, floored to bytes, reaches buffer_size
 * (huffman.c:344-361). Bits past the stream read as 0. Returns symbols written
 * (at most out_cap). */
uint64_t port_huffman_decompress(const uint32_t* words, uint64_t nwords, uint64_t buffer_size,
                                 const uint32_t* codes, const uint8_t* lens, uint8_t* out, uint64_t out_cap) {
    /* rebuild a pointer-free tree from the table */
    int left[512], right[512], sym[512], nn = 1;
    left[0] = right[0] = -1; sym[0] = 0;
    for (int s = 0; s < 256; ++s) {
        if (!lens[s]) continue;
        int cur = 0;
        for (int b = lens[s] - 1; b >= 0; --b) {
            int bit = (codes[s] >> b) & 1;
            int* nx = bit ? &right[cur] : &left[cur];
            if (*nx < 0) { *nx = nn; left[nn] = right[nn] = -1; sym[nn] = 0; ++nn; }
            cur = *nx;
        }
        sym[cur] = s;
    }
    uint64_t consumed = 0, o = 0;
    do {
        int v = 0;
        while (left[v] >= 0 && right[v] >= 0) {
            uint64_t w = consumed >> 5;
            uint32_t word = w < nwords ? words[w] : 0;
            v = (word >> (31 - (consumed & 31))) & 1 ? right[v] : left[v];
            ++consumed;
        }
        if (o < out_cap) out[o] = (uint8_t)sym[v];
        ++o;
    } while ((consumed >> 3) < buffer_size);
    return o;
}
