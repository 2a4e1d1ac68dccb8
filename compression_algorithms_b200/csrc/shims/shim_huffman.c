/* Drop-in for algorithms/huffman (see include/b200_huffman.h). Host code stays C; the
 * histogram, the heap-exact table build, the bit packing and the decoding run on the GPU. */
#include <string.h>
#include "b200_huffman.h"
#include "shim_common.h"

/* ---- small host helpers of the reference's public header ------------------------------ */
void init_bitwriter(BitWriter* w, uint64_t buffer_size) {
    w->buffer = (uint32_t*)calloc(buffer_size ? buffer_size : 1, 1);
    w->word_idx = 0;
    w->bit_idx = 0;
    w->buffer_size = buffer_size;
}

/* MSB-first append of `length` bits (huffman.c:18-48) */
void write_bits(BitWriter* w, uint32_t bits, uint8_t length) {
    const uint32_t room = 32u - (uint32_t)w->bit_idx;
    if (length == 0) return;
    if (length < 32) bits &= (1u << length) - 1u;
    if (length <= room) {
        w->buffer[w->word_idx] |= (length == 32) ? bits : bits << (room - length);
        w->bit_idx += length;
        if (w->bit_idx == 32) { w->bit_idx = 0; ++w->word_idx; }
    } else {
        const uint32_t spill = length - room;
        w->buffer[w->word_idx] |= bits >> spill;
        ++w->word_idx;
        w->buffer[w->word_idx] |= bits << (32u - spill);
        w->bit_idx = spill;
    }
}

PriorityQueue* init_priority_queue(uint64_t capacity) {
    PriorityQueue* q = (PriorityQueue*)malloc(sizeof(*q));
    q->nodes = (Node**)malloc(capacity * sizeof(Node*));
    q->size = 0;
    q->capacity = capacity;
    return q;
}
void swap_nodes(Node** a, Node** b) { Node* t = *a; *a = *b; *b = t; }
void heapify_up(PriorityQueue* q, uint64_t i) {
    while (i > 0) {
        const uint64_t p = (i - 1) / 2;
        if (!(q->nodes[i]->frequency < q->nodes[p]->frequency)) break;
        swap_nodes(&q->nodes[i], &q->nodes[p]);
        i = p;
    }
}
void heapify_down(PriorityQueue* q, uint64_t i) {
    for (;;) {
        const uint64_t l = 2 * i + 1, r = l + 1;
        uint64_t m = i;
        if (l < q->size && q->nodes[l]->frequency < q->nodes[m]->frequency) m = l;
        if (r < q->size && q->nodes[r]->frequency < q->nodes[m]->frequency) m = r;
        if (m == i) return;
        swap_nodes(&q->nodes[i], &q->nodes[m]);
        i = m;
    }
}
void enqueue(PriorityQueue* q, Node* node) {
    if (q->size == q->capacity) { printf("ERROR: Queue is full\n"); exit(1); }
    q->nodes[q->size++] = node;
    heapify_up(q, q->size - 1);
}
Node* dequeue(PriorityQueue* q) {
    if (q->size == 0) { printf("ERROR: Queue is empty\n"); exit(1); }
    Node* top = q->nodes[0];
    q->nodes[0] = q->nodes[--q->size];
    heapify_down(q, 0);
    return top;
}
bool is_empty(PriorityQueue* q) { return q->size == 0; }
Node* init_node(uint8_t value, uint32_t frequency) {
    Node* n = (Node*)malloc(sizeof(*n));
    n->value = value; n->frequency = frequency; n->left = NULL; n->right = NULL;
    return n;
}

void print_bit_string(uint8_t* buffer, uint64_t size) {
    for (uint64_t i = 0; i < size; ++i) {
        for (int b = 7; b >= 0; --b) putchar((buffer[i] >> b) & 1 ? '1' : '0');
        putchar(' ');
    }
    putchar('\n');
}

char* read_input_buffer(const char* filename, uint64_t* size) {
    FILE* f = fopen(filename, "rb");
    if (!f) { printf("ERROR: cannot open %s\n", filename); exit(1); }
    fseek(f, 0, SEEK_END); *size = (uint64_t)ftell(f); fseek(f, 0, SEEK_SET);
    char* buffer = (char*)malloc(*size + 1);
    size_t got = fread(buffer, 1, *size, f); (void)got;
    fclose(f);
    return buffer;
}

void gather_codes(Node* root, uint32_t code, uint32_t length, uint32_t* codes, uint8_t* code_lengths) {
    if (!root->left && !root->right) {
        codes[root->value] = code;
        code_lengths[root->value] = (uint8_t)length;
        return;
    }
    if (root->left) gather_codes(root->left, code << 1, length + 1, codes, code_lengths);
    if (root->right) gather_codes(root->right, (code << 1) | 1u, length + 1, codes, code_lengths);
}

void print_codes(uint32_t* codes, uint8_t* code_lengths) {
    for (int s = 0; s < 256; ++s) {
        if (!code_lengths[s]) continue;
        printf("%c: ", (char)s);
        for (int b = code_lengths[s] - 1; b >= 0; --b) putchar((codes[s] >> b) & 1 ? '1' : '0');
        putchar('\n');
    }
}

/* ---- GPU-backed entry points ------------------------------------------------------------ */

/* The decode index (bit offsets of the 256-symbol sub-chunks) is not part of the words the reference compares and
 * huffman_decompress(writer, root, ...) has no parameter that could carry it, so the library keeps it for every
 * stream it produced, keyed by the words pointer and the exact bit count, behind a mutex: any number of live
 * streams, from any thread. A stream that is not in the registry (e.g. words written by the reference itself) is
 * decoded without an index. Streams that leave the process go through the container calls below instead. */
#include <pthread.h>
typedef struct Stream { struct Stream* next; const uint32_t* buffer; uint64_t n, total_words, total_bits; uint8_t* side; uint64_t side_bytes; } Stream;
static Stream* g_streams = NULL;
static pthread_mutex_t g_streams_mu = PTHREAD_MUTEX_INITIALIZER;

static void remember(const uint32_t* buffer, uint64_t n, uint64_t total_words, uint64_t total_bits, uint8_t* side, uint64_t side_bytes) {
    pthread_mutex_lock(&g_streams_mu);
    Stream* e = g_streams;
    while (e && e->buffer != buffer) e = e->next;       /* the same words buffer written again replaces its entry */
    if (!e) { e = (Stream*)calloc(1, sizeof(Stream)); e->next = g_streams; g_streams = e; }
    free(e->side);
    e->buffer = buffer; e->n = n; e->total_words = total_words; e->total_bits = total_bits; e->side = side; e->side_bytes = side_bytes;
    pthread_mutex_unlock(&g_streams_mu);
}

/* copy of the entry for (buffer, bits), side index included, so that the caller works outside the lock */
static bool lookup(const uint32_t* buffer, uint64_t total_bits, Stream* out) {
    bool hit = false;
    pthread_mutex_lock(&g_streams_mu);
    for (Stream* e = g_streams; e; e = e->next) if (e->buffer == buffer && e->total_bits == total_bits) {
        *out = *e;
        out->side = (uint8_t*)malloc(e->side_bytes);
        memcpy(out->side, e->side, e->side_bytes);
        hit = true;
        break;
    }
    pthread_mutex_unlock(&g_streams_mu);
    return hit;
}

/* host Node tree from the GPU's node array: i16 child[511][2], leaf = {-1, symbol} */
static Node* materialise(const int16_t* tree, const uint32_t* freq, int v) {
    const int l = tree[2 * v], r = tree[2 * v + 1];
    if (l < 0) return init_node((uint8_t)r, freq[r]);
    Node* n = init_node(0, 0);
    n->left = materialise(tree, freq, l);
    n->right = materialise(tree, freq, r);
    n->frequency = n->left->frequency + n->right->frequency;   /* u32 sum like init_node(0, l+r) */
    return n;
}

static Node* tree_from_side(const uint8_t* side, const b200_huff_layout* L) {
    const uint32_t* meta = (const uint32_t*)(side + L->off_meta);
    if (meta[1] == 0) { printf("ERROR: Queue is empty\n"); exit(1); }   /* huffman.c:149-152 */
    return materialise((const int16_t*)(side + L->off_tree), (const uint32_t*)(side + L->off_freq), (int)meta[2]);
}

void build_huffman_tree(char* buffer, uint64_t size, Node** root) {
    b200_ctx* ctx = shim_ctx();
    if (size == 0) { printf("ERROR: Queue is empty\n"); exit(1); }
    b200_huff_layout L;
    SHIM_CHECK(b200_huffman_layout(size, 0, &L));
    uint8_t* side = (uint8_t*)malloc(L.bytes);
    SHIM_CHECK(b200_huffman_tables_host(ctx, (const uint8_t*)buffer, size, 0, side, L.bytes));
    *root = tree_from_side(side, &L);
    free(side);
}

static void set_writer_sizes(BitWriter* w, uint64_t total_bits) {
    w->word_idx = total_bits / 32;
    w->bit_idx = total_bits % 32;
    w->buffer_size = w->word_idx * 4 + w->bit_idx / 8 + (w->bit_idx % 8 > 0);   /* huffman.c:318-320 */
}

Node huffman_compress(char* buffer, uint64_t size, BitWriter* writer) {
    b200_ctx* ctx = shim_ctx();
    if (size == 0) { printf("ERROR: Queue is empty\n"); exit(1); }
    b200_huff_layout L;
    SHIM_CHECK(b200_huffman_layout(size, 0, &L));
    const uint64_t cap = b200_huffman_max_words(size, 0);
    uint8_t* side = (uint8_t*)malloc(L.bytes);
    uint32_t* words = (uint32_t*)malloc(cap * 4);
    uint64_t total_words = 0; uint32_t status = 0;
    const int rc = b200_huffman_compress_host(ctx, (const uint8_t*)buffer, size, 0, words, cap, side, L.bytes, &total_words, &status);
    if (rc == B200_ERR_DOMAIN && status == 1) {   /* root is a leaf -> length 0 -> huffman.c:274-277 */
        printf("ERROR: No code for character %c\n", buffer[0]);
        exit(1);
    }
    SHIM_CHECK(rc);
    const uint64_t total_bits = ((const uint64_t*)(side + L.off_block_bits))[0];
    writer->buffer = (uint32_t*)realloc(words, (total_words ? total_words : 1) * 4);   /* whole words, see the header */
    set_writer_sizes(writer, total_bits);
    Node* root = tree_from_side(side, &L);
    remember(writer->buffer, size, total_words, total_bits, side, L.bytes);
    return *root;
}

void _huffman_compress(char* buffer, uint64_t size, uint32_t* codes, uint8_t* code_lengths, BitWriter* writer) {
    b200_ctx* ctx = shim_ctx();
    if (size == 0) return;
    b200_huff_layout L;
    SHIM_CHECK(b200_huffman_layout(size, 0, &L));
    uint32_t mx = 0;
    for (int s = 0; s < 256; ++s) if (code_lengths[s] > mx) mx = code_lengths[s];
    const uint64_t cap = (size * mx + 31) / 32 + 4;
    uint8_t* side = (uint8_t*)malloc(L.bytes);
    uint32_t* words = (uint32_t*)malloc(cap * 4);
    uint64_t total_words = 0; uint32_t status = 0;
    const int rc = b200_huffman_compress_codes_host(ctx, (const uint8_t*)buffer, size, codes, code_lengths, words, cap, side, L.bytes,
                                                    &total_words, &status);
    if (rc == B200_ERR_DOMAIN && status == 3) {
        uint64_t i = 0;
        while (i < size && code_lengths[(uint8_t)buffer[i]]) ++i;
        printf("ERROR: No code for character %c\n", buffer[i < size ? i : 0]);
        exit(1);
    }
    SHIM_CHECK(rc);
    const uint64_t total_bits = ((const uint64_t*)(side + L.off_block_bits))[0];
    const uint64_t start_bits = writer->word_idx * 32 + writer->bit_idx;
    const uint64_t end_words = (start_bits + total_bits + 31) / 32;
    if (end_words * 4 > ((writer->buffer_size + 3) & ~(uint64_t)3)) {
        printf("ERROR: BitWriter of %lu bytes cannot hold %lu more bits\n", (unsigned long)writer->buffer_size, (unsigned long)total_bits);
        exit(1);
    }
    uint32_t* dst = writer->buffer + writer->word_idx;
    const uint32_t sh = (uint32_t)writer->bit_idx;
    if (sh == 0) {
        for (uint64_t i = 0; i < total_words; ++i) dst[i] |= words[i];
    } else {   /* appending at a bit offset: shift the packed words in behind the existing bits */
        for (uint64_t i = 0; i < total_words; ++i) {
            dst[i] |= words[i] >> sh;
            const uint32_t low = words[i] << (32 - sh);
            if (low) dst[i + 1] |= low;
        }
    }
    const bool fresh = start_bits == 0;
    writer->word_idx = (start_bits + total_bits) / 32;
    writer->bit_idx = (start_bits + total_bits) % 32;
    if (fresh) remember(writer->buffer, size, total_words, total_bits, side, L.bytes);
    else free(side);
    free(words);
}

static uint64_t trailing_symbols(const BitWriter* w, const Node* root, uint64_t consumed, uint64_t nwords) {
    /* the reference stops at the first symbol boundary with consumed/8 >= buffer_size
     * (huffman.c:344-361): the pad bits behind the last code decode as extra symbols */
    uint64_t extra = 0;
    while (consumed / 8 < w->buffer_size) {
        const Node* v = root;
        while (v->left && v->right) {
            const uint64_t wi = consumed / 32;
            const uint32_t word = wi < nwords ? w->buffer[wi] : 0;
            v = (word >> (31 - consumed % 32)) & 1 ? v->right : v->left;
            ++consumed;
        }
        ++extra;
    }
    return extra;
}

void huffman_decompress(BitWriter* writer, Node* root, char* output, uint64_t* output_size) {
    b200_ctx* ctx = shim_ctx();
    const uint64_t capacity = *output_size;
    memset(output, 0, capacity);   /* huffman.c:341 */
    Stream st;
    if (lookup(writer->buffer, writer->word_idx * 32 + writer->bit_idx, &st)) {
        const uint64_t n = st.n;
        if (capacity >= n) {
            SHIM_CHECK(b200_huffman_decompress_host(ctx, writer->buffer, st.total_words, st.side, st.side_bytes, n, 0, (uint8_t*)output));
        } else {
            uint8_t* tmp = (uint8_t*)malloc(n);
            SHIM_CHECK(b200_huffman_decompress_host(ctx, writer->buffer, st.total_words, st.side, st.side_bytes, n, 0, tmp));
            memcpy(output, tmp, capacity);
            free(tmp);
        }
        free(st.side);
        const uint64_t extra = trailing_symbols(writer, root, st.total_bits, st.total_words);
        if (extra && capacity > n) {   /* the symbols the reference decodes out of the pad bits */
            const Node* v = root;
            while (v->left && v->right) v = v->left;   /* pad bits are zero */
            for (uint64_t i = n; i < n + extra && i < capacity; ++i) output[i] = (char)v->value;
        }
        *output_size = n + extra;
        return;
    }
    /* a stream without an index: one GPU thread walks it with the reference's rule */
    uint32_t codes[256] = {0}; uint8_t lens[256] = {0};
    gather_codes(root, 0, 0, codes, lens);
    const uint64_t nwords = (writer->buffer_size + 3) / 4;
    uint64_t count = 0;
    SHIM_CHECK(b200_huffman_decompress_serial_host(ctx, writer->buffer, nwords, writer->buffer_size, codes, lens, (uint8_t*)output, capacity, &count));
    *output_size = count;
}

void huffman_decompress_lookup_table(BitWriter* writer, Node* root, char* output, uint64_t* output_size) {
    huffman_decompress(writer, root, output, output_size);
}

/* ---- file-level helpers on the self-describing container (include/b200comp.h): what a driver that wants to keep
 * the compressed data writes and reads. Returns the container bytes / the decoded bytes. */
static char* slurp_file(const char* path, uint64_t* size) {
    FILE* f = fopen(path, "rb");
    if (!f) { printf("ERROR: cannot open %s\n", path); exit(1); }
    fseek(f, 0, SEEK_END); *size = (uint64_t)ftell(f); fseek(f, 0, SEEK_SET);
    char* p = (char*)malloc(*size + 64);
    if (fread(p, 1, *size, f) != *size) { printf("ERROR: short read on %s\n", path); exit(1); }
    fclose(f);
    return p;
}

uint64_t huffman_compress_file(const char* input_filename, const char* output_filename, uint64_t block_size) {
    b200_ctx* ctx = shim_ctx();
    uint64_t size = 0, total = 0;
    char* in = slurp_file(input_filename, &size);
    if (size == 0) { printf("ERROR: Queue is empty\n"); exit(1); }
    const uint64_t cap = b200_huffman_container_max_bytes(size, block_size);
    void* out = malloc(cap);
    SHIM_CHECK(b200_huffman_compress_container_host(ctx, (const uint8_t*)in, size, block_size, out, cap, &total));
    FILE* f = fopen(output_filename, "wb");
    if (!f || fwrite(out, 1, total, f) != total) { printf("ERROR: cannot write %s\n", output_filename); exit(1); }
    fclose(f);
    free(out); free(in);
    return total;
}

uint64_t huffman_decompress_file(const char* input_filename, const char* output_filename) {
    b200_ctx* ctx = shim_ctx();
    uint64_t bytes = 0, n = 0, bs = 0; uint32_t codec = 0;
    char* c = slurp_file(input_filename, &bytes);
    SHIM_CHECK(b200_container_info(c, bytes, &codec, &n, &bs));
    uint8_t* out = (uint8_t*)malloc(n + 64);
    SHIM_CHECK(b200_huffman_decompress_container_host(ctx, c, bytes, out, n, &n));
    FILE* f = fopen(output_filename, "wb");
    if (!f || fwrite(out, 1, n, f) != n) { printf("ERROR: cannot write %s\n", output_filename); exit(1); }
    fclose(f);
    free(out); free(c);
    return n;
}
