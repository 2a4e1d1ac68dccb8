/* Drop-in for algorithms/deflate (see include/b200_deflate.h). Host code stays C; match
 * finding, greedy parse, token emission, block compaction and decoding run on the GPU. */
#include <string.h>
#include <time.h>
#include "b200_deflate.h"
#include "shim_common.h"

#define IDX_MAGIC 0x3158444946454442ull /* "BDEFIDX1" */
static const char* extension = ".deflate";
#define DIE(...) do { fprintf(stderr, __VA_ARGS__); exit(1); } while (0)
#define GPU_CHECK(expr) do { if ((expr) != B200_OK) DIE("Error: %s\n", b200_last_error()); } while (0)

/* ---- host helpers of the public headers ------------------------------------------------- */
uint64_t min(uint64_t a, uint64_t b) { return a < b ? a : b; }
uint64_t max(uint64_t a, uint64_t b) { return a > b ? a : b; }

uint32_t hash(uint32_t k) {   /* murmur3 mix of one u32 key, seed 0, no length xor */
    k *= 0xcc9e2d51u; k = (k << 15) | (k >> 17); k *= 0x1b873593u;
    uint32_t h = k;
    h = ((h << 13) | (h >> 19)) * 5u + 0xe6546b64u;
    h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
    return h % TABLE_SIZE;
}

void init_hash_table(HashTableArray* t) {
    /* the reference leaves is_set[] and indices[] uninitialised (lz77.c:47-49) */
    t->buckets.patterns = (uint32_t*)calloc(TABLE_SIZE, sizeof(uint32_t));
    t->buckets.indices = (uint64_t*)calloc(TABLE_SIZE, sizeof(uint64_t));
    t->buckets.is_set = (bool*)calloc(TABLE_SIZE, sizeof(bool));
    memset(t->bucket_indices, 0, sizeof(t->bucket_indices));
    t->current_idx = 0;
    t->is_full = false;
}

void insert_hash_table(HashTableArray* t, uint32_t pattern, uint64_t index) {
    uint32_t s = hash(pattern);
    while (t->buckets.is_set[s]) s = (s + 1) % TABLE_SIZE;
    t->buckets.patterns[s] = pattern; t->buckets.indices[s] = index; t->buckets.is_set[s] = true;
    if (t->is_full) {   /* evict the slot recorded WINDOW_SIZE inserts ago (one early: SURVEY.md U10) */
        const uint32_t old = t->bucket_indices[t->current_idx];
        t->buckets.patterns[old] = 0; t->buckets.indices[old] = 0; t->buckets.is_set[old] = false;
    }
    t->bucket_indices[t->current_idx++] = s;
    if (t->current_idx >= WINDOW_SIZE - 1) t->is_full = true;
    t->current_idx %= WINDOW_SIZE;
}

uint64_t find(HashTableArray* t, uint32_t pattern) {
    uint32_t s = hash(pattern);
    while (t->buckets.is_set[s] && t->buckets.patterns[s] != pattern) ++s;   /* no wrap (lz77.c:168) */
    return t->buckets.is_set[s] ? t->buckets.indices[s] : UINT64_MAX;
}

void write_literal(char* buffer, char c, uint64_t* i) { buffer[(*i)++] = 0; buffer[(*i)++] = c; }
void write_length_distance(char* buffer, uint8_t length, uint16_t distance, uint64_t* i) {
    buffer[(*i)++] = 1; buffer[(*i)++] = (char)(distance & 0xFF); buffer[(*i)++] = (char)(distance >> 8); buffer[(*i)++] = (char)length;
}

void init_bitwriter(BitWriter* w, uint64_t buffer_size) {
    w->buffer = (uint32_t*)calloc(buffer_size ? buffer_size : 1, 1);
    w->word_idx = 0; w->bit_idx = 0; w->buffer_size = buffer_size;
}
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
void append_huffman_tree_literal(uint32_t* frequencies, char literal) { ++frequencies[(uint8_t)literal]; }
void append_huffman_tree_pair(uint32_t* frequencies, uint16_t offset) {
    ++frequencies[256 + (uint8_t)(__builtin_clz((unsigned)offset) - 16)];   /* distance bit-width class */
}
void gather_codes(MinHeapNode* root, uint16_t code, uint8_t length, uint16_t* codes, uint8_t* code_lengths) {
    if (!root->left && !root->right) { codes[root->data] = code; code_lengths[root->data] = length; return; }
    if (root->left) gather_codes(root->left, (uint16_t)(code << 1), (uint8_t)(length + 1), codes, code_lengths);
    if (root->right) gather_codes(root->right, (uint16_t)((code << 1) | 1), (uint8_t)(length + 1), codes, code_lengths);
}
void init_huffman_node(HuffmanNode* n) { n->left = NULL; n->right = NULL; n->value = 0; n->frequency = 0; }
void destroy_huffman_node(HuffmanNode* n) {
    if (n->left) destroy_huffman_node(n->left);
    if (n->right) destroy_huffman_node(n->right);
    free(n->left); free(n->right);
    n->left = n->right = NULL;
}
bool compare_huffman_node(const HuffmanNode* a, const HuffmanNode* b) { return a->frequency < b->frequency; }

/* ---- GPU-backed entry points ------------------------------------------------------------ */
/* Every GPU of the box works on a buffer of many blocks (blocks are independent: fresh table per block), one host
 * thread per GPU inside libb200comp.so (b200_lz77_compress_multi_host). B200_DEVICES = "n" limits the count
 * (default: all visible devices); buffers below 64 blocks per device stay on one GPU. */
static b200_multi* shim_multi(uint64_t nblocks) {
    static b200_multi* m = NULL;
    static int tried = 0;
    if (!tried) {
        tried = 1;
        int want = b200_device_count();
        const char* e = getenv("B200_DEVICES");
        if (e && atoi(e) > 0 && atoi(e) < want) want = atoi(e);
        if (want > 1 && b200_multi_create(&m, NULL, want) != B200_OK) { fprintf(stderr, "warning: multi-GPU setup failed (%s); using one GPU\n", b200_last_error()); m = NULL; }
    }
    if (m && nblocks < 64ull * (uint64_t)b200_multi_device_count(m)) return NULL;
    return m;
}

uint64_t deflate_compress_buffer(const char* in, uint64_t size, uint64_t block_size, char* out, uint64_t* block_off) {
    if (size == 0) { block_off[0] = 0; return 0; }
    const uint64_t bs = (block_size == 0 || block_size > size) ? size : block_size;
    const uint64_t nblocks = (size + bs - 1) / bs;
    uint64_t* sizes = (uint64_t*)malloc(nblocks * sizeof(uint64_t));
    uint64_t total = 0;
    b200_multi* m = shim_multi(nblocks);
    if (m) GPU_CHECK(b200_lz77_compress_multi_host(m, B200_LZ_DEFLATE, (const uint8_t*)in, size, block_size, (uint8_t*)out,
                                                   b200_lz77_max_bytes(B200_LZ_DEFLATE, size, block_size), sizes, block_off, &total));
    else GPU_CHECK(b200_lz77_compress_host(shim_ctx(), B200_LZ_DEFLATE, (const uint8_t*)in, size, block_size, (uint8_t*)out,
                                           b200_lz77_max_bytes(B200_LZ_DEFLATE, size, block_size), sizes, block_off, &total));
    free(sizes);
    return total;
}

void deflate_decompress_buffer(const char* tokens, uint64_t token_bytes, const uint64_t* block_off, uint64_t size,
                               uint64_t block_size, char* out) {
    if (size == 0) return;
    const uint64_t bs = (block_size == 0 || block_size > size) ? size : block_size;
    const uint64_t nblocks = (size + bs - 1) / bs;
    uint64_t* sizes = (uint64_t*)malloc(nblocks * sizeof(uint64_t));
    for (uint64_t b = 0; b < nblocks; ++b) sizes[b] = block_off[b + 1] - block_off[b];
    b200_multi* m = shim_multi(nblocks);
    if (m) GPU_CHECK(b200_lz77_decompress_multi_host(m, B200_LZ_DEFLATE, (const uint8_t*)tokens, token_bytes, block_off, sizes, size,
                                                     block_size, (uint8_t*)out));
    else GPU_CHECK(b200_lz77_decompress_host(shim_ctx(), B200_LZ_DEFLATE, (const uint8_t*)tokens, token_bytes, block_off, sizes, size,
                                             block_size, (uint8_t*)out));
    free(sizes);
}

void lz77_compress(const char* input_buffer, uint64_t input_buffer_size, char* compressed_buffer,
                   uint64_t* compressed_buffer_size, HashTableArray* table) {
    (void)table;   /* every call parses against a fresh table; the caller's table is left untouched */
    uint64_t off[2] = {0, 0};
    *compressed_buffer_size = 0;
    if (input_buffer_size == 0) return;
    /* one block = the whole buffer; the host wrapper wants 2n + 64 bytes of room, the
     * reference's contract is 2n: stage through a private buffer */
    char* tmp = (char*)malloc(b200_lz77_max_bytes(B200_LZ_DEFLATE, input_buffer_size, 0));
    const uint64_t total = deflate_compress_buffer(input_buffer, input_buffer_size, 0, tmp, off);
    memcpy(compressed_buffer, tmp, total);
    free(tmp);
    *compressed_buffer_size = total;
}

/* number of bytes a token stream decodes to (2-byte literals, 4-byte matches) */
static uint64_t decoded_length(const uint8_t* t, uint64_t bytes) {
    uint64_t n = 0, i = 0;
    while (i < bytes) {
        if (t[i] == 0) { n += 1; i += 2; }
        else { if (i + 4 > bytes) break; n += t[i + 3]; i += 4; }
    }
    return n;
}

void lz77_decompress(const char* compressed_buffer, uint64_t compressed_buffer_size, char* decompressed_buffer,
                     uint64_t* decompressed_buffer_size) {
    const uint64_t capacity = *decompressed_buffer_size;
    const uint64_t n = decoded_length((const uint8_t*)compressed_buffer, compressed_buffer_size);
    *decompressed_buffer_size = 0;
    if (n == 0) return;
    if (n > capacity) DIE("Error: block decodes to %lu bytes, buffer has %lu\n", (unsigned long)n, (unsigned long)capacity);
    const uint64_t off[2] = {0, compressed_buffer_size};
    deflate_decompress_buffer(compressed_buffer, compressed_buffer_size, off, n, 0, decompressed_buffer);
    *decompressed_buffer_size = n;
}

static char* slurp(const char* path, uint64_t* size) {
    FILE* f = fopen(path, "rb");
    if (!f) return NULL;
    fseek(f, 0, SEEK_END); *size = (uint64_t)ftell(f); fseek(f, 0, SEEK_SET);
    char* p = (char*)malloc(*size + 64);
    if (fread(p, 1, *size, f) != *size) { fclose(f); free(p); return NULL; }
    fclose(f);
    return p;
}

StateData compress(const char* input_filename) {
    static HashTableArray table;   /* the reference returns the address of a stack object (deflate.c:13,16) */
    const char* slash = strrchr(input_filename, '/');
    const char* filename = slash ? slash + 1 : input_filename;
    StateData state = { &table, NULL, (char*)malloc(strlen(filename) + strlen(extension) + 1) };
    strcpy(state.compressed_filename, filename);
    strcat(state.compressed_filename, extension);

    uint64_t size = 0;
    char* in = slurp(input_filename, &size);
    if (!in) DIE("Error: could not open file %s\n", input_filename);
    FILE* out_file = fopen(state.compressed_filename, "wb");
    if (!out_file) DIE("Error: could not open file %s\n", state.compressed_filename);

    const clock_t start = clock();
    const uint64_t nblocks = (size + BUFFER_SIZE - 1) / BUFFER_SIZE;
    char* out = (char*)malloc(2 * size + 64);
    uint64_t* off = (uint64_t*)calloc(nblocks + 1, sizeof(uint64_t));
    const uint64_t total = deflate_compress_buffer(in, size, BUFFER_SIZE, out, off);
    fwrite(out, 1, total, out_file);
    fclose(out_file);
    printf("MB/s: %f\n", (double)size / (1024 * 1024) / ((double)(clock() - start) / CLOCKS_PER_SEC));

    char* idx_name = (char*)malloc(strlen(state.compressed_filename) + 5);
    strcpy(idx_name, state.compressed_filename); strcat(idx_name, ".idx");
    FILE* idx = fopen(idx_name, "wb");
    if (idx) {
        const uint64_t hdr[4] = { IDX_MAGIC, size, BUFFER_SIZE, nblocks };
        fwrite(hdr, sizeof(uint64_t), 4, idx);
        fwrite(off, sizeof(uint64_t), nblocks + 1, idx);
        fclose(idx);
    }
    free(idx_name); free(off); free(out); free(in);
    return state;
}

void decompress(StateData* state_data, const char* input_filename) {
    const char* path = input_filename ? input_filename : (state_data ? state_data->compressed_filename : NULL);
    if (!path) DIE("Error: no compressed file given\n");
    uint64_t tbytes = 0;
    char* tokens = slurp(path, &tbytes);
    if (!tokens) DIE("Error: could not open file %s\n", path);
    char* idx_name = (char*)malloc(strlen(path) + 5);
    strcpy(idx_name, path); strcat(idx_name, ".idx");
    uint64_t ibytes = 0, size = 0, nblocks = 0;
    uint64_t* off = NULL;
    uint64_t* idx = (uint64_t*)slurp(idx_name, &ibytes);
    if (idx && ibytes >= 40 && idx[0] == IDX_MAGIC && idx[2] == BUFFER_SIZE && ibytes >= (5 + idx[3]) * 8 && idx[4 + idx[3]] == tbytes) {
        size = idx[1]; nblocks = idx[3];
        off = (uint64_t*)malloc((nblocks + 1) * sizeof(uint64_t));
        memcpy(off, idx + 4, (nblocks + 1) * sizeof(uint64_t));
    } else {
        /* no index (e.g. a file written by the reference): find the block boundaries by walking
         * the token flags -- a block ends when its tokens cover BUFFER_SIZE input bytes */
        uint64_t cap = tbytes / (2 * BUFFER_SIZE) * 2 + 16, i = 0, covered = 0;
        off = (uint64_t*)malloc(cap * sizeof(uint64_t));
        off[0] = 0;
        const uint8_t* t = (const uint8_t*)tokens;
        while (i < tbytes) {
            if (t[i] == 0) { covered += 1; i += 2; } else { if (i + 4 > tbytes) break; covered += t[i + 3]; i += 4; }
            if (covered >= BUFFER_SIZE || i >= tbytes) {
                if (nblocks + 2 > cap) { cap *= 2; off = (uint64_t*)realloc(off, cap * sizeof(uint64_t)); }
                off[++nblocks] = i;
                size += covered < BUFFER_SIZE ? covered : BUFFER_SIZE;
                covered = 0;
            }
        }
    }
    free(idx); free(idx_name);
    char* out = (char*)malloc(size + 64);
    deflate_decompress_buffer(tokens, tbytes, off, size, BUFFER_SIZE, out);
    char* out_name = (char*)malloc(strlen(path) + 5);
    strcpy(out_name, path); strcat(out_name, ".out");
    FILE* f = fopen(out_name, "wb");
    if (!f) DIE("Error: could not open file %s\n", out_name);
    fwrite(out, 1, size, f);
    fclose(f);
    free(out_name); free(out); free(off); free(tokens);
}

/* ---- the completed pipeline the reference sketches (deflate/lz77.c:279 "TODO: Build huffman tree and encode
 * compressed buffer"): <name>.dfl = ONE self-describing stream (b200comp.h "containers": frequencies[286] of every
 * block, token sizes, decode index, Huffman-coded words). Nothing else is needed to decode it. */
uint64_t compress_entropy(const char* input_filename, const char* output_filename) {
    b200_ctx* ctx = shim_ctx();
    uint64_t size = 0, total = 0;
    char* in = slurp(input_filename, &size);
    if (!in) DIE("Error: could not open file %s\n", input_filename);
    const uint64_t cap = b200_deflate_container_max_bytes(size, BUFFER_SIZE);
    void* out = malloc(cap);
    SHIM_CHECK(b200_deflate_compress_container_host(ctx, (const uint8_t*)in, size, BUFFER_SIZE, out, cap, &total));
    FILE* f = fopen(output_filename, "wb");
    if (!f || fwrite(out, 1, total, f) != total) DIE("Error: could not write file %s\n", output_filename);
    fclose(f);
    free(out); free(in);
    return total;
}

uint64_t decompress_entropy(const char* input_filename, const char* output_filename) {
    b200_ctx* ctx = shim_ctx();
    uint64_t bytes = 0, n = 0, bs = 0; uint32_t codec = 0;
    char* c = slurp(input_filename, &bytes);
    if (!c) DIE("Error: could not open file %s\n", input_filename);
    SHIM_CHECK(b200_container_info(c, bytes, &codec, &n, &bs));
    uint8_t* out = (uint8_t*)malloc(n + 64);
    SHIM_CHECK(b200_deflate_decompress_container_host(ctx, c, bytes, out, n, &n));
    FILE* f = fopen(output_filename, "wb");
    if (!f || fwrite(out, 1, n, f) != n) DIE("Error: could not write file %s\n", output_filename);
    fclose(f);
    free(out); free(c);
    return n;
}
