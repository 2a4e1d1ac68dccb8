/* Drop-in for algorithms/lz77 (see include/b200_lz77.h). Host code stays C. */
#include <string.h>
#include "b200_lz77.h"
#include "shim_common.h"

uint64_t min(uint64_t a, uint64_t b) { return a < b ? a : b; }
uint64_t max(uint64_t a, uint64_t b) { return a > b ? a : b; }

uint32_t hash(uint32_t k) {
    k *= 0xcc9e2d51u; k = (k << 15) | (k >> 17); k *= 0x1b873593u;
    uint32_t h = k;
    h = ((h << 13) | (h >> 19)) * 5u + 0xe6546b64u;
    h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
    return h % (1u << (WINDOW_BITS + 6));
}

/* table helpers of lz77.h:33-37. A guard tail behind the 2^20 slots takes the probes that run past
 * the end (the reference indexes out of bounds there). */
#define SHIM_GUARD (1u << WINDOW_BITS)
void init_hash_table(HashTableArray* t) {
    t->buckets = (ArrayNode*)calloc((size_t)TABLE_SIZE + SHIM_GUARD, sizeof(ArrayNode));
    memset(t->bucket_indices, 0, sizeof(t->bucket_indices));
    t->current_idx = 0;
    t->is_full = false;
}

void insert_hash_table(HashTableArray* t, uint32_t pattern, uint64_t index) {
    uint32_t s = hash(pattern);
    while (t->buckets[s].is_set) ++s;                      /* no wrap (lz77.c:61) */
    t->buckets[s].pattern = pattern; t->buckets[s].index = index; t->buckets[s].is_set = true;
    if (t->is_full) {   /* evict the slot recorded one window ago (flips one insert early: SURVEY.md U10) */
        ArrayNode* old = &t->buckets[t->bucket_indices[t->current_idx]];
        old->pattern = 0; old->index = 0; old->is_set = false;
    }
    t->bucket_indices[t->current_idx++] = s;
    if (t->current_idx >= (1u << WINDOW_BITS) - 1) t->is_full = true;
    t->current_idx %= (1u << WINDOW_BITS);
}

uint64_t find(HashTableArray* t, uint32_t pattern) {
    uint32_t s = hash(pattern);
    while (t->buckets[s].is_set && t->buckets[s].pattern != pattern) ++s;
    return t->buckets[s].is_set ? t->buckets[s].index : UINT64_MAX;
}

void print_bit_string(const char* buffer, uint64_t size) {
    for (uint64_t i = 0; i < size; ++i) for (int b = 7; b >= 0; --b) putchar(((unsigned char)buffer[i] >> b) & 1 ? '1' : '0');
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

void init_bitstream(BitStream* s, uint8_t* buffer) { s->data = buffer; s->bit_index = 0; }
void write_bit(BitStream* s, bool bit) {
    uint64_t by = s->bit_index / 8, off = s->bit_index % 8;
    if (bit) s->data[by] |= (uint8_t)(1u << off); else s->data[by] &= (uint8_t)~(1u << off);
    ++s->bit_index;
}
bool read_bit(BitStream* s) { bool b = (s->data[s->bit_index / 8] >> (s->bit_index % 8)) & 1; ++s->bit_index; return b; }
void write_bits(BitStream* s, uint64_t value, uint64_t num_bits) { for (uint64_t b = 0; b < num_bits; ++b) write_bit(s, (value >> b) & 1); }
uint64_t read_bits(BitStream* s, uint64_t num_bits) { uint64_t v = 0; for (uint64_t b = 0; b < num_bits; ++b) if (read_bit(s)) v |= 1ull << b; return v; }

bool check_buffer_equivalence(const char* a, const char* b, uint64_t size) {
    uint64_t diff = 0;
    for (uint64_t i = 0; i < size; ++i) diff += a[i] != b[i];
    printf("Number of differences: %lu\n", (unsigned long)diff);
    return diff == 0;
}

BitStream* lz77_compress(const char* buffer, uint64_t size) {
    b200_ctx* ctx = shim_ctx();
    BitStream* stream = (BitStream*)malloc(sizeof(BitStream));
    uint64_t cap = b200_lz77_max_bytes(B200_LZ_STANDALONE, size, 0);
    uint8_t* out = (uint8_t*)malloc(cap);
    uint64_t sizes[1] = {0}, off[2] = {0, 0}, total = 0;
    /* block_size 0: the whole buffer is one block = the reference's unblocked call */
    SHIM_CHECK(b200_lz77_compress_host(ctx, B200_LZ_STANDALONE, (const uint8_t*)buffer, size, 0, out, cap, sizes, off, &total));
    stream->bit_index = sizes[0];
    stream->data = (uint8_t*)realloc(out, stream->bit_index / 8 + 1);   /* lz77.c:341-342 */
    return stream;
}

char* lz77_decompress(BitStream* s, uint64_t size, uint64_t* decompressed_size) {
    b200_ctx* ctx = shim_ctx();
    char* out = (char*)malloc(size + 16);
    uint64_t sizes[1] = {s->bit_index}, off[2] = {0, s->bit_index / 8 + 1};
    s->bit_index = 0;   /* the reference resets the read position (lz77.c:356) */
    SHIM_CHECK(b200_lz77_decompress_host(ctx, B200_LZ_STANDALONE, s->data, off[1], off, sizes, size, 0, (uint8_t*)out));
    *decompressed_size = size;
    return out;
}
