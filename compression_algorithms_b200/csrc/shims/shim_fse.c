/* C mirror of algorithms/fse/src/main.zig (see include/b200_fse.h). */
#include <string.h>
#include "b200_fse.h"
#include "shim_common.h"

#define DIE(...) do { fprintf(stderr, __VA_ARGS__); exit(1); } while (0)
#define GPU_CHECK(expr) do { if ((expr) != B200_OK) DIE("error: %s\n", b200_last_error()); } while (0)

void fse_build_frequency_table(const uint8_t* input, size_t n, size_t freq[256]) {
    memset(freq, 0, 256 * sizeof(size_t));
    if (n == 0) return;
    uint32_t f32[256];
    GPU_CHECK(b200_fse_normalize_host(shim_ctx(), input, n, f32, NULL));
    for (int s = 0; s < 256; ++s) freq[s] = f32[s];
}

void fse_normalize_frequency_table(size_t freq[256]) {
    uint32_t f32[256]; uint16_t norm[256];
    size_t total = 0;
    for (int s = 0; s < 256; ++s) {
        if (freq[s] > 0xFFFFFFFFull) DIE("error: symbol count above 2^32 - 1\n");
        f32[s] = (uint32_t)freq[s]; total += freq[s];
    }
    if (total == 0) return;
    GPU_CHECK(b200_fse_tables_host(shim_ctx(), f32, NULL, norm, NULL));
    for (int s = 0; s < 256; ++s) freq[s] = norm[s];
}

void fse_build_transition_table(const size_t norm[256], TT_Entry tt[TT_SIZE]) {
    uint16_t n16[256];
    size_t total = 0;
    for (int s = 0; s < 256; ++s) { n16[s] = (uint16_t)norm[s]; total += norm[s]; }
    if (total != TT_SIZE) DIE("error: normalised counts sum to %zu, not %d\n", total, TT_SIZE);
    GPU_CHECK(b200_fse_tables_host(shim_ctx(), NULL, n16, NULL, tt));
}

size_t fse_compress_bound(size_t n) { return (size_t)b200_fse_container_max_words(n, FSE_BLOCK_SIZE, FSE_SEGMENT_SIZE); }

size_t fse_compress(const uint8_t* input, size_t n, uint64_t* output) {
    uint64_t words = 0;
    GPU_CHECK(b200_fse_compress_host(shim_ctx(), input, n, FSE_BLOCK_SIZE, FSE_SEGMENT_SIZE, output, fse_compress_bound(n), &words));
    return (size_t)words;
}

size_t fse_decompressed_size(const uint64_t* compressed, size_t words) {
    uint64_t n = 0;
    return b200_fse_container_size(compressed, words, &n) == B200_OK ? (size_t)n : 0;
}

size_t fse_decompress(const uint64_t* compressed, size_t words, uint8_t* output) {
    uint64_t n = 0;
    GPU_CHECK(b200_fse_container_size(compressed, words, &n));
    GPU_CHECK(b200_fse_decompress_host(shim_ctx(), compressed, words, output, n, &n));
    return (size_t)n;
}
