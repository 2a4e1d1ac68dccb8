/* Shaped like /root/reference/algorithms/fse/src/main.zig:192-254 (read file, table, compress,
 * report), in C against libb200_fse.so, plus the decoder the reference lacks. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "b200_fse.h"

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 2;
    fseek(f, 0, SEEK_END); size_t n = (size_t)ftell(f); fseek(f, 0, SEEK_SET);
    uint8_t* in = (uint8_t*)malloc(n);
    if (fread(in, 1, n, f) != n) return 2;
    fclose(f);
    size_t freq[256];
    fse_build_frequency_table(in, n, freq);
    fse_normalize_frequency_table(freq);
    size_t sum = 0;
    for (int s = 0; s < 256; ++s) sum += freq[s];
    printf("Normalised sum: %zu\n", sum);
    TT_Entry tt[TT_SIZE];
    fse_build_transition_table(freq, tt);
    uint64_t* out = (uint64_t*)malloc(fse_compress_bound(n) * 8);
    const size_t words = fse_compress(in, n, out);
    printf("Compressed size: %zu\nCompression ratio: %f\n", words * 8, (double)n / (double)(words * 8));
    uint8_t* back = (uint8_t*)malloc(fse_decompressed_size(out, words) + 16);
    const size_t m = fse_decompress(out, words, back);
    const int ok = m == n && memcmp(in, back, n) == 0;
    printf(ok ? "SUCCESS\n" : "FAILURE\n");
    return !ok;
}
