/* Shaped like /root/reference/algorithms/lz77/main.c:10-69, linked against libb200_lz77.so. */
#include <stdio.h>
#include <stdlib.h>
#include <time.h>
#include "b200_lz77.h"

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    uint64_t filesize;
    char* buffer = read_input_buffer(argv[1], &filesize);
    clock_t start = clock();
    BitStream* stream = lz77_compress(buffer, filesize);
    printf("Compression MB/s: %f\n", (double)filesize / (double)(clock() - start) * CLOCKS_PER_SEC / (1024.0 * 1024.0));
    const uint64_t compressed_bits = stream->bit_index;
    if (argc > 2) {
        FILE* f = fopen(argv[2], "wb");
        fwrite(stream->data, 1, compressed_bits / 8 + 1, f);
        fclose(f);
    }
    start = clock();
    uint64_t decompressed_size;
    char* decompressed = lz77_decompress(stream, filesize, &decompressed_size);
    printf("Decompression MB/s: %f\n", (double)filesize / (double)(clock() - start) * CLOCKS_PER_SEC / (1024.0 * 1024.0));
    const bool ok = check_buffer_equivalence(buffer, decompressed, min(filesize, decompressed_size));
    printf(ok ? "SUCCESS\n" : "FAILURE\n");
    printf("Uncompressed size: %lu\nCompressed bits:   %lu\nCompression ratio: %f\n", filesize, compressed_bits,
           (double)filesize * 8 / (double)compressed_bits);
    free(stream->data); free(stream); free(buffer); free(decompressed);
    return !ok;
}
