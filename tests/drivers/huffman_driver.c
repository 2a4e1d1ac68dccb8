/* Shaped like /root/reference/algorithms/huffman/main.c:34-103, linked against
 * libb200_huffman.so: the file name comes from argv instead of being hard-coded. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "b200_huffman.h"

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    uint64_t filesize;
    char* buffer = read_input_buffer(argv[1], &filesize);
    printf("File size uncompressed: %d\n\n", (int)filesize);
    clock_t start = clock();
    BitWriter bit_writer;
    Node root = huffman_compress(buffer, filesize, &bit_writer);
    printf("Compression MB/s: %f\n", (double)filesize / (double)(clock() - start) * CLOCKS_PER_SEC / (1024.0 * 1024.0));
    printf("Compressed size: %lu\n\n", bit_writer.buffer_size);
    char* decompressed = (char*)malloc(filesize + 64);
    uint64_t decompressed_bytes = filesize + 64;
    start = clock();
    huffman_decompress(&bit_writer, &root, decompressed, &decompressed_bytes);
    printf("Decompressed size: %lu\n", decompressed_bytes);
    printf("Decompression MB/s: %f\n\n", (double)filesize / (double)(clock() - start) * CLOCKS_PER_SEC / (1024.0 * 1024.0));
    uint64_t mismatches = 0;
    for (uint64_t i = 0; i < filesize; ++i) mismatches += buffer[i] != decompressed[i];
    printf("Number of mismatches: %d\n", (int)mismatches);
    printf(mismatches == 0 ? "SUCCESS\n" : "FAILURE\n");
    if (argc > 2) {   /* dump the words for the parity check */
        FILE* f = fopen(argv[2], "wb");
        const uint64_t nwords = bit_writer.word_idx + (bit_writer.bit_idx > 0);
        fwrite(bit_writer.buffer, 4, nwords, f);
        fclose(f);
    }
    return mismatches != 0;
}
