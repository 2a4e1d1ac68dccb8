/*
 * libb200_lz77.so -- drop-in for the functions /root/reference/algorithms/lz77/main.c
 * links from algorithms/lz77/lz77.c. Same names, argument meaning, ownership and error
 * behaviour as algorithms/lz77/lz77.h:10-63; the work runs on the GPU through
 * libb200comp.so (include/b200comp.h). No CPU fallback: without a CUDA device the
 * calls print the error and exit(1), the reference's own failure mode (lz77.c:315-326).
 *
 * The reference call is unblocked: one table sliding over the whole buffer, which is one
 * sequential stream. The drop-in reproduces exactly that (the buffer is
 * one block); throughput comes from the blocked entry points of b200comp.h.
 */
#ifndef B200_LZ77_H
#define B200_LZ77_H
#include <stdint.h>
#include <stdbool.h>
#ifdef __cplusplus
extern "C" {
#endif

#define LENGTH_BITS 4   /* lz77.h:6 */
#define WINDOW_BITS 14  /* lz77.h:7 */

typedef struct {        /* lz77.h:14-17 */
    uint8_t* data;
    uint64_t bit_index;
} BitStream;

#define TABLE_SIZE (1 << (WINDOW_BITS + 6))   /* lz77.h:8 */

typedef struct ArrayNode {   /* lz77.h:19-23 */
    uint32_t pattern;
    uint64_t index;
    bool is_set;
} ArrayNode;

typedef struct {             /* lz77.h:25-30 */
    ArrayNode* buckets;
    uint32_t bucket_indices[1 << WINDOW_BITS];
    uint32_t current_idx;
    bool is_full;
} HashTableArray;

/* host-side table helpers of lz77.h:33-37 (the drivers do not call them; exported so that every
 * public name of the header resolves). They run on the host: one table, one caller. */
void     init_hash_table(HashTableArray* table);                                    /* lz77.c:43-53  */
void     insert_hash_table(HashTableArray* table, uint32_t pattern, uint64_t index); /* lz77.c:55-86  */
uint64_t find(HashTableArray* table, uint32_t pattern);                             /* lz77.c:94-108 */
void     print_bit_string(const char* buffer, uint64_t size);                       /* lz77.c:111-119 */

uint64_t min(uint64_t a, uint64_t b);                      /* lz77.c:9  */
uint64_t max(uint64_t a, uint64_t b);                      /* lz77.c:10 */
uint32_t hash(uint32_t pattern);                           /* lz77.c:13-41 */
char*    read_input_buffer(const char* filename, uint64_t* size);   /* lz77.c:121-137 */
void     init_bitstream(BitStream* stream, uint8_t* buffer);        /* lz77.c:139-142 */
void     write_bit(BitStream* stream, bool bit);                    /* lz77.c:144-156 */
bool     read_bit(BitStream* stream);                               /* lz77.c:159-167 */
void     write_bits(BitStream* stream, uint64_t value, uint64_t num_bits);  /* lz77.c:169-174 */
uint64_t read_bits(BitStream* stream, uint64_t num_bits);                   /* lz77.c:176-184 */
bool     check_buffer_equivalence(const char* buffer1, const char* buffer2, uint64_t size); /* lz77.c:379-392 */
/* lz77.c:264-345: returns a malloc'd BitStream with malloc'd data of bit_index/8+1 bytes */
BitStream* lz77_compress(const char* buffer, uint64_t size);
/* lz77.c:347-377: returns malloc(size [+ slack for a final match]); resets stream->bit_index */
char*    lz77_decompress(BitStream* compressed_stream, uint64_t size, uint64_t* decompressed_size);

#ifdef __cplusplus
}
#endif
#endif
