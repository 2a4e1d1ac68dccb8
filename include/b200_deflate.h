/*
 * libb200_deflate.so -- drop-in for the functions /root/reference/algorithms/deflate/main.c
 * links from algorithms/deflate/{deflate,lz77,huffman}.c. Names, argument meaning, file
 * naming and error behaviour follow algorithms/deflate/deflate.h:8-30, lz77.h:5-59 and
 * huffman.h:6-92. The match finder + greedy parse + byte-token emission of every 64 KiB
 * block (lz77.c:199-280) and the raw concatenation of the blocks (deflate.c:47-63) run on
 * the GPU through libb200comp.so. No CPU fallback: without a CUDA device the calls print
 * the error to stderr and exit(1) like deflate.c:24-34.
 *
 * Differences a maintainer must know (SURVEY.md §8a "Parity contract for D-rows"):
 *  - every block is parsed against a FRESH table. The reference's compress() shares one
 *    table between blocks, which serialises them and only shadows matches (entries of the
 *    previous block are rejected by lz77.c:223); lz77_compress() called with a freshly
 *    initialised table is reproduced byte for byte. The table argument is not updated.
 *  - compress() additionally writes "<name>.deflate.idx" (u64 magic, n, block size, block
 *    count, then block count + 1 byte offsets) so that decompress() can decode the blocks
 *    in parallel; without it the offsets are recovered by scanning the token flags.
 *  - lz77_decompress() and decompress() are working decoders; the reference's are broken /
 *    empty (lz77.c:282-311, deflate.c:78-79). decompress(state, path) decodes the file
 *    `path` (default: state->compressed_filename) into "<path>.out".
 *  - push_heap / pop_heap / new_node / build_huffman_tree are declared by the reference
 *    (huffman.h:16-32,84) but defined nowhere; they are not provided here either.
 */
#ifndef B200_DEFLATE_H
#define B200_DEFLATE_H
#include <stdint.h>
#include <stdbool.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- lz77.h:5-59 ---------------------------------------------------------------------- */
#define MAX_WINDOW_BITS 15
#define WINDOW_SIZE (1 << MAX_WINDOW_BITS)
#define MAX_LENGTH_BITS 5
#define TABLE_SIZE (1 << (MAX_WINDOW_BITS + 5))

typedef struct ArrayNode { uint32_t pattern; uint64_t index; bool is_set; } ArrayNode;
typedef struct Buckets { uint32_t* patterns; uint64_t* indices; bool* is_set; } Buckets;
typedef struct {
    Buckets  buckets;
    uint32_t bucket_indices[1 << MAX_WINDOW_BITS];
    uint32_t current_idx;
    bool     is_full;
} HashTableArray;

uint64_t min(uint64_t a, uint64_t b);                                              /* lz77.c:10 */
uint64_t max(uint64_t a, uint64_t b);                                              /* lz77.c:11 */
uint32_t hash(uint32_t pattern);                                                   /* lz77.c:14-42   */
void     init_hash_table(HashTableArray* table);                                   /* lz77.c:44-66 (is_set/indices zeroed too) */
void     insert_hash_table(HashTableArray* table, uint32_t pattern, uint64_t index); /* lz77.c:77-145 */
uint64_t find(HashTableArray* table, uint32_t pattern);                            /* lz77.c:147-174 */
void     write_literal(char* buffer, char c, uint64_t* buffer_index);              /* lz77.c:176-184 */
void     write_length_distance(char* buffer, uint8_t length, uint16_t distance,
                               uint64_t* buffer_index);                            /* lz77.c:186-197 */
/* lz77.c:199-280, GPU. compressed_buffer must hold 2 * input_buffer_size bytes. */
void     lz77_compress(const char* input_buffer, uint64_t input_buffer_size, char* compressed_buffer,
                       uint64_t* compressed_buffer_size, HashTableArray* table);
/* GPU decoder of one block's byte tokens. *decompressed_buffer_size: in = capacity, out = bytes written. */
void     lz77_decompress(const char* compressed_buffer, uint64_t compressed_buffer_size,
                         char* decompressed_buffer, uint64_t* decompressed_buffer_size);

/* ---- huffman.h:6-92 (deflate's copy) --------------------------------------------------- */
#define NUM_CODES 286
typedef struct MinHeapNode MinHeapNode;
struct MinHeapNode { uint8_t data; uint32_t frequency; MinHeapNode* left; MinHeapNode* right; };
typedef struct { uint32_t* buffer; uint64_t bit_idx; uint64_t word_idx; uint64_t buffer_size; } BitWriter;
void init_bitwriter(BitWriter* writer, uint64_t buffer_size);                      /* huffman.c:7-13  */
void write_bits(BitWriter* writer, uint32_t bits, uint8_t length);                 /* huffman.c:16-46 */
void append_huffman_tree_literal(uint32_t* frequencies, char literal);             /* huffman.c:49-54 */
void append_huffman_tree_pair(uint32_t* frequencies, uint16_t offset);             /* huffman.c:56-62 */
void gather_codes(MinHeapNode* root, uint16_t code, uint8_t length,
                  uint16_t* codes, uint8_t* code_lengths);                         /* huffman.c:64-97 */

/* ---- deflate.h:8-30 -------------------------------------------------------------------- */
#define BUFFER_SIZE 65536
typedef struct HuffmanNode {
    struct HuffmanNode* left;
    struct HuffmanNode* right;
    uint16_t value;
    uint64_t frequency;
} HuffmanNode;
void init_huffman_node(HuffmanNode* node);                                         /* deflate.c:81-86  */
void destroy_huffman_node(HuffmanNode* node);                                      /* deflate.c:88-98  */
bool compare_huffman_node(const HuffmanNode* a, const HuffmanNode* b);             /* deflate.c:100-102 */

typedef struct StateData {
    HashTableArray* table;
    HuffmanNode*    huffman_root;
    char*           compressed_filename;
} StateData;
StateData compress(const char* input_filename);                                    /* deflate.c:10-76, GPU */
void      decompress(StateData* state_data, const char* input_filename);           /* deflate.c:78-79, GPU */

/* Extension (not in the reference): the in-memory form of compress(): all blocks of
 * `size` bytes at once, tokens concatenated into `out` (capacity 2*size + 64), block
 * byte offsets into block_off[nblocks + 1]. Returns the token bytes. */
uint64_t deflate_compress_buffer(const char* in, uint64_t size, uint64_t block_size, char* out, uint64_t* block_off);
void     deflate_decompress_buffer(const char* tokens, uint64_t token_bytes, const uint64_t* block_off,
                                   uint64_t size, uint64_t block_size, char* out);

/* Extension: the pipeline with the entropy stage the reference leaves as a TODO (deflate/lz77.c:279), file to
 * file through ONE self-describing container (b200comp.h "containers"). Returns the bytes written. */
uint64_t compress_entropy(const char* input_filename, const char* output_filename);
uint64_t decompress_entropy(const char* input_filename, const char* output_filename);

#ifdef __cplusplus
}
#endif
#endif
