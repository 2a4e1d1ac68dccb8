/*
 * libb200_huffman.so -- drop-in for the functions /root/reference/algorithms/huffman/main.c
 * links from algorithms/huffman/huffman.c. Same names, argument meaning, ownership and
 * error behaviour as algorithms/huffman/huffman.h:42-113. The three hot loops (histogram
 * huffman.c:184-187, bit packing :267-285, decode :330-364) and the heap-exact table build
 * (:189-250) run on the GPU through libb200comp.so (include/b200comp.h). There is no CPU
 * fallback: without a CUDA device the calls print "ERROR: ..." and exit(1), the
 * reference's own failure mode (huffman.c:137-140,149-152,278-281).
 *
 * Notes for a maintainer switching over:
 *  - huffman_compress returns the root by value with malloc'd children, exactly like
 *    huffman.c:327; the tree is rebuilt on the host from the GPU's node array, so
 *    gather_codes on it gives the reference's codes[] / code_lengths[] bit for bit.
 *  - writer->buffer holds whole u32 words (the reference shrinks it to buffer_size bytes
 *    and then reads the last word out of bounds); buffer_size, word_idx and
 *    bit_idx carry the reference's values (huffman.c:318-320).
 *  - huffman_decompress reproduces the reference's termination rule and symbol count
 *    (n + the symbols decoded out of the pad bits) but never writes past
 *    the capacity passed in *output_size.
 *  - a stream produced by this library in the same process is decoded by the parallel
 *    table-lookup decoder (the library keeps the bit-offset index of every stream it produced,
 *    keyed by words pointer + bit count, mutex protected); any other stream is decoded by one GPU
 *    thread walking it like the reference. Streams that leave the process use the container calls.
 */
#ifndef B200_HUFFMAN_H
#define B200_HUFFMAN_H
#include <stdint.h>
#include <stdbool.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {                /* huffman.h:42-47 */
    uint32_t* buffer;
    uint64_t  bit_idx;
    uint64_t  word_idx;
    uint64_t  buffer_size;
} BitWriter;

typedef struct Node Node;
struct Node {                   /* huffman.h:53-59 */
    uint8_t  value;
    uint32_t frequency;
    Node*    left;
    Node*    right;
};

typedef struct PriorityQueue {  /* huffman.h:61-65 */
    Node**   nodes;
    uint64_t size;
    uint64_t capacity;
} PriorityQueue;

void init_bitwriter(BitWriter* writer, uint64_t buffer_size);                 /* huffman.c:9-15  */
void write_bits(BitWriter* writer, uint32_t bits, uint8_t length);            /* huffman.c:18-48 */

PriorityQueue* init_priority_queue(uint64_t capacity);                        /* huffman.c:80-89   */
void  swap_nodes(Node** a, Node** b);                                         /* huffman.c:91-98   */
void  heapify_up(PriorityQueue* queue, uint64_t idx);                         /* huffman.c:100-110 */
void  heapify_down(PriorityQueue* queue, uint64_t idx);                       /* huffman.c:112-131 */
void  enqueue(PriorityQueue* queue, Node* node);                              /* huffman.c:133-144 */
Node* dequeue(PriorityQueue* queue);                                          /* huffman.c:146-159 */
bool  is_empty(PriorityQueue* queue);                                         /* huffman.c:161-163 */
Node* init_node(uint8_t value, uint32_t frequency);                           /* huffman.c:165-176 */

void  print_bit_string(uint8_t* buffer, uint64_t size);                       /* huffman.c:50-59   */
char* read_input_buffer(const char* filename, uint64_t* size);                /* huffman.c:61-78   */
void  build_huffman_tree(char* buffer, uint64_t size, Node** root);           /* huffman.c:179-215, GPU */
void  gather_codes(Node* root, uint32_t code, uint32_t length,
                   uint32_t* codes, uint8_t* code_lengths);                   /* huffman.c:217-250 */
void  print_codes(uint32_t* codes, uint8_t* code_lengths);                    /* huffman.c:252-265 */
void  _huffman_compress(char* buffer, uint64_t size, uint32_t* codes,
                        uint8_t* code_lengths, BitWriter* writer);            /* huffman.c:267-285, GPU */
Node  huffman_compress(char* buffer, uint64_t size, BitWriter* writer);       /* huffman.c:288-328, GPU */
void  huffman_decompress(BitWriter* writer, Node* root, char* output,
                         uint64_t* output_size);                              /* huffman.c:330-364, GPU */
/* the reference's version (huffman.c:366-401) is unfinished; this one is a working
 * table-lookup decoder with the contract of huffman_decompress */
void  huffman_decompress_lookup_table(BitWriter* writer, Node* root, char* output,
                                      uint64_t* output_size);

/* Extension (not in the reference, which never writes its stream anywhere): file in, self-describing container out
 * (b200comp.h "containers": tables + decode index + words in ONE stream), and back, in any process.
 * block_size 0 = one table for the whole file (what huffman_compress does), else a multiple of 4096. */
uint64_t huffman_compress_file(const char* input_filename, const char* output_filename, uint64_t block_size);
uint64_t huffman_decompress_file(const char* input_filename, const char* output_filename);

#ifdef __cplusplus
}
#endif
#endif
