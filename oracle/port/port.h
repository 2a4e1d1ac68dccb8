/* TEST INFRASTRUCTURE ONLY -- prototypes of the CPU restatement (oracle port). */
#pragma once
#include <stdint.h>

uint32_t port_lz77_hash(uint32_t k);
int port_lz77_compress(const uint8_t* in, uint64_t n, uint8_t* out, uint64_t* bit_index, uint32_t* F_or_null);
int port_deflate_lz77_compress(const uint8_t* in, uint64_t n, uint8_t* out, uint64_t* out_n, uint32_t* F_or_null);
int port_lz77_compress_blocks(const uint8_t* in, uint64_t n, uint64_t block, int variant,
                              uint8_t* out, uint64_t out_stride, uint64_t* sizes, int threads);
uint64_t port_lz77_decompress(const uint8_t* stream, uint64_t size, uint8_t* out);
uint64_t port_deflate_lz77_decompress(const uint8_t* tok, uint64_t ntokbytes, uint8_t* out);

void port_histogram(const uint8_t* in, uint64_t n, uint64_t* freq);
#define PORT_HUFF_MAX_SYMS 288
int port_huffman_build_n(const uint64_t* freq, int nsym, uint32_t* codes, uint8_t* lens, int* nodes_out, int* root_out);
int port_huffman_build(const uint64_t* freq, uint32_t* codes, uint8_t* lens, int* nodes_out, int* root_out);
uint64_t port_huffman_encode(const uint8_t* in, uint64_t n, const uint32_t* codes, const uint8_t* lens, uint32_t* words);
int port_huffman_compress(const uint8_t* in, uint64_t n, uint32_t* words, uint64_t* word_idx, uint64_t* bit_idx,
                          uint64_t* buffer_size, uint32_t* codes, uint8_t* lens);
uint64_t port_huffman_decompress(const uint32_t* words, uint64_t nwords, uint64_t buffer_size,
                                 const uint32_t* codes, const uint8_t* lens, uint8_t* out, uint64_t out_cap);

void port_fse_histogram(const uint8_t* in, uint64_t n, uint64_t* freq);
int port_fse_normalize(uint64_t* freq);
void port_fse_build_tables(const uint64_t* norm, uint32_t* tt, uint8_t* enc_state, uint16_t* cum);
uint64_t port_fse_encode_stream(const uint8_t* in, uint64_t n, const uint64_t* norm,
                                const uint8_t* enc_state, const uint16_t* cum, uint64_t* words);
int port_fse_decode_stream(const uint64_t* words, uint64_t total_bits, uint64_t n, const uint32_t* tt, uint8_t* out);
uint64_t port_fse_compress(const uint8_t* in, uint64_t n, uint64_t* words, uint64_t* norm_out, uint64_t* total_bits);
int port_fse_decompress(const uint64_t* words, uint64_t total_bits, uint64_t n, const uint64_t* norm, uint8_t* out);
uint64_t port_lz77_decompress_blocks(const uint8_t* stream, const uint64_t* off, uint64_t nblocks, uint64_t block,
                                     uint64_t n, int variant, uint8_t* out, int threads);

/* deflate token entropy stage (deflate_huff_port.c) */
#define PORT_DFL_NSYM 286
void port_dfl_frequencies(const uint8_t* tok, uint64_t nbytes, uint64_t* freq);
int port_dfl_build(const uint64_t* freq, uint32_t* codes, uint8_t* lens);
uint64_t port_dfl_encode(const uint8_t* tok, uint64_t nbytes, const uint32_t* codes, const uint8_t* lens, uint32_t* words);
int port_dfl_decode(const uint32_t* words, uint64_t nwords, const uint32_t* codes, const uint8_t* lens,
                    uint64_t nbytes, uint8_t* tok_out, uint64_t* bits_used);

/* Zig Huffman file format (zig_huffman_port.c) -- parity unpinned */
uint64_t port_zig_huffman_compress(const uint8_t* in, uint64_t n, uint8_t* out);
uint64_t port_zig_huffman_decompress(const uint8_t* in, uint64_t bytes, uint8_t* out, uint64_t out_cap);
