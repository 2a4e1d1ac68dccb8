/*
 * libb200comp.so -- C-ABI of the B200 (sm_100a) block-parallel codecs.
 *
 * This is the boundary the reference's four directories bind to. The reference
 * has no FFI/plugin registry: its "interface" is the set of global C functions
 * each driver links. Those exact prototypes are provided by the
 * thin per-directory shims (include/b200_lz77.h, b200_huffman.h, b200_deflate.h,
 * b200_fse.h -> libb200_{lz77,huffman,deflate,fse}.so, one per directory because
 * the reference's symbols collide). The shims are host wrappers around the
 * device-resident entry points declared here:  H2D -> *_dev -> D2H.
 *
 * Conventions
 *   - plain pointers and sizes only; `d_` = device pointer, `h_` = host pointer;
 *   - every call returns 0 on success, else a B200_ERR_* code; the text is in
 *     b200_last_error();
 *   - all kernels of a context run on that context's CUDA stream; `*_dev` calls
 *     are asynchronous unless they have an `h_` result parameter, in which case
 *     they synchronise the stream before returning;
 *   - there is NO CPU fallback: without a CUDA device every compute call fails.
 */
#ifndef B200COMP_H
#define B200COMP_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_OK 0
#define B200_ERR_CUDA 1
#define B200_ERR_ARG 2
#define B200_ERR_CAPACITY 3
#define B200_ERR_DOMAIN 4
#define B200_ERR_FORMAT 5

typedef struct b200_ctx b200_ctx;

/* ---- context / memory ------------------------------------------------------ */
int         b200_device_count(void);
/* cuda_stream: a cudaStream_t to launch on (cudaStreamLegacy / cudaStreamPerThread are
 * accepted), or NULL to create a private non-blocking stream. */
int         b200_ctx_create(b200_ctx** out, int device, void* cuda_stream);
void        b200_ctx_destroy(b200_ctx* ctx);
int         b200_ctx_sync(b200_ctx* ctx);
uint64_t    b200_ctx_launches(b200_ctx* ctx); /* kernels launched so far */
/* When enabled, every codec call brackets its dominant kernel with CUDA events on
 * the context's stream (kind: 0 LZ77 parse, 1 LZ77 decode, 2 Huffman encode,
 * 3 Huffman decode, 4 FSE encode, 5 FSE decode, 6 deflate entropy encode, 7 deflate entropy
 * decode). Up to 512 entries are kept since the
 * last b200_ctx_set_timing call; b200_ctx_timing_get waits for entry i. */
int         b200_ctx_set_timing(b200_ctx* ctx, int enable);
int         b200_ctx_timing_count(b200_ctx* ctx);
int         b200_ctx_timing_get(b200_ctx* ctx, int i, int* kind, float* ms);
const char* b200_last_error(void);
int         b200_dev_alloc(void** d_ptr, uint64_t bytes);
int         b200_dev_free(void* d_ptr);
int         b200_host_alloc(void** h_ptr, uint64_t bytes); /* pinned */
int         b200_host_free(void* h_ptr);
int         b200_copy_h2d(b200_ctx* ctx, void* d_dst, const void* h_src, uint64_t bytes); /* async */
int         b200_copy_d2h(b200_ctx* ctx, void* h_dst, const void* d_src, uint64_t bytes); /* async */
int         b200_memset(b200_ctx* ctx, void* d_dst, int value, uint64_t bytes);           /* async */

/* ---- Huffman (replaces algorithms/huffman/huffman.c:179-364) --------------- */
#define B200_HUFF_CHUNK 4096u /* symbols per encode chunk (one CTA)              */
#define B200_HUFF_SUB   256u  /* symbols per decode sub-chunk (one thread)       */

/* Layout of the caller-allocated device "side" buffer: code tables plus the index
 * that makes the stream parallel-decodable. It is NOT part of the compressed words
 * the reference compares. All offsets are in bytes. */
typedef struct {
    uint64_t bytes;          /* total size to allocate                                   */
    uint64_t nblocks;        /* table scopes: ceil(n / block_size)                       */
    uint64_t nchunks;        /* ceil over blocks of len/4096                             */
    uint64_t chunks_per_block;
    uint64_t off_freq;       /* u32[nblocks][256]  histogram (huffman.c:184-187)         */
    uint64_t off_codes;      /* u32[nblocks][256]  codes, right-aligned (huffman.c:225)  */
    uint64_t off_lens;       /* u8 [nblocks][256]  code lengths (huffman.c:226)          */
    uint64_t off_tree;       /* i16[nblocks][511][2] children; leaf = {-1, symbol}       */
    uint64_t off_meta;       /* u32[nblocks][4] = {status, distinct, root, max_len}      */
    uint64_t off_block_bits; /* u64[nblocks]   bits of each block's stream               */
    uint64_t off_block_word; /* u64[nblocks+1] first u32 word of each block; [nblocks]=total */
    uint64_t off_chunk_bits; /* u32[nchunks]   bits of each chunk                        */
    uint64_t off_chunk_off;  /* u64[nchunks+1] scratch prefix, then absolute bit offset  */
    uint64_t off_sub_off;    /* u32[nchunks*16] bit offset of each sub-chunk in its chunk*/
} b200_huff_layout;

/* block_size = table scope in bytes; 0 means one table for the whole buffer
 * (what huffman_compress does). Otherwise it must be a multiple of 4096. */
int b200_huffman_layout(uint64_t n, uint64_t block_size, b200_huff_layout* out);
/* upper bound of the stream size in u32 words for (n, block_size) */
uint64_t b200_huffman_max_words(uint64_t n, uint64_t block_size);

/* histogram -> tables -> MSB-first u32 word stream (bit-exact with
 * build_huffman_tree + gather_codes + _huffman_compress per block). meta.status
 * != 0 marks a block the reference would not encode (1: fewer than 2 distinct
 * symbols, exit(1) at huffman.c:278-281; 2: a code longer than 32 bits, which the
 * reference silently corrupts). h_total_words may be NULL (then no sync). */
int b200_huffman_encode_dev(b200_ctx* ctx, const uint8_t* d_in, uint64_t n, uint64_t block_size,
                            uint32_t* d_words, uint64_t words_capacity,
                            uint8_t* d_side, uint64_t side_bytes,
                            uint64_t* h_total_words, uint32_t* h_worst_status);
/* only the histogram + table build (build_huffman_tree + gather_codes) */
int b200_huffman_tables_dev(b200_ctx* ctx, const uint8_t* d_in, uint64_t n, uint64_t block_size,
                            uint8_t* d_side, uint64_t side_bytes);
/* _huffman_compress (huffman.c:267-285): bit packing with the caller's code table (host
 * arrays of 256 entries), one table for the whole buffer. status 3 = a symbol of the input
 * has no code (the reference prints an error and exits, huffman.c:274-277). */
int b200_huffman_encode_with_codes_dev(b200_ctx* ctx, const uint8_t* d_in, uint64_t n,
                                       const uint32_t* h_codes, const uint8_t* h_lens,
                                       uint32_t* d_words, uint64_t words_capacity, uint8_t* d_side, uint64_t side_bytes,
                                       uint64_t* h_total_words, uint32_t* h_worst_status);
/* One table over several shards (the reference builds one tree over the whole
 * input, huffman.c:184-211). Per rank: b200_huffman_histogram_dev -> the caller sums the
 * 256 x u64 bins over the ranks (all-reduce) -> b200_huffman_encode_with_freq_dev builds the
 * table from the sum (bins truncated to the reference's uint32_t) and packs this shard from
 * bit 0 of d_words (side layout = b200_huffman_layout(n, 0); an empty shard is allowed) ->
 * the caller all-gathers the shard bit counts -> b200_huffman_splice_dev ORs a shard stream
 * into a ZEROED destination at the exclusive prefix of the bit counts. The spliced words equal
 * huffman_compress on the concatenated input. Decoding stays per shard (own side index).
 * words_capacity: the table is not the shard's own, so allow the reference's 32 bits per symbol (n + 4 words). */
int b200_huffman_histogram_dev(b200_ctx* ctx, const uint8_t* d_in, uint64_t n, uint64_t* d_freq64);
int b200_huffman_encode_with_freq_dev(b200_ctx* ctx, const uint8_t* d_in, uint64_t n, const uint64_t* d_freq64,
                                      uint32_t* d_words, uint64_t words_capacity, uint8_t* d_side, uint64_t side_bytes,
                                      uint64_t* h_total_words, uint64_t* h_total_bits, uint32_t* h_worst_status);
int b200_huffman_splice_dev(b200_ctx* ctx, uint32_t* d_dst, uint64_t dst_words_capacity, uint64_t dst_bit,
                            const uint32_t* d_src, uint64_t src_bits);
/* table-lookup decoder driven by the side index written by the encoder */
int b200_huffman_decode_dev(b200_ctx* ctx, const uint32_t* d_words, uint64_t total_words,
                            const uint8_t* d_side, uint64_t side_bytes,
                            uint64_t n, uint64_t block_size, uint8_t* d_out);
/* Index-free decoder for one foreign stream (e.g. produced by the reference):
 * one GPU thread walks the stream with the reference's termination rule
 * (huffman.c:344-361) and reports the symbol count the reference would report.
 * d_codes/d_lens: 256 entries each. */
int b200_huffman_decode_serial_dev(b200_ctx* ctx, const uint32_t* d_words, uint64_t nwords,
                                   uint64_t buffer_size, const uint32_t* d_codes, const uint8_t* d_lens,
                                   uint8_t* d_out, uint64_t out_capacity, uint64_t* h_count);

/* ---- LZ77 (replaces algorithms/lz77/lz77.c:55-108,264-377 and
 *            algorithms/deflate/lz77.c:77-280) -------------------------------- */
#define B200_LZ_STANDALONE 0 /* W=2^14, MAX_LEN 15, LSB-first 9/19-bit tokens   */
#define B200_LZ_DEFLATE    1 /* W=2^15, MAX_LEN 31, 2/4-byte tokens             */

/* worst-case bytes of one block's token stream (the reference allocates 2*size) */
uint64_t b200_lz77_block_stride(uint64_t block_size);
/* Greedy parse of every block with a fresh table (the parity contract of SURVEY.md
 * §8a), tokens compacted back to back into d_out:
 *   variant 0: block b occupies bits/8+1 bytes (lz77.c:341), d_block_sizes[b] = bit_index
 *   variant 1: block b occupies d_block_sizes[b] bytes (deflate/lz77.c:277)
 * d_block_off[nblocks+1] = byte offset of each block in d_out (exclusive scan;
 * last entry = total). This is deflate.c:47-63's raw concatenation plus the index
 * the reference lacks. block_size 0 = whole buffer as one block. */
int b200_lz77_encode_dev(b200_ctx* ctx, int variant, const uint8_t* d_in, uint64_t n, uint64_t block_size,
                         uint8_t* d_out, uint64_t out_capacity,
                         uint64_t* d_block_sizes, uint64_t* d_block_off, uint64_t* h_total_bytes);
/* Test hook: as b200_lz77_encode_dev, additionally dumping the match finder's per-position
 * token candidates (0 = literal, else offset | len << 16), 65536 entries per block.
 * Blocks must be <= 65536 bytes (the shared-memory path). */
int b200_lz77_encode_debug_dev(b200_ctx* ctx, int variant, const uint8_t* d_in, uint64_t n, uint64_t block_size,
                               uint8_t* d_out, uint64_t out_capacity, uint64_t* d_block_sizes, uint64_t* d_block_off,
                               uint64_t* h_total_bytes, uint32_t* d_tok);
int b200_lz77_decode_dev(b200_ctx* ctx, int variant, const uint8_t* d_stream,
                         const uint64_t* d_block_off, const uint64_t* d_block_sizes,
                         uint64_t n, uint64_t block_size, uint8_t* d_out);

/* ---- deflate token entropy stage (completes algorithms/deflate/lz77.c:279
 *      "TODO: Build huffman tree and encode compressed buffer") -----------------
 * Per LZ block: frequencies[286] exactly as lz77_compress counts them (lz77.c:206,231,273 with
 * huffman.c:49-62: literal byte, or 256 + clz16(offset) per match), codes by the heap rule of
 * algorithms/huffman/huffman.c:100-250 over the 286 symbols, MSB-first u32 words as write_bits
 * (deflate/huffman.c:18-48). literal = code[byte]; match = code[256+k], the 15-k offset bits below
 * the leading one, the length in 5 bits. The parts the reference only declares are specified by
 * oracle/port/deflate_huff_port.c (parity unpinned there). */
#define B200_DFL_NSYM   286u  /* NUM_CODES, algorithms/deflate/huffman.h:6          */
#define B200_DFL_STRIDE 288u  /* row stride of freq / codes / lens                  */
#define B200_DFL_CHUNK  4096u /* token bytes per encode chunk (one CTA)             */
#define B200_DFL_SUB    256u  /* token bytes per decode sub-chunk (one thread)      */
typedef struct {
    uint64_t bytes;           /* total size of the side buffer                              */
    uint64_t nblocks;         /* LZ blocks = table scopes: ceil(n / block_size)             */
    uint64_t nchunks;         /* nblocks * chunks_per_block chunk slots                     */
    uint64_t chunks_per_block;/* ceil((2 * block_size + 2) / 4096): worst-case token bytes  */
    uint64_t off_freq;        /* u32[nblocks][288] frequencies[286] (lz77.c:206)            */
    uint64_t off_codes;       /* u32[nblocks][288] codes, right-aligned                     */
    uint64_t off_lens;        /* u8 [nblocks][288] code lengths                             */
    uint64_t off_tree;        /* i16[nblocks][571][2] children; leaf = {-1, symbol}         */
    uint64_t off_meta;        /* u32[nblocks][4] = {status, distinct, root, max_len}; status 1 = no tokens, 2 = a code > 32 bits */
    uint64_t off_tok_off;     /* u64[nblocks+1] byte offset of each block's tokens          */
    uint64_t off_tok_sizes;   /* u64[nblocks]   token bytes of each block                   */
    uint64_t off_block_bits;  /* u64[nblocks]   bits of each block's stream                 */
    uint64_t off_block_word;  /* u64[nblocks+1] first u32 word of each block; [nblocks]=total */
    uint64_t off_chunk_state; /* u8 [nchunks]   1 = the chunk's first 2-byte unit is the tail of a match */
    uint64_t off_chunk_bits;  /* u32[nchunks]   bits of each chunk                          */
    uint64_t off_chunk_off;   /* u64[nchunks+1] absolute bit offset of each chunk           */
    uint64_t off_sub_off;     /* u32[nchunks*16] bit offset of each sub-chunk in its chunk; bit 31 = starts with a tail unit */
} b200_dfl_layout;
/* n / block_size: the UNCOMPRESSED size and LZ block size the tokens came from (0 = one block) */
int b200_dfl_layout_for(uint64_t n, uint64_t block_size, b200_dfl_layout* out);
uint64_t b200_dfl_max_words(uint64_t n, uint64_t block_size);
/* byte tokens (b200_lz77_encode_dev variant 1: d_out, d_block_off, d_block_sizes) -> packed words.
 * d_tokens must be 16-byte aligned with tokens_capacity bytes allocated. */
int b200_dfl_encode_dev(b200_ctx* ctx, const uint8_t* d_tokens, uint64_t tokens_capacity,
                        const uint64_t* d_tok_off, const uint64_t* d_tok_sizes, uint64_t n, uint64_t block_size,
                        uint32_t* d_words, uint64_t words_capacity, uint8_t* d_side, uint64_t side_bytes,
                        uint64_t* h_total_words, uint32_t* h_worst_status);
/* packed words -> the byte tokens, placed at the offsets recorded in the side buffer */
int b200_dfl_decode_dev(b200_ctx* ctx, const uint32_t* d_words, uint64_t total_words, const uint8_t* d_side,
                        uint64_t side_bytes, uint64_t n, uint64_t block_size, uint8_t* d_tokens_out);
/* both stages: lz77_compress per block + the entropy stage / its inverse + the LZ77 decoder.
 * d_tokens is a scratch of b200_lz77_max_bytes(1, n, block_size) bytes. */
int b200_deflate_compress_dev(b200_ctx* ctx, const uint8_t* d_in, uint64_t n, uint64_t block_size,
                              uint8_t* d_tokens, uint64_t tokens_capacity, uint64_t* d_tok_sizes, uint64_t* d_tok_off,
                              uint32_t* d_words, uint64_t words_capacity, uint8_t* d_side, uint64_t side_bytes,
                              uint64_t* h_total_words, uint32_t* h_worst_status);
int b200_deflate_decompress_dev(b200_ctx* ctx, const uint32_t* d_words, uint64_t total_words, const uint8_t* d_side,
                                uint64_t side_bytes, uint64_t n, uint64_t block_size, uint8_t* d_tokens, uint8_t* d_out);

/* ---- FSE (C mirror of algorithms/fse/src/main.zig:50-189) ------------------- */
#define B200_FSE_TABLE_LOG 8u
typedef struct {
    uint64_t bytes;
    uint64_t nblocks;       /* table scopes                                             */
    uint64_t nsegs;         /* independent single-state streams                         */
    uint64_t segs_per_block;
    uint64_t off_freq;      /* u32[nblocks][256] histogram (main.zig:88-96)             */
    uint64_t off_norm;      /* u16[nblocks][256] normalised counts (main.zig:106-149)   */
    uint64_t off_tt;        /* u32[nblocks][256] TT_Entry {symbol,next_state,num_bits}  */
    uint64_t off_seg_bits;  /* u32[nsegs] exact bits of each segment stream             */
    uint64_t off_seg_word;  /* u64[nsegs+1] first u64 word of each segment; last=total  */
} b200_fse_layout;
int b200_fse_layout_for(uint64_t n, uint64_t block_size, uint64_t seg_size, b200_fse_layout* out);
uint64_t b200_fse_max_words(uint64_t n, uint64_t seg_size);
int b200_fse_encode_dev(b200_ctx* ctx, const uint8_t* d_in, uint64_t n, uint64_t block_size, uint64_t seg_size,
                        uint64_t* d_words, uint64_t words_capacity, uint8_t* d_side, uint64_t side_bytes,
                        uint64_t* h_total_words);
int b200_fse_decode_dev(b200_ctx* ctx, const uint64_t* d_words, const uint8_t* d_side, uint64_t side_bytes,
                        uint64_t n, uint64_t block_size, uint64_t seg_size, uint8_t* d_out, uint32_t* h_bad_segments);
/* histogram + normalisation only (the part of FSE whose parity is pinned) */
int b200_fse_normalize_dev(b200_ctx* ctx, const uint8_t* d_in, uint64_t n, uint64_t block_size,
                           uint8_t* d_side, uint64_t side_bytes);
/* decoder side of a stored stream: with norm[] and seg_bits[] already in d_side, rebuild
 * the transition tables and the segment word offsets */
int b200_fse_rebuild_index_dev(b200_ctx* ctx, uint64_t n, uint64_t block_size, uint64_t seg_size,
                               uint8_t* d_side, uint64_t side_bytes);

/* ---- host-buffer wrappers: H2D -> *_dev -> D2H, synchronous ------------------
 * These are what the reference-named shims call; bench.py's "e2e" number is
 * measured through them with the copies inside the timed region. Host buffers may
 * be pageable; pinned ones (b200_host_alloc) copy faster. */
uint64_t b200_lz77_max_bytes(int variant, uint64_t n, uint64_t block_size);
int b200_lz77_compress_host(b200_ctx* ctx, int variant, const uint8_t* h_in, uint64_t n, uint64_t block_size,
                            uint8_t* h_out, uint64_t out_capacity, uint64_t* h_block_sizes,
                            uint64_t* h_block_off, uint64_t* h_total_bytes);
int b200_lz77_decompress_host(b200_ctx* ctx, int variant, const uint8_t* h_stream, uint64_t stream_bytes,
                              const uint64_t* h_block_off, const uint64_t* h_block_sizes,
                              uint64_t n, uint64_t block_size, uint8_t* h_out);
int b200_huffman_compress_host(b200_ctx* ctx, const uint8_t* h_in, uint64_t n, uint64_t block_size,
                               uint32_t* h_words, uint64_t words_capacity, uint8_t* h_side, uint64_t side_bytes,
                               uint64_t* h_total_words, uint32_t* h_worst_status);
int b200_huffman_decompress_host(b200_ctx* ctx, const uint32_t* h_words, uint64_t total_words,
                                 const uint8_t* h_side, uint64_t side_bytes, uint64_t n, uint64_t block_size,
                                 uint8_t* h_out);
int b200_huffman_decompress_serial_host(b200_ctx* ctx, const uint32_t* h_words, uint64_t nwords,
                                        uint64_t buffer_size, const uint32_t* h_codes, const uint8_t* h_lens,
                                        uint8_t* h_out, uint64_t out_capacity, uint64_t* h_count);

/* deflate with the entropy stage (lz77_compress per block + the Huffman-coded token stream);
 * h_side receives b200_dfl_layout.bytes of tables and decode index */
int b200_deflate_compress_host(b200_ctx* ctx, const uint8_t* h_in, uint64_t n, uint64_t block_size,
                               uint32_t* h_words, uint64_t words_capacity, uint8_t* h_side, uint64_t side_bytes,
                               uint64_t* h_total_words, uint32_t* h_worst_status);
int b200_deflate_decompress_host(b200_ctx* ctx, const uint32_t* h_words, uint64_t total_words,
                                 const uint8_t* h_side, uint64_t side_bytes, uint64_t n, uint64_t block_size,
                                 uint8_t* h_out);

int b200_huffman_tables_host(b200_ctx* ctx, const uint8_t* h_in, uint64_t n, uint64_t block_size,
                             uint8_t* h_side, uint64_t side_bytes);
int b200_huffman_compress_codes_host(b200_ctx* ctx, const uint8_t* h_in, uint64_t n,
                                     const uint32_t* h_codes, const uint8_t* h_lens,
                                     uint32_t* h_words, uint64_t words_capacity, uint8_t* h_side, uint64_t side_bytes,
                                     uint64_t* h_total_words, uint32_t* h_worst_status);
/* FSE container (self-describing, what fse_compress of the C mirror returns):
 *   u64 header[8] = {magic "B00FSE01", n, block_size, seg_size, nblocks, nsegs, stream_words, 0}
 *   u16 norm[nblocks][256] | u32 seg_bits[nsegs] | u64 stream[stream_words], each part padded to a u64 */
uint64_t b200_fse_container_max_words(uint64_t n, uint64_t block_size, uint64_t seg_size);
int b200_fse_compress_host(b200_ctx* ctx, const uint8_t* h_in, uint64_t n, uint64_t block_size, uint64_t seg_size,
                           uint64_t* h_out, uint64_t out_capacity_words, uint64_t* h_total_words);
int b200_fse_container_size(const uint64_t* h_container, uint64_t words, uint64_t* h_n);
int b200_fse_decompress_host(b200_ctx* ctx, const uint64_t* h_container, uint64_t words,
                             uint8_t* h_out, uint64_t out_capacity, uint64_t* h_n);
int b200_fse_normalize_host(b200_ctx* ctx, const uint8_t* h_in, uint64_t n, uint32_t* h_freq, uint16_t* h_norm);
/* one table: normalise 256 raw counts and build the transition table (h_freq given), or
 * build the table of already normalised counts (h_norm_in given) */
int b200_fse_tables_host(b200_ctx* ctx, const uint32_t* h_freq, const uint16_t* h_norm_in,
                         uint16_t* h_norm_out, uint32_t* h_tt_out);

/* ---- self-describing containers (csrc/container.cu documents the layout) ----------
 * ONE buffer = header {magic "B200CONT", version, codec, n, block_size, ...} + serialized code tables (the
 * histogram of every table scope: the decoder replays the reference's heap on it) + per-chunk bit counts and
 * the parallel-decode index + the payload words (bit-exact with the non-container calls). A decoder needs
 * nothing else; model: the Zig Huffman's tree dump + size header, zig_huffman/src/main.zig:11-18,155-200,513-530.
 * codec 1 = Huffman (algorithms/huffman), codec 2 = deflate with the entropy stage (algorithms/deflate). */
#define B200_CODEC_HUFFMAN 1u
#define B200_CODEC_DEFLATE 2u
uint64_t b200_huffman_container_max_bytes(uint64_t n, uint64_t block_size);
uint64_t b200_deflate_container_max_bytes(uint64_t n, uint64_t block_size);
int b200_container_info(const void* h_container, uint64_t bytes, uint32_t* codec, uint64_t* n, uint64_t* block_size);
int b200_huffman_compress_container_host(b200_ctx* ctx, const uint8_t* h_in, uint64_t n, uint64_t block_size,
                                         void* h_out, uint64_t out_capacity, uint64_t* h_total_bytes);
int b200_huffman_decompress_container_host(b200_ctx* ctx, const void* h_container, uint64_t bytes,
                                           uint8_t* h_out, uint64_t out_capacity, uint64_t* h_n);
int b200_deflate_compress_container_host(b200_ctx* ctx, const uint8_t* h_in, uint64_t n, uint64_t block_size,
                                         void* h_out, uint64_t out_capacity, uint64_t* h_total_bytes);
int b200_deflate_decompress_container_host(b200_ctx* ctx, const void* h_container, uint64_t bytes,
                                           uint8_t* h_out, uint64_t out_capacity, uint64_t* h_n);
/* ---- multi-GPU host entry points in C (csrc/multi.cu) -------------------------------------
 * ONE buffer sharded by block over the GPUs of the box: contiguous block ranges, ceil(nblocks / G) per device, one
 * host thread per GPU, no data-path collective. The one exchange is an ncclAllGather of the G shard sizes (NCCL is
 * bound at run time from libnccl.so.2; a handle with one device needs no NCCL), after which every device copies its
 * shard to its global offset of h_out: the result is byte for byte what the single-device call returns. */
typedef struct b200_multi b200_multi;
int      b200_multi_create(b200_multi** out, const int* devices /* NULL = 0..ndev-1 */, int ndev);
void     b200_multi_destroy(b200_multi* m);
int      b200_multi_device_count(b200_multi* m);
uint64_t b200_multi_allgathers(b200_multi* m);   /* size exchanges issued so far */
uint64_t b200_multi_launches(b200_multi* m);     /* kernels launched on all devices */
int b200_lz77_compress_multi_host(b200_multi* m, int variant, const uint8_t* h_in, uint64_t n, uint64_t block_size,
                                  uint8_t* h_out, uint64_t out_capacity, uint64_t* h_block_sizes, uint64_t* h_block_off,
                                  uint64_t* h_total_bytes);
int b200_lz77_decompress_multi_host(b200_multi* m, int variant, const uint8_t* h_stream, uint64_t stream_bytes,
                                    const uint64_t* h_block_off, const uint64_t* h_block_sizes, uint64_t n,
                                    uint64_t block_size, uint8_t* h_out);

/* ---- Zig-Huffman-compatible chunked mode: the FILE FORMAT of
 * algorithms/huffman/zig_huffman/src/main.zig -- 4 MiB chunks (:5), per chunk the tree dumped pre-order (value u8 +
 * freq u32 per node, i32 -1 for a missing child, :155-176), CompressedSize{last_block:1, value:31} (:11-18,513-520)
 * and the codes packed MSB-first into bytes (:316-338), whole bytes only (the format loses the tail bits of a chunk,
 * :523; reproduced as it is). The tree comes from std.PriorityQueue over the histogram of the whole read buffer
 * (:100-153). PARITY UNPINNED (no Zig toolchain here): oracle/port/zig_huffman_port.c is the written spec. */
uint64_t b200_zig_huffman_max_bytes(uint64_t n);
int b200_zig_huffman_compress_host(b200_ctx* ctx, const uint8_t* h_in, uint64_t n, uint8_t* h_out, uint64_t out_capacity,
                                   uint64_t* h_total_bytes);
/* h_out needs 4 MiB per chunk at most; *h_n = bytes the reference's decoder would write */
int b200_zig_huffman_decompress_host(b200_ctx* ctx, const uint8_t* h_in, uint64_t bytes, uint8_t* h_out, uint64_t out_capacity,
                                     uint64_t* h_n);

/* decoder side: rebuild codes / lengths / trees from the histograms already in d_side */
int b200_huffman_tables_from_freq_dev(b200_ctx* ctx, uint8_t* d_side, uint64_t side_bytes, uint64_t n, uint64_t block_size);
int b200_dfl_tables_from_freq_dev(b200_ctx* ctx, uint8_t* d_side, uint64_t side_bytes, uint64_t n, uint64_t block_size);

#ifdef __cplusplus
}
#endif
#endif /* B200COMP_H */
