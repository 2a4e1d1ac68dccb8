/*
 * libb200_fse.so -- C mirror of /root/reference/algorithms/fse/src/main.zig. The reference is
 * Zig, exports nothing (src/root.zig:4 is the `zig init` template) and does not compile
 * (main.zig:47), so there is no binding to replace; this header keeps the Zig file's function
 * split and constants so a C driver shaped like main.zig:192-254 can call them:
 *   buildFrequencyTable      main.zig:88-96    -> fse_build_frequency_table      (GPU, pinned)
 *   normalizeFrequencyTable  main.zig:106-149  -> fse_normalize_frequency_table  (GPU, pinned)
 *   buildTransitionTable     main.zig:151-189  -> fse_build_transition_table     (defective in
 *                            the reference; here a valid tANS table, parity unpinned)
 *   compress                 main.zig:50-68    -> fse_compress                   (GPU)
 *   (no decoder in the reference)              -> fse_decompress                 (GPU)
 * Kept from the reference: TABLE_LOG 8, 256 states, TT_Entry {symbol u8, next_state u16,
 * num_bits u8} packed in a u32, single-state streams starting at state 0, the last byte
 * stored raw in the first 8 bits, symbols visited from the end, the final state flushed in
 * 8 bits, LSB-first u64 words. No CPU fallback: errors print to stderr and exit(1).
 */
#ifndef B200_FSE_H
#define B200_FSE_H
#include <stdint.h>
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

#define TABLE_LOG 8                 /* main.zig:80 */
#define TT_SIZE (1 << TABLE_LOG)    /* main.zig:81 */
#define FSE_BLOCK_SIZE 65536        /* table scope of fse_compress */
#define FSE_SEGMENT_SIZE 1024       /* one single-state stream per segment */

typedef uint32_t TT_Entry;          /* main.zig:73-78: symbol | next_state << 8 | num_bits << 24 */

/* histogram of the whole buffer into freq[256] (main.zig:88-96) */
void   fse_build_frequency_table(const uint8_t* input, size_t n, size_t freq[256]);
/* in-place normalisation of freq[256] to a sum of 256 (main.zig:106-149): f64 scale,
 * truncation, at least 1 per present symbol, the remainder to the first maximum */
void   fse_normalize_frequency_table(size_t freq[256]);
/* decode-style table of 256 states from normalised counts */
void   fse_build_transition_table(const size_t norm[256], TT_Entry tt[TT_SIZE]);
/* words a compressed stream of n input bytes may need (output: []u64 of main.zig:233-236
 * only holds n/8 words, which incompressible input overflows) */
size_t fse_compress_bound(size_t n);
/* returns the number of u64 words written to `output` (a self-describing container, see
 * include/b200comp.h "FSE container") */
size_t fse_compress(const uint8_t* input, size_t n, uint64_t* output);
/* size of the original data of a container, or 0 if it is not one */
size_t fse_decompressed_size(const uint64_t* compressed, size_t words);
/* returns the number of bytes written to `output` (capacity >= fse_decompressed_size) */
size_t fse_decompress(const uint64_t* compressed, size_t words, uint8_t* output);

#ifdef __cplusplus
}
#endif
#endif
