/*
 * TEST INFRASTRUCTURE ONLY (oracle/_build/liboracle_port.so).
 *
 * FSE. The reference (/root/reference/algorithms/fse/src/main.zig) is Zig that
 * does not compile (syntax error at main.zig:47) and there is no Zig toolchain,
 * so nothing here can be checked against an executed reference.
 *
 *   PARITY PINNED (restatement of well-defined code, hand-derived KAT in
 *   SURVEY.md §4.3):
 *     port_fse_histogram     main.zig:88-96   buildFrequencyTable
 *     port_fse_normalize     main.zig:106-149 normalizeFrequencyTable
 *
 *   PARITY UNPINNED (the reference only fixes constants and intent; its table
 *   builder main.zig:151-189 overwrites entries and its encoder main.zig:42-68 is
 *   unfinished; there is no decoder): everything below "tANS completion". Kept
 *   from the reference: TABLE_LOG = 8 / 256 states (main.zig:80-81), the
 *   {symbol u8, next_state u16, num_bits u8} entry (main.zig:73-77), single state
 *   starting at 0 (main.zig:52), the last input byte stored raw in the first 8
 *   bits (main.zig:55-56), symbols encoded from the end of the input backwards
 *   (main.zig:59-62; the loop's off-by-one that skips input[0] and re-encodes
 *   input[len-1] is treated as a bug: symbols len-2 .. 0 are encoded), the final
 *   state flushed in TABLE_LOG bits (main.zig:65), an LSB-first u64 word
 *   container (main.zig:28-39) and the size formula 8*word_idx + bit_idx/8
 *   (main.zig:67). Chosen here because the reference is silent: the symbol spread
 *   (step 163 = 256/2 + 256/8 + 3, the classic FSE walk) and the state
 *   transition rule (standard tANS).
 */
#include <stdint.h>
#include <string.h>
#include "port.h"

#define TLOG 8
#define TSIZE 256

void port_fse_histogram(const uint8_t* in, uint64_t n, uint64_t* freq) {
    memset(freq, 0, 256 * sizeof(uint64_t));
    for (uint64_t i = 0; i < n; ++i) ++freq[in[i]];
}

/* in place; returns number of present symbols (0 -> table untouched). */
int port_fse_normalize(uint64_t* freq) {
    uint64_t total = 0, num_symbols = 0;
    for (int s = 0; s < 256; ++s) if (freq[s] > 0) { total += freq[s]; ++num_symbols; }
    if (total == 0 || num_symbols == 0) return 0;
    uint64_t distributable = TSIZE - num_symbols;
    double scale = (double)distributable / (double)total;
    uint64_t remaining = TSIZE;
    for (int s = 0; s < 256; ++s) {
        if (freq[s] == 0) continue;
        uint64_t nf = (uint64_t)((double)freq[s] * scale); /* truncation toward zero */
        if (nf == 0) nf = 1;
        freq[s] = nf;
        remaining -= nf;
    }
    while (remaining > 0) {
        uint64_t max_freq = 0; int max_idx = 0;
        for (int s = 0; s < 256; ++s) if (freq[s] > max_freq) { max_freq = freq[s]; max_idx = s; }
        if (max_freq == 0) break;
        freq[max_idx] += 1;
        remaining -= 1;
    }
    return (int)num_symbols;
}

/* ---------------- tANS completion (parity unpinned) ---------------- */

static inline int floor_log2_u32(uint32_t x) { return 31 - __builtin_clz(x); }

/* tt[u] = symbol | next_state << 8 | num_bits << 24 (packed like TT_Entry).
 * enc_state[cum[s] + (x - norm[s])] = state value u reached when symbol s is
 * pushed with sub-state x. cum[257]. */
void port_fse_build_tables(const uint64_t* norm, uint32_t* tt, uint8_t* enc_state, uint16_t* cum) {
    uint8_t sym_at[TSIZE];
    uint32_t pos = 0;
    cum[0] = 0;
    for (int s = 0; s < 256; ++s) {
        cum[s + 1] = (uint16_t)(cum[s] + norm[s]);
        for (uint64_t k = 0; k < norm[s]; ++k) { sym_at[pos] = (uint8_t)s; pos = (pos + 163) & (TSIZE - 1); }
    }
    uint32_t next[256];
    for (int s = 0; s < 256; ++s) next[s] = (uint32_t)norm[s];
    for (uint32_t u = 0; u < TSIZE; ++u) {
        uint32_t s = sym_at[u];
        uint32_t x = next[s]++;
        uint32_t nb = TLOG - (uint32_t)floor_log2_u32(x);
        uint32_t base = (x << nb) - TSIZE;
        tt[u] = s | (base << 8) | (nb << 24);
        enc_state[cum[s] + (x - (uint32_t)norm[s])] = (uint8_t)u;
    }
}

static inline void put_lsb(uint64_t* w, uint64_t* bitpos, uint64_t value, uint32_t nbits) {
    if (!nbits) return;
    uint64_t i = *bitpos >> 6; uint32_t o = (uint32_t)(*bitpos & 63);
    w[i] |= value << o;
    if (o + nbits > 64) w[i + 1] |= value >> (64 - o);
    *bitpos += nbits;
}
static inline uint32_t get_lsb(const uint64_t* w, uint64_t bitpos, uint32_t nbits) {
    if (!nbits) return 0;
    uint64_t i = bitpos >> 6; uint32_t o = (uint32_t)(bitpos & 63);
    uint64_t v = w[i] >> o;
    if (o + nbits > 64) v |= w[i + 1] << (64 - o);
    return (uint32_t)(v & ((1ull << nbits) - 1));
}

/* One stream. words must be zeroed, capacity >= (16 + 8*n)/64 + 2. Returns bits. */
uint64_t port_fse_encode_stream(const uint8_t* in, uint64_t n, const uint64_t* norm,
                                const uint8_t* enc_state, const uint16_t* cum, uint64_t* words) {
    if (n == 0) return 0;
    uint64_t bp = 0;
    put_lsb(words, &bp, in[n - 1], 8);
    uint32_t X = TSIZE; /* state value 0 */
    for (uint64_t i = n - 1; i-- > 0;) {
        uint32_t s = in[i];
        uint32_t f = (uint32_t)norm[s];
        uint32_t nb = TLOG - (uint32_t)floor_log2_u32(f);
        if ((X >> nb) < f) --nb;
        put_lsb(words, &bp, X & ((1u << nb) - 1), nb);
        X = TSIZE + enc_state[cum[s] + ((X >> nb) - f)];
    }
    put_lsb(words, &bp, X - TSIZE, TLOG);
    return bp;
}

/* Returns 0 on success, 1 if the stream does not land back on state 0 / bit 8. */
int port_fse_decode_stream(const uint64_t* words, uint64_t total_bits, uint64_t n,
                           const uint32_t* tt, uint8_t* out) {
    if (n == 0) return total_bits != 0;
    uint64_t pos = total_bits - TLOG;
    uint32_t u = get_lsb(words, pos, TLOG);
    for (uint64_t i = 0; i + 1 < n; ++i) {
        uint32_t e = tt[u];
        out[i] = (uint8_t)(e & 0xFF);
        uint32_t nb = e >> 24;
        pos -= nb;
        u = ((e >> 8) & 0xFFFF) + get_lsb(words, pos, nb);
    }
    out[n - 1] = (uint8_t)get_lsb(words, 0, 8);
    return !(pos == 8 && u == 0);
}

/* C mirror of compress() (main.zig:50-68) for one buffer: histogram, normalise,
 * build tables, encode. norm_out[256] receives the normalised counts the decoder
 * needs. Returns 8*word_idx + bit_idx/8; *total_bits gets the exact bit count. */
uint64_t port_fse_compress(const uint8_t* in, uint64_t n, uint64_t* words, uint64_t* norm_out, uint64_t* total_bits) {
    uint64_t norm[256];
    uint32_t tt[TSIZE]; uint8_t enc_state[TSIZE]; uint16_t cum[257];
    port_fse_histogram(in, n, norm);
    if (!port_fse_normalize(norm)) { *total_bits = 0; memset(norm_out, 0, 256 * 8); return 0; }
    port_fse_build_tables(norm, tt, enc_state, cum);
    memset(words, 0, ((16 + 8 * n) / 64 + 2) * 8);
    uint64_t bits = port_fse_encode_stream(in, n, norm, enc_state, cum, words);
    memcpy(norm_out, norm, sizeof(norm));
    *total_bits = bits;
    return 8 * (bits >> 6) + (bits & 63) / 8;
}

int port_fse_decompress(const uint64_t* words, uint64_t total_bits, uint64_t n, const uint64_t* norm, uint8_t* out) {
    uint32_t tt[TSIZE]; uint8_t enc_state[TSIZE]; uint16_t cum[257];
    if (n == 0) return total_bits != 0;
    port_fse_build_tables(norm, tt, enc_state, cum);
    return port_fse_decode_stream(words, total_bits, n, tt, out);
}
