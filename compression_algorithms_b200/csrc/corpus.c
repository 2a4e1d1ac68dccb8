/*
 * Seeded synthetic inputs (host only, no CUDA): the reference drivers read
 * ../../data/enwik8|9 (/root/reference/algorithms/huffman/main.c:35-39,
 * algorithms/lz77/main.c:11-16, algorithms/deflate/main.c:7) fetched by
 * get_data.sh; there is no network here, so benches and tests use these
 * generators instead (SURVEY.md §8d "Inputs").
 *
 * kind 0  enwik-shaped : Zipf(1.0) words from a 50 000-word vocabulary, sentence
 *                        punctuation/capitalisation, ~5 % wiki/XML markup lines,
 *                        ~2 % multi-byte UTF-8; ~190 distinct byte values, no 0x00.
 * kind 1  low-entropy  : i.i.d. "acgt".
 * kind 2  skewed       : two symbols, P('a') = 0.95.
 * kind 3  near-random  : uniform bytes 1..255.
 *
 * The buffer is a concatenation of independent 1 MiB chunks, chunk k seeded from
 * (seed, k), so generation is parallel and any prefix is reproducible.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <omp.h>

#define VOCAB 50000
#define CHUNK (1u << 20)

typedef struct { uint64_t s; } rng_t;
static inline uint64_t rng_next(rng_t* r) {
    uint64_t z = (r->s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline uint32_t rng_below(rng_t* r, uint32_t n) { return (uint32_t)((rng_next(r) >> 32) * (uint64_t)n >> 32); }

typedef struct {
    char     words[VOCAB][13];
    uint8_t  wlen[VOCAB];
    uint32_t cum[VOCAB]; /* cumulative Zipf weights scaled to 2^32 */
} vocab_t;

static const char LETTERS[] = /* rough English letter frequencies, 100 slots */
    "eeeeeeeeeeeettttttttttaaaaaaaaooooooooiiiiiiinnnnnnnsssssshhhhhhrrrrrrddddlllluuucccmmmwwffggyyppbbvkjxqz";

static void build_vocab(vocab_t* v, uint64_t seed) {
    rng_t r = { seed ^ 0xC0FFEE1234ull };
    static const char* COMMON[] = { "the", "of", "and", "in", "to", "a", "is", "was", "for", "as", "by", "with",
        "that", "on", "from", "at", "his", "it", "an", "are", "which", "be", "this", "or", "he", "also", "has",
        "were", "had", "not", "but", "one", "their", "its", "first", "have", "been", "other", "new", "after" };
    int ncommon = (int)(sizeof(COMMON) / sizeof(COMMON[0]));
    for (int i = 0; i < VOCAB; ++i) {
        if (i < ncommon) {
            strcpy(v->words[i], COMMON[i]);
            v->wlen[i] = (uint8_t)strlen(COMMON[i]);
            continue;
        }
        /* longer words further down the ranking, 2..12 letters */
        int base = 3 + (32 - __builtin_clz((unsigned)i + 1u)) / 3;
        int len = base + (int)rng_below(&r, 5) - 2;
        if (len < 2) len = 2;
        if (len > 12) len = 12;
        for (int k = 0; k < len; ++k) v->words[i][k] = LETTERS[rng_below(&r, (uint32_t)(sizeof(LETTERS) - 1))];
        v->words[i][len] = 0;
        v->wlen[i] = (uint8_t)len;
    }
    double total = 0.0;
    for (int i = 0; i < VOCAB; ++i) total += 1.0 / (double)(i + 1);
    double acc = 0.0;
    for (int i = 0; i < VOCAB; ++i) {
        acc += 1.0 / (double)(i + 1);
        double c = acc / total * 4294967296.0;
        v->cum[i] = c >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)c;
    }
    v->cum[VOCAB - 1] = 0xFFFFFFFFu;
}

static inline int zipf_draw(const vocab_t* v, rng_t* r) {
    uint32_t u = (uint32_t)(rng_next(r) >> 32);
    int lo = 0, hi = VOCAB - 1;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (v->cum[mid] < u) lo = mid + 1; else hi = mid;
    }
    return lo;
}

typedef struct { uint8_t* p; uint8_t* end; } sink_t;
static inline void put(sink_t* s, const void* src, size_t n) {
    size_t room = (size_t)(s->end - s->p);
    if (n > room) n = room;
    memcpy(s->p, src, n);
    s->p += n;
}
static inline void putc_(sink_t* s, char c) { if (s->p < s->end) *s->p++ = (uint8_t)c; }
static inline void puts_(sink_t* s, const char* z) { put(s, z, strlen(z)); }

static void put_word(sink_t* s, const vocab_t* v, rng_t* r, int capital) {
    int w = zipf_draw(v, r);
    if (capital) {
        putc_(s, (char)(v->words[w][0] - 32));
        put(s, v->words[w] + 1, v->wlen[w] - 1u);
    } else {
        put(s, v->words[w], v->wlen[w]);
    }
}

static void put_utf8(sink_t* s, rng_t* r) {
    uint32_t k = rng_below(r, 100);
    if (k < 55) {                 /* Latin-1 supplement: C3 80..BF */
        putc_(s, (char)0xC3); putc_(s, (char)(0x80 + rng_below(r, 64)));
    } else if (k < 80) {          /* Cyrillic: D0/D1 80..BF, 2-5 letters */
        int n = 2 + (int)rng_below(r, 4);
        for (int i = 0; i < n; ++i) { putc_(s, (char)(0xD0 + rng_below(r, 2))); putc_(s, (char)(0x80 + rng_below(r, 64))); }
    } else if (k < 95) {          /* CJK: E4..E9 80..BF 80..BF, 1-3 glyphs */
        int n = 1 + (int)rng_below(r, 3);
        for (int i = 0; i < n; ++i) { putc_(s, (char)(0xE4 + rng_below(r, 6))); putc_(s, (char)(0x80 + rng_below(r, 64))); putc_(s, (char)(0x80 + rng_below(r, 64))); }
    } else {                      /* punctuation block: E2 80 90..A6 */
        putc_(s, (char)0xE2); putc_(s, (char)0x80); putc_(s, (char)(0x90 + rng_below(r, 23)));
    }
}

static void put_number(sink_t* s, rng_t* r, int digits) {
    for (int i = 0; i < digits; ++i) putc_(s, (char)('0' + rng_below(r, 10)));
}

static void markup_line(sink_t* s, const vocab_t* v, rng_t* r) {
    switch (rng_below(r, 9)) {
    case 0: puts_(s, "  <page>\n    <title>"); put_word(s, v, r, 1); putc_(s, ' '); put_word(s, v, r, 1); puts_(s, "</title>\n"); break;
    case 1: puts_(s, "    <id>"); put_number(s, r, 3 + (int)rng_below(r, 5)); puts_(s, "</id>\n"); break;
    case 2: puts_(s, "      <timestamp>200"); put_number(s, r, 1); putc_(s, '-'); put_number(s, r, 2); putc_(s, '-'); put_number(s, r, 2);
            putc_(s, 'T'); put_number(s, r, 2); putc_(s, ':'); put_number(s, r, 2); putc_(s, ':'); put_number(s, r, 2); puts_(s, "Z</timestamp>\n"); break;
    case 3: puts_(s, "      <contributor>\n        <username>"); put_word(s, v, r, 1); put_number(s, r, 2); puts_(s, "</username>\n      </contributor>\n"); break;
    case 4: puts_(s, "== "); put_word(s, v, r, 1); putc_(s, ' '); put_word(s, v, r, 0); puts_(s, " ==\n"); break;
    case 5: puts_(s, "{{"); put_word(s, v, r, 0); putc_(s, '|'); put_word(s, v, r, 0); putc_(s, '='); put_word(s, v, r, 0); puts_(s, "}}\n"); break;
    case 6: puts_(s, "* [[Category:"); put_word(s, v, r, 1); putc_(s, ' '); put_word(s, v, r, 0); puts_(s, "]]\n"); break;
    case 7: puts_(s, "      <text xml:space=\"preserve\">#REDIRECT [["); put_word(s, v, r, 1); puts_(s, "]]</text>\n    </revision>\n  </page>\n"); break;
    default: puts_(s, "* [http://www."); put_word(s, v, r, 0); puts_(s, ".org/"); put_word(s, v, r, 0); putc_(s, '_'); put_number(s, r, 4);
             puts_(s, ".html "); put_word(s, v, r, 1); puts_(s, "]\n"); break;
    }
}

static void gen_enwik_chunk(uint8_t* out, size_t n, const vocab_t* v, uint64_t seed, uint64_t chunk) {
    rng_t r = { seed * 0x9E3779B97F4A7C15ull + chunk * 0xD1B54A32D192ED03ull + 1 };
    sink_t s = { out, out + n };
    while (s.p < s.end) {
        if (rng_below(&r, 100) < 12) { markup_line(&s, v, &r); continue; }
        /* one paragraph of 2..7 sentences */
        int sentences = 2 + (int)rng_below(&r, 6);
        for (int q = 0; q < sentences && s.p < s.end; ++q) {
            int words = 6 + (int)rng_below(&r, 20);
            for (int w = 0; w < words; ++w) {
                uint32_t k = rng_below(&r, 1000);
                if (k < 25) { puts_(&s, "[["); put_word(&s, v, &r, 1); if (k < 10) { putc_(&s, '|'); put_word(&s, v, &r, 0); } puts_(&s, "]]"); }
                else if (k < 33) { puts_(&s, "&quot;"); put_word(&s, v, &r, 0); puts_(&s, "&quot;"); }
                else if (k < 40) { puts_(&s, "''"); put_word(&s, v, &r, 0); puts_(&s, "''"); }
                else if (k < 52) { put_number(&s, &r, 1 + (int)rng_below(&r, 4)); }
                else if (k < 60) { put_utf8(&s, &r); }
                else if (k < 64) { puts_(&s, "&amp;"); }
                else if (k < 67) { puts_(&s, "&lt;ref&gt;"); put_word(&s, v, &r, 1); puts_(&s, "&lt;/ref&gt;"); }
                else put_word(&s, v, &r, w == 0 || k > 960);
                if (w + 1 < words) {
                    uint32_t p = rng_below(&r, 100);
                    if (p < 7) puts_(&s, ", "); else if (p < 8) puts_(&s, "; "); else if (p < 9) puts_(&s, " ("); else if (p < 10) puts_(&s, ") ");
                    else if (p < 11) puts_(&s, " - "); else if (p < 12) puts_(&s, ": "); else putc_(&s, ' ');
                }
            }
            uint32_t e = rng_below(&r, 100);
            puts_(&s, e < 90 ? ". " : (e < 95 ? "? " : "! "));
        }
        puts_(&s, "\n\n");
    }
}

static void gen_simple_chunk(uint8_t* out, size_t n, int kind, uint64_t seed, uint64_t chunk) {
    rng_t r = { seed * 0x9E3779B97F4A7C15ull + chunk * 0xD1B54A32D192ED03ull + (uint64_t)kind * 77 + 1 };
    for (size_t i = 0; i < n; ++i) {
        uint32_t u = (uint32_t)(rng_next(&r) >> 32);
        if (kind == 1) out[i] = (uint8_t)"acgt"[u >> 30];
        else if (kind == 2) out[i] = (uint8_t)(u < 4080218931u ? 'a' : 'b'); /* 0.95 * 2^32 */
        else out[i] = (uint8_t)(1u + (uint32_t)(((uint64_t)u * 255u) >> 32));
    }
}

/* C-ABI: fill out[0..n) ; returns 0 on success. */
int b200_corpus_generate(uint8_t* out, uint64_t n, int kind, uint64_t seed) {
    if (kind < 0 || kind > 3) return 1;
    vocab_t* v = NULL;
    if (kind == 0) {
        v = (vocab_t*)malloc(sizeof(vocab_t));
        if (!v) return 2;
        build_vocab(v, seed);
    }
    uint64_t nchunks = (n + CHUNK - 1) / CHUNK;
#pragma omp parallel for schedule(dynamic, 8)
    for (int64_t c = 0; c < (int64_t)nchunks; ++c) {
        uint64_t off = (uint64_t)c * CHUNK;
        size_t len = (size_t)(n - off < CHUNK ? n - off : CHUNK);
        if (kind == 0) gen_enwik_chunk(out + off, len, v, seed, (uint64_t)c);
        else gen_simple_chunk(out + off, len, kind, seed, (uint64_t)c);
    }
    free(v);
    return 0;
}

/* Bytes [start, start + len) of the buffer b200_corpus_generate(out, N, kind, seed) would produce for any
 * N >= start + len (chunks are independent), without generating what lies before: lets every rank of a
 * multi-GPU run materialise only its own shard of ONE global buffer. */
int b200_corpus_generate_range(uint8_t* out, uint64_t start, uint64_t len, int kind, uint64_t seed) {
    if (kind < 0 || kind > 3) return 1;
    if (len == 0) return 0;
    vocab_t* v = NULL;
    if (kind == 0) {
        v = (vocab_t*)malloc(sizeof(vocab_t));
        if (!v) return 2;
        build_vocab(v, seed);
    }
    const uint64_t c0 = start / CHUNK, c1 = (start + len - 1) / CHUNK;
    int rc = 0;
#pragma omp parallel
    {
        uint8_t* tmp = (uint8_t*)malloc(CHUNK);
        if (!tmp) {
#pragma omp atomic write
            rc = 2;
        }
#pragma omp for schedule(dynamic, 8)
        for (int64_t c = (int64_t)c0; c <= (int64_t)c1; ++c) {
            if (!tmp) continue;
            const uint64_t off = (uint64_t)c * CHUNK;
            const uint64_t lo = off > start ? off : start, hi = off + CHUNK < start + len ? off + CHUNK : start + len;
            if (lo == off && hi == off + CHUNK) {   /* whole chunk: straight into place */
                if (kind == 0) gen_enwik_chunk(out + (lo - start), CHUNK, v, seed, (uint64_t)c);
                else gen_simple_chunk(out + (lo - start), CHUNK, kind, seed, (uint64_t)c);
            } else {
                /* a chunk's bytes do not depend on how much of it is asked for beyond the prefix property, so
                 * generate the chunk up to `hi` and keep the tail */
                const size_t want = (size_t)(hi - off);
                if (kind == 0) gen_enwik_chunk(tmp, want, v, seed, (uint64_t)c);
                else gen_simple_chunk(tmp, want, kind, seed, (uint64_t)c);
                memcpy(out + (lo - start), tmp + (lo - off), (size_t)(hi - lo));
            }
        }
        free(tmp);
    }
    free(v);
    return rc;
}
