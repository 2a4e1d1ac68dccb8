/*
 * CPU prototype of the v2 (shared-memory) LZ77 match finder, written warp-by-warp the
 * way the CUDA kernel runs, to validate exactness against the oracle port before any
 * GPU time is spent. Not part of the product or the oracle.
 *
 *   1. no-expiry linear-probing occupancy bitmap  -> compact slot index c(h) = rank(h)
 *   2. 32 slot-space ranges cut at cluster ends    -> per-range time-ordered lists
 *   3. per range: batches of 32 positions walk the compact table speculatively;
 *      conflicts = lanes sharing a first-dead slot; commit the conflict-free prefix
 *   4. slot-0 cluster handled serially with the reference's early-clear rule
 *
 * build: gcc -O2 -o /tmp/proto_v2 tools/proto_v2.c oracle/port/lz77_port.c -Ioracle/port -fopenmp
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "port.h"

#define SLOTS (1u << 20)
#define GUARDB 65536u
#define NONE 0xFFFFFFFFu
#define NR 32

static uint32_t word_at(const uint8_t* d, uint32_t n, uint32_t p) {
    uint32_t w = 0;
    for (int k = 0; k < 4; ++k) if (p + k < n) w |= (uint32_t)d[p + k] << (8 * k);
    return w;
}
static uint8_t byte_at(const uint8_t* d, uint32_t n, uint32_t p) { return p < n ? d[p] : 0; }

typedef struct { uint64_t steps, committed, walk, maxwalk, special, maxrange, minrange; } stats_t;

/* F[q] for every position (find result after inserts 0..q-1), NONE if none */
static void v2_block(const uint8_t* d, uint32_t n, int variant, uint32_t* F, stats_t* st) {
    const uint32_t W = variant ? 32768u : 16384u;
    const uint32_t nbits = SLOTS + GUARDB;
    uint32_t* bm = (uint32_t*)calloc(nbits / 32 + 2, 4);
    uint32_t* hs = (uint32_t*)malloc(4 * n);
    /* 1. occupancy */
    for (uint32_t i = 0; i < n; ++i) {
        uint32_t h = port_lz77_hash(word_at(d, n, i));
        hs[i] = h;
        uint32_t s = h;
        for (;;) {
            if (!(bm[s >> 5] >> (s & 31) & 1)) { bm[s >> 5] |= 1u << (s & 31); break; }
            ++s;
            if (variant && s == SLOTS) s = 0;
        }
    }
    /* rank prefix per word */
    uint32_t nwords = nbits / 32;
    uint32_t* pre = (uint32_t*)malloc(4 * (nwords + 1));
    pre[0] = 0;
    for (uint32_t w = 0; w < nwords; ++w) pre[w + 1] = pre[w] + (uint32_t)__builtin_popcount(bm[w]);
#define RANK(s) (pre[(s) >> 5] + (uint32_t)__builtin_popcount(bm[(s) >> 5] & ((1u << ((s) & 31)) - 1)))
#define BIT(s) ((bm[(s) >> 5] >> ((s) & 31)) & 1)
    /* 2. cuts: first zero bit at/after r << 15 */
    uint32_t cut[NR + 1];
    for (int r = 0; r < NR; ++r) { uint32_t s = (uint32_t)r << 15; while (BIT(s)) ++s; cut[r] = s; }
    cut[NR] = nbits;
    /* special region: the run of ones starting at slot 0 = [0, cut[0]); for the wrapping
     * variant also the run ending at the last slot */
    uint32_t top_start = SLOTS;
    if (variant && BIT(SLOTS - 1)) { top_start = SLOTS - 1; while (top_start > 0 && BIT(top_start - 1)) --top_start; }
    uint32_t sp_lo_end = RANK(cut[0]);                 /* compact indices [0, sp_lo_end) */
    uint32_t sp_hi_start = top_start < SLOTS ? RANK(top_start) : n; /* [sp_hi_start, n) */
    /* lists */
    uint32_t* cnt = (uint32_t*)calloc(NR, 4);
    uint8_t* rid = (uint8_t*)malloc(n);
    uint16_t* cidx = (uint16_t*)malloc(2 * n);
    for (uint32_t i = 0; i < n; ++i) {
        uint32_t h = hs[i];
        int r = (int)(h >> 15);
        if (r >= NR) r = NR - 1;
        while (r > 0 && h < cut[r]) --r;
        if (h >= top_start) r = 0;                      /* wrapped cluster belongs to the special warp */
        rid[i] = (uint8_t)r; ++cnt[r];
        cidx[i] = (uint16_t)RANK(h);
    }
    uint32_t* start = (uint32_t*)malloc(4 * (NR + 1));
    start[0] = 0;
    for (int r = 0; r < NR; ++r) start[r + 1] = start[r] + cnt[r];
    uint32_t* list = (uint32_t*)malloc(4 * n);
    uint32_t* fill = (uint32_t*)calloc(NR, 4);
    for (uint32_t i = 0; i < n; ++i) list[start[rid[i]] + fill[rid[i]]++] = i;
    st->maxrange = 0; st->minrange = n;
    for (int r = 0; r < NR; ++r) { if (cnt[r] > st->maxrange) st->maxrange = cnt[r]; if (cnt[r] < st->minrange) st->minrange = cnt[r]; }

    /* 3. simulation. T holds keys: 0 = empty, else position + 1 */
    uint16_t* T = (uint16_t*)calloc(n + 64, 2);
    uint16_t* B1 = (uint16_t*)calloc(n / 32 + 2, 2);
    uint32_t gmin[32], gstart[32]; for (int k = 0; k < 32; ++k) { gmin[k] = 0; gstart[k] = 0xFFFFFFFFu; }
    uint32_t clr[64]; int qh = 0, qt = 0; clr[qt++] = W - 1;   /* slot-0 clear times */
    for (int r = 0; r < NR; ++r) {
        uint32_t cur = start[r], end = start[r + 1];
        while (cur < end) {
            uint32_t lanes = end - cur < 32 ? end - cur : 32;
            uint32_t e[32], f[32]; int special[32];
            uint32_t L = lanes;
            for (uint32_t k = 0; k < lanes; ++k) {
                uint32_t q = list[cur + k], c = cidx[q];
                special[k] = (c < sp_lo_end) || (c >= sp_hi_start);
            }
            if (special[0]) {
                /* serial path with the exact reference semantics */
                uint32_t q = list[cur], c = cidx[q];
                while (qh < qt && clr[qh] < q) { if (sp_lo_end) { T[0] = 0; B1[0] = 0; } ++qh; }   /* clears that happened before q */
                uint32_t dthr = q > W ? q - W : 0;
                uint32_t w = word_at(d, n, q);
                uint32_t k = c, m = NONE;
                for (;;) {
                    uint32_t v = T[k];
                    if (v <= dthr) break;
                    if (word_at(d, n, v - 1) == w) { m = v - 1; break; }
                    if (k + 1 == n) { k = n; break; }               /* find does not wrap: runs into never-written guard */
                    ++k;
                }
                F[q] = m;
                uint32_t kk = (k == n) ? 0 : k;                     /* insert continues, wrapping (deflate variant) */
                for (;;) { if (T[kk] <= dthr) break; ++kk; if (kk == n) kk = 0; }
                if (q != 65535) T[kk] = (uint16_t)(q + 1);
                if (kk == 0 && sp_lo_end > 0) { if (qt < 64) clr[qt++] = q + W; }
                if (qh < qt && clr[qh] == q) { if (sp_lo_end) { T[0] = 0; B1[0] = 0; } ++qh; }
                ++cur; ++st->steps; ++st->committed; ++st->special;
                continue;
            }
            /* lanes up to the first special one take part; in-batch resolution: every lane keeps a
             * cursor kk (all slots in [c,kk) proven live), re-validates it after each commit round,
             * and a lane commits once no lower uncommitted lane shares its cursor */
            for (uint32_t k = 0; k < lanes; ++k) if (special[k]) { L = k; break; }
            uint32_t kk[32], fm[32], w[32], dth[32]; int pend[32], done[32];
            for (uint32_t k = 0; k < L; ++k) {
                uint32_t q = list[cur + k];
                kk[k] = cidx[q]; fm[k] = NONE; pend[k] = 1; done[k] = 0; gstart[k] = (kk[k] & 31) == 0 ? kk[k] : 0xFFFFFFFFu; gmin[k] = 0xFFFF;
                w[k] = word_at(d, n, q); dth[k] = q > W ? q - W : 0;
            }
            uint32_t ncommitted = 0;
            while (ncommitted < L) {
                ++st->steps;
                for (uint32_t k = 0; k < L; ++k) {          /* advance (parallel over lanes) */
                    if (done[k]) continue;
                    uint32_t steps = 0;
                    for (;;) {
                        uint32_t c = kk[k];
                        /* whole-group skip through the lower-bound table once the find is resolved */
                        if (!pend[k] && (c & 31) == 0 && c + 32 <= n && B1[c >> 5] > dth[k]) { kk[k] = c + 32; ++steps; continue; }
                        uint32_t v = T[c];
                        ++steps;
                        if (v <= dth[k]) break;
                        if (pend[k] && word_at(d, n, v - 1) == w[k]) { fm[k] = v - 1; pend[k] = 0; }
                        /* refresh the bound when a full group was scanned from its first slot */
                        if ((c & 31) == 0) gmin[k] = v; else if (v < gmin[k]) gmin[k] = v;
                        if ((c & 31) == 31 && gstart[k] == (c & ~31u)) B1[c >> 5] = (uint16_t)gmin[k];
                        kk[k] = c + 1;
                        if ((kk[k] & 31) == 0) gstart[k] = kk[k];
                    }
                    st->walk += steps; if (steps > st->maxwalk) st->maxwalk = steps;
                }
                uint32_t first_conf = L;
                for (uint32_t k = 0; k < L && first_conf == L; ++k) {
                    if (done[k]) continue;
                    for (uint32_t j = 0; j < k; ++j) if (!done[j] && kk[j] == kk[k]) { first_conf = k; break; }
                }
                for (uint32_t k = 0; k < first_conf; ++k) {
                    if (done[k]) continue;
                    uint32_t q = list[cur + k];
                    F[q] = fm[k];
                    if (q != 65535) T[kk[k]] = (uint16_t)(q + 1);
                    done[k] = 1; ++ncommitted;
                }
            }
            cur += L; st->committed += L;
        }
    }
    free(bm); free(hs); free(pre); free(cnt); free(rid); free(cidx); free(start); free(list); free(fill); free(T); free(B1);
}

int main(int argc, char** argv) {
    if (argc < 4) { fprintf(stderr, "usage: proto_v2 file variant block\n"); return 2; }
    FILE* fp = fopen(argv[1], "rb");
    fseek(fp, 0, SEEK_END); long sz = ftell(fp); fseek(fp, 0, SEEK_SET);
    uint8_t* data = (uint8_t*)malloc(sz); fread(data, 1, sz, fp); fclose(fp);
    int variant = atoi(argv[2]); uint32_t block = (uint32_t)atoi(argv[3]);
    uint64_t mism = 0, checked = 0;
    stats_t st; memset(&st, 0, sizeof(st));
    uint64_t maxr = 0, minr = 1 << 30;
    for (long off = 0; off < sz; off += block) {
        uint32_t n = (uint32_t)(sz - off < block ? sz - off : block);
        uint32_t* F = (uint32_t*)malloc(4 * n); uint32_t* Fp = (uint32_t*)malloc(4 * n);
        for (uint32_t i = 0; i < n; ++i) Fp[i] = 0xFFFFFFFEu;
        v2_block(data + off, n, variant, F, &st);
        if (st.maxrange > maxr) maxr = st.maxrange; if (st.minrange < minr) minr = st.minrange;
        uint8_t* out = (uint8_t*)malloc(2 * n + 16); uint64_t sz2;
        if (variant) port_deflate_lz77_compress(data + off, n, out, &sz2, Fp); else port_lz77_compress(data + off, n, out, &sz2, Fp);
        for (uint32_t i = 0; i < n; ++i) if (Fp[i] != 0xFFFFFFFEu) { ++checked; if (Fp[i] != F[i]) { if (mism < 5) printf("block@%ld pos %u: port %u proto %u\n", off, i, Fp[i], F[i]); ++mism; } }
        free(F); free(Fp); free(out);
    }
    printf("variant %d block %u: checked %lu token-start finds, mismatches %lu | steps %lu committed/step %.2f walk/pos %.2f maxwalk %lu special %lu range max %lu min %lu\n",
           variant, block, checked, mism, st.steps, (double)st.committed / st.steps, (double)st.walk / st.committed, st.maxwalk, st.special, maxr, minr);
    return mism != 0;
}
