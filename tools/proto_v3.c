/*
 * CPU prototype of the v3 LZ77 match finder (lz77_v3.cu), written phase by phase the way the
 * CUDA kernel runs, to validate exactness before any GPU time is spent. Not part of the
 * product or the oracle.
 *
 * Idea (DESIGN.md "LZ77 v3"): in the no-expiry occupancy a cluster (maximal run of occupied
 * slots) has as many entries as slots and never interacts with another cluster. An entry whose
 * home is the first slot of its cluster is a HEAD: slot `home` can only ever be taken by heads of
 * that home, so find() of a head is decided by who currently sits in the home slot (a "jump chain":
 * the occupant changes only when it has expired at an arrival). Only clusters that contain an
 * INTRUDER (home strictly inside the cluster) or two different patterns at the head home need a
 * simulation of the placements, and there only the intruders need find().
 *
 *   1. occupancy bitmap, rank            4. jump chains for heads of uniform homes
 *   2. classify: loner / head / intruder 5. per-lane serial simulation of the mixed clusters
 *   3. mark mixed + non-uniform clusters     (liveness bitmask + FIFO expiry + first-fit)
 *   6. slot-0 / table-end cluster serially with the reference's early clear
 *
 * Reference for every F[p]: a plain serial table with lazy expiry + the slot-0 rule.
 *
 * build: gcc -O2 -o /tmp/proto_v3 tools/proto_v3.c oracle/port/lz77_port.c -Ioracle/port -fopenmp
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "port.h"

#define SLOTS (1u << 20)
#define GUARDB 65536u
#define NONE 0xFFFFFFFFu
#define NBINS 1024

static uint32_t word_at(const uint8_t* d, uint32_t n, uint32_t p) {
    uint32_t w = 0;
    for (int k = 0; k < 4; ++k) if (p + k < n) w |= (uint32_t)d[p + k] << (8 * k);
    return w;
}

/* ---- reference: F[p] for EVERY position (find after inserts 0..p-1) */
static void ref_block(const uint8_t* d, uint32_t n, int variant, uint32_t* F) {
    const uint32_t W = variant ? 32768u : 16384u;
    uint32_t* pat = (uint32_t*)malloc(4 * (SLOTS + GUARDB));
    uint32_t* idx = (uint32_t*)malloc(4 * (SLOTS + GUARDB));
    memset(idx, 0xFF, 4 * (SLOTS + GUARDB));
    uint8_t* placed0 = (uint8_t*)calloc(W, 1);
#define LIVE(s, P) (idx[s] != NONE && (uint64_t)idx[s] + W >= (P))
    for (uint32_t p = 0; p < n; ++p) {
        uint32_t w = word_at(d, n, p);
        uint32_t s = port_lz77_hash(w);
        while (LIVE(s, p) && pat[s] != w) ++s;
        F[p] = LIVE(s, p) ? idx[s] : NONE;
        s = port_lz77_hash(w);
        while (LIVE(s, p)) { ++s; if (variant && s == SLOTS) s = 0; }
        pat[s] = w; idx[s] = p;
        uint32_t ring = p & (W - 1);
        int clear0 = (p == W - 1) || (p >= W && placed0[ring]);
        placed0[ring] = (s == 0);
        if (clear0) idx[0] = NONE;
    }
#undef LIVE
    free(pat); free(idx); free(placed0);
}

typedef struct { uint64_t loner, head_pure, head_mixed, intruder, special, nonuni_clusters, mixed_clusters, maxlane, levels, simsteps, walk; } stats_t;

static void v3_block(const uint8_t* d, uint32_t n, int variant, uint32_t* F, stats_t* st) {
    const uint32_t W = variant ? 32768u : 16384u;
    const uint32_t nbits = SLOTS + GUARDB;
    uint32_t* bm = (uint32_t*)calloc(nbits / 32 + 2, 4);
    uint32_t* hs = (uint32_t*)malloc(4 * n);
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
    uint32_t nwords = nbits / 32;
    uint32_t* pre = (uint32_t*)malloc(4 * (nwords + 1));
    pre[0] = 0;
    for (uint32_t w = 0; w < nwords; ++w) pre[w + 1] = pre[w] + (uint32_t)__builtin_popcount(bm[w]);
#define RANK(s) (pre[(s) >> 5] + (uint32_t)__builtin_popcount(bm[(s) >> 5] & ((1u << ((s) & 31)) - 1)))
#define BIT(s) ((bm[(s) >> 5] >> ((s) & 31)) & 1)
    const uint32_t nslots = n;
    uint32_t cut0 = 0; while (BIT(cut0)) ++cut0;
    uint32_t top_start = SLOTS;
    if (variant && BIT(SLOTS - 1)) { top_start = SLOTS - 1; while (top_start > 0 && BIT(top_start - 1)) --top_start; }
    const uint32_t sp_lo_end = RANK(cut0);
    const uint32_t sp_hi_start = top_start < SLOTS ? RANK(top_start) : 0xFFFFFFFFu;

    /* pass A: class + compact home c + compact cluster start ca (tokb[p] = c | ca << 16 on the GPU) */
    uint32_t* cc = (uint32_t*)malloc(4 * n);
    uint32_t* cca = (uint32_t*)malloc(4 * n);
    uint8_t* cls = (uint8_t*)malloc(n);   /* 0 loner, 1 head, 2 intruder, 3 special */
    for (uint32_t p = 0; p < n; ++p) {
        const uint32_t h = hs[p];
        const uint32_t c = RANK(h);
        if (h < cut0 || (variant && h >= top_start)) { cls[p] = 3; cc[p] = c; cca[p] = c; ++st->special; continue; }
        const uint32_t below = BIT(h - 1), above = BIT(h + 1);
        if (!below) {
            if (!above) { cls[p] = 0; F[p] = NONE; ++st->loner; continue; }
            cls[p] = 1; cc[p] = c; cca[p] = c;
        } else {
            uint32_t a = h; while (BIT(a - 1)) --a;
            cls[p] = 2; cc[p] = c; cca[p] = c - (h - a); ++st->intruder;
        }
    }
    /* pass B0: mixed marks + first occurrence per head home */
    uint8_t* MIX = (uint8_t*)calloc(n + 2, 1);
    uint8_t* NONUNI = (uint8_t*)calloc(n + 2, 1);
    uint32_t* cur = (uint32_t*)malloc(4 * (n + 2));
    for (uint32_t i = 0; i < n + 2; ++i) cur[i] = NONE;
    for (uint32_t p = 0; p < n; ++p) {
        if (cls[p] == 2) MIX[cca[p]] = 1;
        else if (cls[p] == 1 && p < cur[cc[p]]) cur[cc[p]] = p;
    }
    /* pass B1: pattern uniformity of every head home */
    for (uint32_t p = 0; p < n; ++p)
        if (cls[p] == 1) { const uint32_t e0 = cur[cc[p]]; if (e0 != p && word_at(d, n, e0) != word_at(d, n, p)) { NONUNI[cc[p]] = 1; MIX[cc[p]] = 1; } }
    for (uint32_t c = 0; c < n; ++c) { st->nonuni_clusters += NONUNI[c]; st->mixed_clusters += MIX[c]; }
    /* pass B2: jump chains, level by level */
    uint8_t* unres = (uint8_t*)calloc(n, 1);
    for (uint32_t p = 0; p < n; ++p) if (cls[p] == 1 && !NONUNI[cc[p]]) unres[p] = 1;
    for (uint32_t level = 0;; ++level) {
        int any = 0;
        if (level) {
            for (uint32_t p = 0; p < n; ++p) if (unres[p]) cur[cc[p]] = NONE;
            for (uint32_t p = 0; p < n; ++p) if (unres[p] && p < cur[cc[p]]) cur[cc[p]] = p;
        }
        for (uint32_t p = 0; p < n; ++p) if (unres[p]) {
            const uint32_t e = cur[cc[p]];
            if (e == p) { F[p] = NONE; unres[p] = 0; }
            else if (p - e <= W) { F[p] = e; unres[p] = 0; }
            else any = 1;
        }
        if (level + 1 > st->levels) st->levels = level + 1;
        if (!any) break;
    }
    for (uint32_t p = 0; p < n; ++p) if (cls[p] == 1) { if (MIX[cc[p]]) ++st->head_mixed; else ++st->head_pure; }

    /* partition: lane lists (bin = ca >> 6) in time order; entries of mixed clusters only */
    uint32_t* cnt = (uint32_t*)calloc(NBINS + 1, 4);
    for (uint32_t p = 0; p < n; ++p) if ((cls[p] == 1 || cls[p] == 2) && MIX[cca[p]]) ++cnt[cca[p] >> 6];
    uint32_t* start = (uint32_t*)malloc(4 * (NBINS + 2));
    start[0] = 0; for (int b = 0; b < NBINS; ++b) start[b + 1] = start[b] + cnt[b];
    uint32_t* lpos = (uint32_t*)malloc(4 * (n + 1));
    uint32_t* lslot = (uint32_t*)malloc(4 * (n + 1));
    uint32_t* fill = (uint32_t*)calloc(NBINS, 4);
    for (uint32_t p = 0; p < n; ++p) if ((cls[p] == 1 || cls[p] == 2) && MIX[cca[p]]) { const uint32_t b = cca[p] >> 6; lpos[start[b] + fill[b]++] = p; }

    /* simulation: T[slot] = position of the entry placed there, M = liveness bits (compact space) */
    uint16_t* T = (uint16_t*)calloc(n + 64, 2);
    uint32_t* M = (uint32_t*)calloc(n / 32 + 4, 4);
    for (int b = 0; b < NBINS; ++b) {
        const uint32_t lo = start[b], hi = start[b + 1];
        if (hi - lo > st->maxlane) st->maxlane = hi - lo;
        uint32_t ei = lo;
        uint32_t hint_c = NONE, hint_w = 0, hw = 0;
        for (uint32_t i = lo; i < hi; ++i) {
            const uint32_t p = lpos[i], c = cc[p];
            ++st->simsteps;
            /* FIFO expiry: entry j is live at time p iff j + W >= p */
            while (ei < i && lpos[ei] + W < p) {
                const uint32_t s = lslot[ei], sw = s >> 5, sbit = 1u << (s & 31);
                M[sw] &= ~sbit;
                if (s >= hint_c) {
                    if (sw < hint_w) { hint_w = sw; hw = ~sbit; }
                    else if (sw == hint_w) hw &= ~sbit;
                }
                ++ei;
            }
            const int need_find = cls[p] == 2 || NONUNI[cca[p]];
            const int is_head = cls[p] == 1;
            if (need_find) {
                uint32_t s = c;
                const uint32_t w = word_at(d, n, p);
                uint32_t m = NONE;
                while ((M[s >> 5] >> (s & 31)) & 1u) {
                    ++st->walk;
                    if (word_at(d, n, T[s]) == w) { m = T[s]; break; }
                    ++s;
                }
                F[p] = m;
            }
            uint32_t wi, z;
            if (c == hint_c) { wi = hint_w; z = ~hw; if (wi == (c >> 5)) z &= 0xFFFFFFFFu << (c & 31); }
            else { wi = c >> 5; z = ~M[wi] & (0xFFFFFFFFu << (c & 31)); }
            while (!z) { ++wi; z = ~M[wi]; }
            const uint32_t e = (wi << 5) + (uint32_t)__builtin_ctz(z);
            const uint32_t ebit = 1u << (e & 31);
            /* check against the plain search */
            { uint32_t s2 = c; while ((M[s2 >> 5] >> (s2 & 31)) & 1u) ++s2; if (s2 != e) { printf("HINT MISMATCH p %u c %u: %u vs %u\n", p, c, e, s2); exit(3); } }
            if (is_head) {
                if (c != hint_c || wi != hint_w) { hw = ~z; if (wi == (c >> 5)) hw |= ~(0xFFFFFFFFu << (c & 31)); }
                hint_c = c; hint_w = wi; hw |= ebit;
            } else if (e >= hint_c && (e >> 5) == hint_w) hw |= ebit;
            M[e >> 5] |= ebit;
            T[e] = (uint16_t)p;
            lslot[i] = e;
        }
    }

    /* special clusters (slot 0 / table end), serially in time order with the reference's clear queue */
    {
        uint16_t* TS = (uint16_t*)calloc(n + 64, 2);   /* position + 1, 0 = never used (lazy expiry) */
        uint32_t clr[64]; uint32_t qh = 0, qt = 0; clr[qt++ & 63] = W - 1;
        for (uint32_t q = 0; q < n; ++q) {
            if (cls[q] != 3) continue;
            const uint32_t c = cc[q];
            while (qh < qt && clr[qh & 63] < q) { if (sp_lo_end) TS[0] = 0; ++qh; }
            const uint32_t dthr = q > W ? q - W : 0;
            const uint32_t w = word_at(d, n, q);
            uint32_t k = c, m = NONE; int ran_off = 0;
            for (;;) {
                const uint32_t v = TS[k];
                if (v <= dthr) break;
                if (word_at(d, n, v - 1) == w) { m = v - 1; break; }
                if (k + 1 == nslots) { ran_off = 1; break; }
                ++k;
            }
            F[q] = m;
            uint32_t e = ran_off ? 0 : k;
            for (;;) { if (TS[e] <= dthr) break; ++e; if (e == nslots) e = 0; }
            if (q != 65535u) TS[e] = (uint16_t)(q + 1);
            if (e == 0 && sp_lo_end) { clr[qt & 63] = q + W; ++qt; }
            if (qh < qt && clr[qh & 63] == q) { if (sp_lo_end) TS[0] = 0; ++qh; }
        }
        free(TS);
    }
    free(bm); free(hs); free(pre); free(cc); free(cca); free(cls); free(MIX); free(NONUNI); free(cur); free(unres);
    free(cnt); free(start); free(lpos); free(lslot); free(fill); free(T); free(M);
}

int main(int argc, char** argv) {
    if (argc < 4) { fprintf(stderr, "usage: proto_v3 file variant block\n"); return 2; }
    FILE* fp = fopen(argv[1], "rb");
    if (!fp) { perror(argv[1]); return 2; }
    fseek(fp, 0, SEEK_END); long sz = ftell(fp); fseek(fp, 0, SEEK_SET);
    uint8_t* data = (uint8_t*)malloc(sz); if (fread(data, 1, sz, fp) != (size_t)sz) return 2; fclose(fp);
    int variant = atoi(argv[2]); uint32_t block = (uint32_t)atoi(argv[3]);
    uint64_t mism = 0, checked = 0, nblk = 0;
    stats_t st; memset(&st, 0, sizeof(st));
    for (long off = 0; off < sz; off += block) {
        uint32_t n = (uint32_t)(sz - off < block ? sz - off : block);
        uint32_t* F = (uint32_t*)malloc(4 * n); uint32_t* Fr = (uint32_t*)malloc(4 * n);
        for (uint32_t i = 0; i < n; ++i) F[i] = 0xFFFFFFFDu;
        v3_block(data + off, n, variant, F, &st);
        ref_block(data + off, n, variant, Fr);
        for (uint32_t i = 0; i < n; ++i) { ++checked; if (Fr[i] != F[i]) { if (mism < 8) printf("block@%ld pos %u: ref %u proto %u\n", off, i, Fr[i], F[i]); ++mism; } }
        free(F); free(Fr); ++nblk;
    }
    printf("variant %d block %u: %lu positions, mismatches %lu | per block: loner %.0f head(pure) %.0f head(mixed) %.0f intruder %.0f special %.1f | mixed clusters %.0f nonuni %.1f | max lane list %lu levels %lu sim steps %.0f walk/step %.2f\n",
           variant, block, checked, mism, (double)st.loner / nblk, (double)st.head_pure / nblk, (double)st.head_mixed / nblk, (double)st.intruder / nblk,
           (double)st.special / nblk, (double)st.mixed_clusters / nblk, (double)st.nonuni_clusters / nblk, st.maxlane, st.levels, (double)st.simsteps / nblk,
           st.simsteps ? (double)st.walk / st.simsteps : 0.0);
    return mism != 0;
}
