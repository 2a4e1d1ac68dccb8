/* CPU prototype of the v4 match finder (csrc/lz77_v4.cu): per-cluster processing, clusters above small_max by the two-phase slot sweeps; checks every F(p) against a serial table. Not part of the product or the oracle.
   build: gcc -O2 -o /tmp/proto_v4 tools/proto_v4.c oracle/port/lz77_port.c -Ioracle/port ; run: /tmp/proto_v4 file 65536 [small_max] */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "port.h"
#define SLOTS (1u << 20)
#define GUARDB 65536u
#define NONE 0xFFFFFFFFu
static uint32_t W = 32768;
static int SMALL_MAX = 16;
static uint32_t word_at(const uint8_t* d, uint32_t n, uint32_t p) { uint32_t w = 0; for (int k = 0; k < 4; ++k) if (p + k < n) w |= (uint32_t)d[p + k] << (8 * k); return w; }

static void ref_block(const uint8_t* d, uint32_t n, int variant, uint32_t* F) {
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

typedef struct { uint32_t t, o; } ent_t;   /* time, home offset inside the cluster */
static int cmp_ent(const void* a, const void* b) { return (int)((const ent_t*)a)->t - (int)((const ent_t*)b)->t; }

/* general serial simulation (any number of phases) */
static void sim_small(const uint8_t* d, uint32_t n, ent_t* e, uint32_t m, uint32_t* F) {
    uint32_t occ[64]; /* entry index in slot, NONE */
    for (uint32_t j = 0; j < m; ++j) occ[j] = NONE;
    for (uint32_t i = 0; i < m; ++i) {
        const uint32_t p = e[i].t, w = word_at(d, n, p);
        uint32_t j = e[i].o, f = NONE;
        for (;; ++j) {
            const int live = occ[j] != NONE && e[occ[j]].t + W >= p;
            if (!live) break;
            if (f == NONE && word_at(d, n, e[occ[j]].t) == w) { f = e[occ[j]].t; break; }
        }
        F[p] = f;
        j = e[i].o;
        while (occ[j] != NONE && e[occ[j]].t + W >= p) ++j;
        occ[j] = i;
    }
}

static uint64_t st_big, st_subseg, st_walk, st_walkn; static uint64_t g_n, g_m, g_h, g_parts, g_groups, g_groups1, g_longwalk, g_maxwalk;
/* two-phase sweeps; entries sorted by time */
static void sim_big(const uint8_t* d, uint32_t n, ent_t* e, uint32_t m, uint32_t* F) {
    uint32_t* e1 = (uint32_t*)malloc(4 * m); uint32_t* e2 = (uint32_t*)malloc(4 * m); uint32_t* slot_of = (uint32_t*)malloc(4 * m);
    uint8_t* ishome = (uint8_t*)calloc(m + 1, 1); uint8_t* placed = (uint8_t*)calloc(m, 1);
    uint32_t* pool = (uint32_t*)malloc(4 * m);
    for (uint32_t j = 0; j < m; ++j) { e1[j] = NONE; e2[j] = NONE; }
    uint32_t n1 = 0; while (n1 < m && e[n1].t <= W) ++n1;   /* phase-1 entries: [0, n1) */
    for (uint32_t i = 0; i < m; ++i) ishome[e[i].o] = 1;
    ++st_big; int giant = m > 256; uint64_t parts=0, groups=0, groups1=0, H=0; for (uint32_t j = 0; j < m; ++j) H += ishome[j];
    /* phase 1: per home segment, the earliest active unplaced entries in time order */
    for (uint32_t x = 0; x < m;) {
        uint32_t xe = x + 1; while (xe < m && !ishome[xe]) ++xe;
        groups1 += (xe - x + 31) / 32;
        uint32_t j = x;
        for (uint32_t i = 0; i < n1 && j < xe; ++i) if (!placed[i] && e[i].o <= x) { e1[j] = i; slot_of[i] = j; placed[i] = 1; ++j; }
        x = xe;
    }
    /* phase 2: sub-segments (used part with increasing release times, then never-used part) */
    for (uint32_t x = 0; x < m;) {
        uint32_t xe = x + 1; while (xe < m && !ishome[xe]) ++xe;
        uint32_t used = x; while (used < xe && e1[used] != NONE) ++used;
        for (int part = 0; part < 2; ++part) {
            const uint32_t lo = part ? used : x, hi = part ? xe : used;
            if (lo >= hi) continue;
            ++st_subseg; ++parts; groups += (hi - lo + 31) / 32;
            uint32_t np = 0;
            for (uint32_t i = n1; i < m; ++i) if (!placed[i] && e[i].o <= x) pool[np++] = i;
            int64_t prev = -1;
            for (uint32_t j = lo; j < hi; ++j) {
                const uint32_t r = part ? 0 : e[e1[j]].t + W + 1;
                uint32_t lb = 0; { uint32_t a = 0, b = np; while (a < b) { uint32_t mid = (a + b) / 2; if (e[pool[mid]].t >= r) b = mid; else a = mid + 1; } lb = a; }
                int64_t idx = prev + 1 > (int64_t)lb ? prev + 1 : (int64_t)lb;
                if (idx < (int64_t)np) { const uint32_t i = pool[idx]; e2[j] = i; slot_of[i] = j; placed[i] = 1; }
                prev = idx;
            }
        }
        x = xe;
    }
    if (giant) { ++g_n; g_m += m; g_h += H; g_parts += parts; g_groups += groups; g_groups1 += groups1; }
    for (uint32_t i = 0; i < m; ++i) if (!placed[i]) { printf("UNPLACED entry in big cluster m=%u\n", m); exit(4); }
    /* finds: scan [home, own slot) */
    for (uint32_t i = 0; i < m; ++i) {
        const uint32_t p = e[i].t, w = word_at(d, n, p);
        uint32_t f = NONE;
        ++st_walkn;
        if (giant && slot_of[i] - e[i].o > 6) { ++g_longwalk; }
        for (uint32_t j = e[i].o; j < slot_of[i]; ++j) {
            ++st_walk; if (giant) ++g_maxwalk;
            uint32_t oc;
            if (e1[j] != NONE && e[e1[j]].t < p && e[e1[j]].t + W >= p) oc = e1[j];
            else { oc = e2[j]; if (oc == NONE || e[oc].t >= p) { printf("walk hit a dead slot: m=%u i=%u j=%u\n", m, i, j); exit(5); } }
            if (word_at(d, n, e[oc].t) == w) { f = e[oc].t; break; }
        }
        F[p] = f;
    }
    free(e1); free(e2); free(slot_of); free(ishome); free(placed); free(pool);
}

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: proto_v4 file block [small_max]\n"); return 2; }
    FILE* fp = fopen(argv[1], "rb"); if (!fp) { perror(argv[1]); return 2; }
    fseek(fp, 0, SEEK_END); long sz = ftell(fp); fseek(fp, 0, SEEK_SET);
    uint8_t* data = (uint8_t*)malloc(sz); if (fread(data, 1, sz, fp) != (size_t)sz) return 2; fclose(fp);
    uint32_t block = (uint32_t)atoi(argv[2]);
    if (argc > 3) SMALL_MAX = atoi(argv[3]);
    uint64_t mism = 0, checked = 0, nblk = 0, fallback = 0;
    for (long off = 0; off < sz; off += block) {
        uint32_t n = (uint32_t)(sz - off < block ? sz - off : block);
        const uint8_t* d = data + off;
        uint32_t* F = (uint32_t*)malloc(4 * n); uint32_t* Fr = (uint32_t*)malloc(4 * n);
        for (uint32_t i = 0; i < n; ++i) F[i] = 0xFFFFFFFDu;
        ref_block(d, n, 1, Fr);
        /* occupancy + claims in a scrambled order */
        uint8_t* occ = (uint8_t*)calloc(SLOTS + GUARDB, 1);
        uint32_t* claim = (uint32_t*)malloc(4 * n); uint32_t* hs = (uint32_t*)malloc(4 * n);
        uint32_t* at = (uint32_t*)malloc(4 * (SLOTS + GUARDB)); /* position claimed at slot */
        int wrap = 0;
        for (uint32_t k = 0; k < n; ++k) {
            uint32_t i = (uint32_t)(((uint64_t)k * 40503u + 12345u) % n);   /* 40503 odd; n power of two in the common case, else fall back */
            if ((n & (n - 1)) != 0) i = k;
            uint32_t h = port_lz77_hash(word_at(d, n, i)); hs[i] = h;
            uint32_t s = h; while (occ[s]) { ++s; if (s == SLOTS) { s = 0; wrap = 1; } }
            occ[s] = 1; claim[i] = s; at[s] = i;
        }
        int special = occ[0] || occ[SLOTS - 1] || wrap;
        if (special) { ++fallback; }
        else {
            ent_t* e = (ent_t*)malloc(sizeof(ent_t) * n);
            for (uint32_t s = 0; s < SLOTS;) {
                if (!occ[s]) { ++s; continue; }
                uint32_t en = s; while (occ[en]) ++en;
                uint32_t m = en - s;
                for (uint32_t j = 0; j < m; ++j) { e[j].t = at[s + j]; e[j].o = hs[at[s + j]] - s; }
                if (m == 1) F[e[0].t] = NONE;
                else {
                    qsort(e, m, sizeof(ent_t), cmp_ent);
                    if ((int)m <= SMALL_MAX) sim_small(d, n, e, m, F); else sim_big(d, n, e, m, F);
                }
                s = en;
            }
            free(e);
            for (uint32_t i = 0; i < n; ++i) { ++checked; if (Fr[i] != F[i]) { if (mism < 8) printf("block@%ld pos %u: ref %u proto %u\n", off, i, Fr[i], F[i]); ++mism; } }
        }
        free(F); free(Fr); free(occ); free(claim); free(hs); free(at); ++nblk;
    }
    printf("%s block %u small_max %d: blocks %lu (fallback %lu), positions %lu, mismatches %lu | big clusters/blk %.1f subseg/blk %.1f walks/blk %.0f walk steps/blk %.0f\n",
           argv[1], block, SMALL_MAX, nblk, fallback, checked, mism, (double)st_big / nblk, (double)st_subseg / nblk, (double)st_walkn / nblk, (double)st_walk / nblk);
    printf("giant(>256): n/blk %.2f avg m %.0f homes %.1f parts %.1f groups(ph2) %.1f groups(ph1) %.1f walks>6 per cluster %.1f walk steps per cluster %.0f\n", (double)g_n/nblk, (double)g_m/g_n, (double)g_h/g_n, (double)g_parts/g_n, (double)g_groups/g_n, (double)g_groups1/g_n, (double)g_longwalk/g_n, (double)g_maxwalk/g_n);
    return mism != 0;
}
