/*
 * TEST INFRASTRUCTURE ONLY (oracle). Nothing under compression_algorithms_b200/
 * may include, link or call this.
 *
 * Per-thread bump arena that the UNMODIFIED reference sources are redirected to
 * with -Dmalloc=orc_malloc -Drealloc=orc_realloc -Dfree=orc_free -Dexit=orc_exit
 * on the compiler command line (lz77_compress leaks its
 * 24 MiB table on every call, /root/reference/algorithms/lz77/lz77.c:278-344;
 * huffman_compress leaks the tree, algorithms/huffman/huffman.c:295-327).
 * exit(1) in the reference (huffman.c:278-281, lz77.c:315-326) becomes a longjmp
 * back into the harness so a test can observe the error instead of dying.
 */
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <setjmp.h>
#include <stdlib.h>
#include <string.h>

static __thread uint8_t* orc_arena_base = NULL;
static __thread size_t   orc_arena_cap  = 0;
static __thread size_t   orc_arena_off  = 0;
static __thread jmp_buf  orc_exit_jmp;
static __thread int      orc_exit_armed = 0;

/* real libc allocators: this header is only included by harness files that are
 * compiled WITHOUT the -D redirections. */
static void orc_arena_reserve(size_t bytes) {
    bytes += 4096;
    if (orc_arena_cap < bytes) {
        free(orc_arena_base);
        orc_arena_base = (uint8_t*)malloc(bytes);
        orc_arena_cap  = bytes;
    }
    orc_arena_off = 0;
}

void* orc_malloc(size_t n) {
    size_t off = (orc_arena_off + 63) & ~(size_t)63;
    if (off + n + 16 > orc_arena_cap) {
        /* arena too small: fall back to libc so the reference still runs */
        uint8_t* p = (uint8_t*)malloc(n + 16);
        *(size_t*)p = n;
        return p + 16;
    }
    *(size_t*)(orc_arena_base + off) = n;
    orc_arena_off = off + 16 + n;
    return orc_arena_base + off + 16;
}

void* orc_realloc(void* p, size_t n) {
    if (!p) return orc_malloc(n);
    size_t old = *(size_t*)((uint8_t*)p - 16);
    if (n <= old) return p; /* shrink in place: bytes past n stay readable (U4) */
    void* q = orc_malloc(n);
    memcpy(q, p, old);
    return q;
}

void orc_free(void* p) { (void)p; }

void orc_exit(int code) {
    if (orc_exit_armed) longjmp(orc_exit_jmp, code ? code : 1);
    _Exit(code);
}
