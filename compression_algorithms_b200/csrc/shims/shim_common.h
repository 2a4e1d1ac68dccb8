/* Shared by the reference-named shims: one lazily created GPU context per process. */
#ifndef B200_SHIM_COMMON_H
#define B200_SHIM_COMMON_H
#include <stdio.h>
#include <stdlib.h>
#include "b200comp.h"

static b200_ctx* shim_ctx(void) {
    static b200_ctx* ctx = NULL;
    if (!ctx) {
        const char* dev = getenv("B200_DEVICE");
        if (b200_ctx_create(&ctx, dev ? atoi(dev) : 0, NULL) != B200_OK) {
            /* the reference reports errors with printf + exit(1) (lz77.c:315-326, huffman.c:137-140) */
            printf("ERROR: %s\n", b200_last_error());
            exit(1);
        }
    }
    return ctx;
}
#define SHIM_CHECK(expr) do { if ((expr) != B200_OK) { printf("ERROR: %s\n", b200_last_error()); exit(1); } } while (0)
#endif
