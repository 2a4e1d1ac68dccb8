/* Shaped like /root/reference/algorithms/deflate/main.c:6-16, linked against
 * libb200_deflate.so; additionally exercises decompress(), which the reference leaves empty. */
#include <stdio.h>
#include <time.h>
#include "b200_deflate.h"

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    clock_t start = clock();
    StateData state_data = compress(argv[1]);
    clock_t end = clock();
    printf("Compression took %f seconds\n", (double)(end - start) / CLOCKS_PER_SEC);
    printf("Wrote %s\n", state_data.compressed_filename);
    decompress(&state_data, state_data.compressed_filename);
    return 0;
}
