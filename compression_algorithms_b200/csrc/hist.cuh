// Byte histogram shared by the Huffman and FSE front ends.
// Replaces the counting loops of algorithms/huffman/huffman.c:184-187 and
// algorithms/fse/src/main.zig:88-96.
#pragma once
#include "common.cuh"

constexpr uint32_t HIST_TILE = 65536;  // bytes per CTA

// One CTA per 64 KiB tile of one block; 16-byte loads; one private 256-bin
// histogram per warp in shared memory, merged into the block's global histogram.
static __global__ void __launch_bounds__(256) byte_hist_kernel(const uint8_t* __restrict__ in, uint64_t n, uint64_t bs,
                                                        uint32_t tiles_per_block, uint32_t* __restrict__ freq) {
    __shared__ uint32_t h[8][256];
    for (int i = threadIdx.x; i < 8 * 256; i += 256) (&h[0][0])[i] = 0;
    __syncthreads();
    const uint64_t b = blockIdx.x / tiles_per_block, k = blockIdx.x % tiles_per_block;
    const uint64_t start = b * bs + k * (uint64_t)(HIST_TILE);
    uint64_t end = start + HIST_TILE;
    if (end > (b + 1) * bs) end = (b + 1) * bs;
    if (end > n) end = n;
    uint32_t* my = h[threadIdx.x >> 5];
    for (uint64_t i = start + (uint64_t)threadIdx.x * 16; i < end; i += 256 * 16) {
        if (i + 16 <= end) {
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(in + i));
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                atomicAdd(&my[w[q] & 0xFF], 1u);
                atomicAdd(&my[(w[q] >> 8) & 0xFF], 1u);
                atomicAdd(&my[(w[q] >> 16) & 0xFF], 1u);
                atomicAdd(&my[w[q] >> 24], 1u);
            }
        } else {
            for (uint64_t j = i; j < end; ++j) atomicAdd(&my[in[j]], 1u);
        }
    }
    __syncthreads();
    uint32_t s = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += h[w][threadIdx.x];
    if (s) atomicAdd(&freq[b * 256 + threadIdx.x], s);
}

