"""B200-native (sm_100a) block-parallel LZ77 / Huffman / FSE / deflate codecs behind
the C entry points of jdm365/Compression_Algorithms.

Layout:
  csrc/        hand-written CUDA kernels + the C-ABI (include/b200comp.h)
  device.py    device-resident Python mirror of the C-ABI (torch = memory + streams)
  csrc/shims/  the reference's own function names (one C library per reference directory)
  corpus.py    seeded synthetic inputs
  sharding.py  block-range sharding across ranks + all-gather of shard sizes
"""
__all__ = ["device", "corpus"]
