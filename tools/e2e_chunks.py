"""Host-buffer deflate-variant LZ77 compress / decompress times (b200_lz77_*_host, pinned buffers, 1 GB) for the
pipeline chunk count given in B200_LZ_CHUNKS: B200_LZ_CHUNKS=16 python tools/e2e_chunks.py"""
import os, sys, time, ctypes as C, torch
sys.path.insert(0, '.')
from compression_algorithms_b200 import _lib, corpus, device as dv
ctx = dv.Context(0); lib = _lib.core()
n = 1_000_000_000; block = 65536; nb = (n + block - 1) // block
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); corpus.generate(n, 0, 20261018, out=h_in.numpy())
def t(fn, reps=3):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    return min(ts)
cap = int(lib.b200_lz77_max_bytes(1, n, block))
h_out = torch.empty(cap, dtype=torch.uint8).pin_memory(); sizes = torch.empty(nb, dtype=torch.int64).pin_memory(); off = torch.empty(nb + 1, dtype=torch.int64).pin_memory()
h_dec = torch.empty(n, dtype=torch.uint8).pin_memory(); tot = C.c_uint64(0)
comp = lambda: _lib.check(lib.b200_lz77_compress_host(ctx.handle, 1, h_in.data_ptr(), n, block, h_out.data_ptr(), cap, sizes.data_ptr(), off.data_ptr(), C.byref(tot)))
dec = lambda: _lib.check(lib.b200_lz77_decompress_host(ctx.handle, 1, h_out.data_ptr(), tot.value, off.data_ptr(), sizes.data_ptr(), n, block, h_dec.data_ptr()))
tc, td = t(comp), t(dec)
print("chunks %s: compress_host %.2f ms, decompress_host %.2f ms, sum %.2f ms -> %.2f GB/s, roundtrip %s" % (
    os.environ.get("B200_LZ_CHUNKS", "8"), tc, td, tc + td, n / 1e6 / (tc + td), bool(torch.equal(h_dec, h_in))))
