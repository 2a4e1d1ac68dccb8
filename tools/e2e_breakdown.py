import sys, time, ctypes as C, numpy as np, torch
sys.path.insert(0, '.')
from compression_algorithms_b200 import _lib, corpus, device as dv
ctx = dv.Context(0); lib = _lib.core()
n = 1_000_000_000; block = 65536; nb = (n + block - 1) // block
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); corpus.generate(n, 0, 20261018, out=h_in.numpy())
d = torch.empty(n, dtype=torch.uint8, device='cuda')
def t(fn, reps=3):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3
print("H2D 1 GB pinned: %.1f ms" % t(lambda: d.copy_(h_in, non_blocking=True)))
h_back = torch.empty(n, dtype=torch.uint8).pin_memory()
print("D2H 1 GB pinned: %.1f ms" % t(lambda: h_back.copy_(d, non_blocking=True)))
s2 = torch.cuda.Stream()
def both():
    d.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2): h_back.copy_(d, non_blocking=True)
print("H2D + D2H concurrently: %.1f ms" % t(both))
cap = int(lib.b200_lz77_max_bytes(1, n, block))
h_out = torch.empty(cap, dtype=torch.uint8).pin_memory(); sizes = torch.empty(nb, dtype=torch.int64).pin_memory(); off = torch.empty(nb + 1, dtype=torch.int64).pin_memory()
h_dec = torch.empty(n, dtype=torch.uint8).pin_memory(); tot = C.c_uint64(0)
comp = lambda: _lib.check(lib.b200_lz77_compress_host(ctx.handle, 1, h_in.data_ptr(), n, block, h_out.data_ptr(), cap, sizes.data_ptr(), off.data_ptr(), C.byref(tot)))
dec = lambda: _lib.check(lib.b200_lz77_decompress_host(ctx.handle, 1, h_out.data_ptr(), tot.value, off.data_ptr(), sizes.data_ptr(), n, block, h_dec.data_ptr()))
print("compress_host: %.1f ms" % t(comp)); print("decompress_host: %.1f ms (stream %.2f GB)" % (t(dec), tot.value / 1e9))
st = dv.lz77_alloc(ctx, n, block, 1); d.copy_(h_in)
print("device encode: %.1f ms" % t(lambda: dv.lz77_encode(ctx, d, 1, block, stream=st, sync=False)))
st = dv.lz77_encode(ctx, d, 1, block, stream=st); o = torch.empty_like(d)
print("device decode: %.1f ms" % t(lambda: dv.lz77_decode(ctx, st, out=o)))
