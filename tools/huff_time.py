"""Huffman / FSE / entropy-stage timings at 100 MB (device resident, CUDA events)."""
import sys, torch
sys.path.insert(0, '.')
from compression_algorithms_b200 import corpus, device as dv
ctx = dv.Context(0)
n = 100_000_000
d = torch.from_numpy(corpus.generate(n, 0, 20261018)).to(ctx.device)
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
for block in (0, 65536):
    st = dv.huffman_alloc(ctx, n, block)
    te = timed(lambda: dv.huffman_encode(ctx, d, block, stream=st, sync=False))
    st = dv.huffman_encode(ctx, d, block, stream=st)
    out = torch.empty_like(d)
    td = timed(lambda: dv.huffman_decode(ctx, st, out=out))
    print("huffman block %d: encode %.3f ms (%.0f GB/s) decode %.3f ms (%.0f GB/s) ok %s" % (block, te, n / te / 1e6, td, n / td / 1e6, bool(torch.equal(out, d))))
ctx.set_timing(True)
st = dv.huffman_encode(ctx, d, 0, stream=dv.huffman_alloc(ctx, n, 0)); dv.huffman_decode(ctx, st)
print("kernel timings (kind, ms):", [(k, round(m, 3)) for k, m in ctx.timings()])
