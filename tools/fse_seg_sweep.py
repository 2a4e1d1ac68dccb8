"""FSE segment size sweep (100 MB enwik-shaped, 64 KiB table scopes): GB/s and ratio per segment size."""
import sys, torch
sys.path.insert(0, '.')
from compression_algorithms_b200 import corpus, device as dv
ctx = dv.Context(0)
n = 100_000_000
d = torch.from_numpy(corpus.generate(n, 0, corpus.DEFAULT_SEED)).to(ctx.device)
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): r = fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps / 1e3, r
for seg in (2048, 1024, 512, 256, 128):
    st = dv.fse_alloc(ctx, n, 65536, seg)
    te, st = timed(lambda: dv.fse_encode(ctx, d, 65536, seg, stream=st, sync=False))
    st = dv.fse_encode(ctx, d, 65536, seg, stream=st)
    dec = torch.empty_like(d)
    td, _ = timed(lambda: dv.fse_decode(ctx, st, out=dec, sync=False))
    print("seg %5d: compress %6.1f GB/s  decompress %6.1f GB/s  ratio %.4f  ok %s" % (seg, n / 1e9 / te, n / 1e9 / td, n / (st.total_words * 8.0), bool(torch.equal(dec, d))))
