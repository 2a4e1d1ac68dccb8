"""Deflate-variant decoder timing on each corpus kind: python tools/dec_time.py [MB] [block size]"""
import sys, torch
sys.path.insert(0, '.')
from compression_algorithms_b200 import corpus, device as dv
ctx = dv.Context(0)
mb = int(sys.argv[1]) if len(sys.argv) > 1 else 400
bs = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
for kind, name in enumerate(["enwik", "acgt", "skewed", "random"]):
    data = corpus.generate(mb * 1000000, kind, 5)
    d = torch.from_numpy(data).to(ctx.device)
    st = dv.lz77_encode(ctx, d, dv.LZ_DEFLATE, bs)
    out = torch.empty_like(d)
    for _ in range(2): dv.lz77_decode(ctx, st, out=out)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(5): dv.lz77_decode(ctx, st, out=out)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    print("%-7s %d MB: decode %.3f ms (%.1f GB/s out, stream %.2f B/B) ok=%s" % (name, mb, ms, mb / ms, int(st.total_bytes) / (mb * 1e6), bool(torch.equal(out, d))))
