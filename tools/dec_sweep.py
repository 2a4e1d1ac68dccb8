import os, sys, torch
sys.path.insert(0, '.')
from compression_algorithms_b200 import corpus, device as dv
ctx = dv.Context(0)
n = 300_000_000
d = torch.from_numpy(corpus.generate(n, 0, 5)).to(ctx.device)
for variant in (1, 0):
    st = dv.lz77_encode(ctx, d, variant, 65536)
    out = torch.empty_like(d)
    for G in (8, 16, 32):
        os.environ["B200_LZ_DEC_G"] = str(G)
        dv.lz77_decode(ctx, st, out=out); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3): dv.lz77_decode(ctx, st, out=out)
        b.record(); torch.cuda.synchronize()
        print("variant", variant, "G", G, "decode ms/GB %.2f" % (a.elapsed_time(b) / 3 / n * 1e9), "ok", bool(torch.equal(out, d)))
