"""Deflate-variant decoders side by side: token-serial (B200_LZ_DEC_SERIAL=1) vs token-parallel."""
import os, sys, torch
sys.path.insert(0, '.')
from compression_algorithms_b200 import corpus, device as dv
ctx = dv.Context(0)
for n in (100_000_000, 1_000_000_000):
    for kind in (0, 1, 3):
        if n > 100_000_000 and kind != 0:
            continue
        d = torch.from_numpy(corpus.generate(n, kind, 5)).to(ctx.device)
        st = dv.lz77_encode(ctx, d, 1, 65536)
        out = torch.empty_like(d)
        for serial in ("1", "0"):
            os.environ["B200_LZ_DEC_SERIAL"] = serial
            out.zero_()
            dv.lz77_decode(ctx, st, out=out); torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(3): dv.lz77_decode(ctx, st, out=out)
            b.record(); torch.cuda.synchronize()
            print("n %d kind %d %s decode ms/GB %.2f ok %s" % (n, kind, "serial  " if serial == "1" else "parallel", a.elapsed_time(b) / 3 / n * 1e9, bool(torch.equal(out, d))))
        del d, st, out
