"""LZ77 decode: one warp per block (default for small blocks) against the token-parallel decoder (lz77_pdec.cu), per
variant and block size: python tools/pdec_sweep.py [MB]"""
import os, sys, torch
sys.path.insert(0, '.')
from compression_algorithms_b200 import corpus, device as dv
ctx = dv.Context(0)
n = (int(sys.argv[1]) if len(sys.argv) > 1 else 64) << 20
d = torch.from_numpy(corpus.generate(n, 0, 7)).to(ctx.device)
def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
for variant in (0, 1):
    for block in (65536, 262144, 1 << 20, 4 << 20, 0):
        st = dv.lz77_encode(ctx, d, variant, block)
        out = torch.zeros(n, dtype=torch.uint8, device=ctx.device)
        res = []
        for mode in ("0", "1"):
            os.environ["B200_LZ_PDEC"] = mode
            if mode == "0" and (block == 0 or block >= (4 << 20)) and n > (64 << 20): res.append(float('nan')); res.append(None); continue
            t = timed(lambda: dv.lz77_decode(ctx, st, out=out))
            ok = bool(torch.equal(out, d))
            res.append(t); res.append(ok)
        os.environ.pop("B200_LZ_PDEC", None)
        print("variant %d block %8d: warp-per-block %8.2f ms (%6.2f GB/s, ok %s) | token-parallel %8.2f ms (%6.2f GB/s, ok %s)" % (
            variant, block, res[0], n / res[0] / 1e6, res[1], res[2], n / res[2] / 1e6, res[3]))
