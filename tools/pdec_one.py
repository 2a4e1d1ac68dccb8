import os, sys, torch
sys.path.insert(0, '.')
from compression_algorithms_b200 import corpus, device as dv
ctx = dv.Context(0)
n = 256 << 20
d = torch.from_numpy(corpus.generate(n, 0, 7)).to(ctx.device)
st = dv.lz77_encode(ctx, d, int(sys.argv[1]) if len(sys.argv) > 1 else 1, 1 << 20)
out = torch.zeros(n, dtype=torch.uint8, device=ctx.device)
os.environ["B200_LZ_PDEC"] = "1"
dv.lz77_decode(ctx, st, out=out); dv.lz77_decode(ctx, st, out=out)
torch.cuda.synchronize(); print(bool(torch.equal(out, d)))
