"""Runs every codec once each way on a 100 MB enwik-shaped buffer (for ncu launch lists / captures):
Huffman whole-buffer and per block, FSE, deflate-variant LZ77 + the token entropy stage + both decoders."""
import sys, torch
sys.path.insert(0, '.')
from compression_algorithms_b200 import corpus, device as dv
ctx = dv.Context(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
d = torch.from_numpy(corpus.generate(n, 0, 5)).to(ctx.device)
for rep in range(reps):
    for block in (0, 65536):
        st = dv.huffman_encode(ctx, d, block)
        out = dv.huffman_decode(ctx, st)
        assert torch.equal(out, d)
    fs = dv.fse_encode(ctx, d, 65536, 1024)
    out = dv.fse_decode(ctx, fs)
    assert torch.equal(out, d)
    ds = dv.deflate_compress(ctx, d, 65536)
    out = dv.deflate_decompress(ctx, ds)
    assert torch.equal(out, d)
torch.cuda.synchronize()
print("ok")
