"""Runs the Huffman and FSE codecs once each way on a 100 MB enwik-shaped buffer (for ncu launch lists)."""
import sys, torch
sys.path.insert(0, '.')
from compression_algorithms_b200 import corpus, device as dv
ctx = dv.Context(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
d = torch.from_numpy(corpus.generate(n, 0, 5)).to(ctx.device)
for rep in range(2):
    for block in (0, 65536):
        st = dv.huffman_encode(ctx, d, block)
        out = dv.huffman_decode(ctx, st)
        assert torch.equal(out, d)
    fs = dv.fse_encode(ctx, d, 65536, 1024)
    out = dv.fse_decode(ctx, fs)
    assert torch.equal(out, d)
torch.cuda.synchronize()
print("ok")
