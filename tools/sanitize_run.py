"""Small invocations of every codec for compute-sanitizer (memcheck / racecheck / initcheck):
   compute-sanitizer --tool memcheck python tools/sanitize_run.py"""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
from compression_algorithms_b200 import corpus, device as dv
ctx = dv.Context(0)
for kind, n in ((0, 3 * 65536 + 123), (1, 70000), (3, 65536)):
    data = corpus.generate(n, kind, 1)
    d = torch.from_numpy(data.copy()).to(ctx.device)
    for variant in (1, 0):
        st = dv.lz77_encode(ctx, d, variant, 65536)
        assert torch.equal(dv.lz77_decode(ctx, st), d)
    st = dv.lz77_encode(ctx, d, 1, 0)      # one block above 64 KiB: the sliced path
    assert torch.equal(dv.lz77_decode(ctx, st), d)
    for block in (0, 65536):
        hs = dv.huffman_encode(ctx, d, block)
        assert torch.equal(dv.huffman_decode(ctx, hs), d)
    fs = dv.fse_encode(ctx, d, 65536, dv.DEFAULT_FSE_SEG)
    assert torch.equal(dv.fse_decode(ctx, fs), d)
    ds = dv.deflate_compress(ctx, d, 65536)
    assert torch.equal(dv.deflate_decompress(ctx, ds), d)
ctx.close()
print("sanitize_run ok")
