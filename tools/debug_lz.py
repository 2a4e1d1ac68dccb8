"""Compares the v2 match finder's token candidates with the oracle port's find results at
every token start (port F array) and reports the first differences per block."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from compression_algorithms_b200 import corpus, device as dv
from oracle import bindings as ob

ctx = dv.Context(0)
def run(name, data, variant, block):
    n = data.size
    d = torch.from_numpy(data.copy()).to(ctx.device)
    st, tok = dv.lz77_encode_debug(ctx, d, variant, block)
    tok = tok.cpu().numpy().view(np.uint32)
    bs = n if block == 0 or block > n else block
    W = 32768 if variant else 16384
    MAXLEN = 31 if variant else 15
    nbad = 0
    for b in range((n + bs - 1) // bs):
        blk = data[b*bs:(b+1)*bs]
        if variant: _, F = ob.port_deflate_lz77_compress(blk, want_F=True)
        else: _, _, F = ob.port_lz77_compress(blk, want_F=True)
        starts = np.nonzero(F != 0xFFFFFFFE)[0]
        for p in starts:
            m = int(F[p]); t = int(tok[b, p])
            rej = (m == 0xFFFFFFFF) or ((p - m >= W - 1) if variant else (p - m == W))
            if rej: want = 0
            else:
                l = 4
                pad = np.concatenate([blk, np.zeros(64, dtype=np.uint8)])
                while l < MAXLEN and pad[m + l] == pad[p + l]: l += 1
                want = (p - m) | (l << 16)
            if want != t:
                nbad += 1
                if nbad <= 5: print("  %s v%d block %d pos %d: want off %d len %d got off %d len %d" % (name, variant, b, p, want & 0xFFFF, want >> 16, t & 0xFFFF, t >> 16))
    # stream parity
    exp_out, exp_sizes = ob.port_lz77_compress_blocks(data, bs, variant)
    sizes = st.block_sizes.cpu().numpy().astype(np.uint64)
    dec = dv.lz77_decode(ctx, st).cpu().numpy()
    print("%s variant %d block %d: tok mismatches at token starts %d, sizes equal %s, roundtrip %s" % (name, variant, block, nbad, np.array_equal(sizes, exp_sizes), np.array_equal(dec, data)), flush=True)

for variant in (1, 0):
    run("enwik", corpus.generate(200000, 0, 5), variant, 65536)
    run("acgt", corpus.generate(150000, 1, 5), variant, 65536)
    run("rand", corpus.generate(70000, 3, 5), variant, 65536)
    run("tiny", corpus.generate(1000, 0, 3), variant, 0)
