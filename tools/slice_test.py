import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from compression_algorithms_b200 import corpus, device as dv
from oracle import bindings as ob
ctx = dv.Context(0)
ok_all = True
for kind in (0, 1, 3):
    n = 700_000
    data = corpus.generate(n, kind, 7)
    d = torch.from_numpy(data).to(ctx.device)
    for variant in (1, 0):
        for block in (65536, 98304, 131072, 262144, 0):
            bs = n if block == 0 else block
            st = dv.lz77_encode(ctx, d, variant, block)
            exp, sizes = ob.port_lz77_compress_blocks(data, bs, variant)
            off = st.block_off.cpu().numpy(); out = st.out[: st.total_bytes].cpu().numpy(); bsz = st.block_sizes.cpu().numpy()
            good = True
            for b in range(len(sizes)):
                nb = int(sizes[b]) if variant else int(sizes[b]) // 8
                if int(bsz[b]) != int(sizes[b]) or not np.array_equal(out[off[b]: off[b] + nb], exp[b, :nb]):
                    good = False
                    a = out[off[b]: off[b] + nb]; e = exp[b, :nb]; m = min(len(a), len(e)); dd = np.nonzero(a[:m] != e[:m])[0]
                    print("  MISMATCH kind", kind, "variant", variant, "block", block, "blk", b, "sizes", int(bsz[b]), int(sizes[b]), "first diff", int(dd[0]) if dd.size else -1)
                    break
            rt = bool(torch.equal(dv.lz77_decode(ctx, st), d))
            ok_all &= good and rt
            print("kind", kind, "variant", variant, "block", block, "parity", good, "roundtrip", rt)
print("ALL OK" if ok_all else "FAILED")
