import os, sys, numpy as np, torch
os.environ["B200_LZ_V4"] = "0"   # phase statistics of lz77_v2_kernel
sys.path.insert(0, '.')
from compression_algorithms_b200 import corpus, device as dv
ctx = dv.Context(0)
kind = int(sys.argv[1]) if len(sys.argv) > 1 else 0
mult = int(sys.argv[2]) if len(sys.argv) > 2 else 1   # blocks per CTA; statistics are taken over the last 148 blocks
data = corpus.generate(148 * mult * 65536, kind, 5)
d = torch.from_numpy(data).to(ctx.device)
for variant in (1,):
    st, tok = dv.lz77_encode_debug(ctx, d, variant, 65536)
    st, tok = dv.lz77_encode_debug(ctx, d, variant, 65536)
    s = st.debug_stats.cpu().numpy().astype(np.int64)[-148:]
    ph = s[:, :7]
    names = ["P0 load", "P1 bitmap", "P2 prefix", "P3 partition", "P4 sim", "P5 parse", "P6 emit"]
    prev = np.zeros(len(s), dtype=np.int64)
    print("variant", variant, "blocks", len(s), "total cycles/block median", int(np.median(ph[:, 6])))
    for k, nm in enumerate(names):
        dtk = ph[:, k] - prev; prev = ph[:, k]
        print("  %-13s median %8d  p90 %8d  max %8d cycles" % (nm, np.median(dtk), np.percentile(dtk, 90), dtk.max()))
    w = s[:, 8:].reshape(len(s), 32, 4)
    cyc = w[:, :, 0]
    print("  P4 per-warp cycles: median of max %d, median of median %d, median of min %d" % (np.median(cyc.max(1)), np.median(np.median(cyc, 1)), np.median(cyc.min(1))))
    b = 3
    order = np.argsort(-cyc[b])
    print("  block %d warps:" % b)
    for wi in list(order[:8]) + list(order[-2:]):
        r = w[b, wi]
        ent, rounds = r[1] & 0xFFFF, r[2] & 0xFFFF
        print("    warp %2d: cycles %7d entries %4d rounds %4d (%.1f per batch, %d cycles each) walks %d t_walk %dK t_sync %dK group steps %d" % (
            wi, r[0], ent, rounds, rounds / max(1.0, ent / 32.0), r[0] // max(1, rounds), r[3] & 0xFFFF, r[2] >> 16, r[3] >> 16, r[1] >> 16))
    print("  max over warps: t_commit %dK t_matchlen %dK" % (s[b, 7] >> 16, s[b, 7] & 0xFFFF))
