"""Phase cycles of lz77_v3_kernel (debug instantiation): B200_LZ_V3=1 python tools/lz_stats3.py [kind] [blocks per CTA] [variant]"""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from compression_algorithms_b200 import corpus, device as dv
ctx = dv.Context(0)
kind = int(sys.argv[1]) if len(sys.argv) > 1 else 0
mult = int(sys.argv[2]) if len(sys.argv) > 2 else 1
variant = int(sys.argv[3]) if len(sys.argv) > 3 else 1
data = corpus.generate(148 * mult * 65536, kind, 5)
d = torch.from_numpy(data).to(ctx.device)
st, tok = dv.lz77_encode_debug(ctx, d, variant, 65536)
st, tok = dv.lz77_encode_debug(ctx, d, variant, 65536)
s = st.debug_stats.cpu().numpy().astype(np.int64)[-148:]
ph = s[:, :8]
names = ["P0+P1 bitmap", "P2 prefix", "P3 classify", "P4 marks", "P5 jump", "P6 partition", "P7 simulate", "P8+P9 parse/emit"]
prev = np.zeros(len(s), dtype=np.int64)
print("variant", variant, "kind", kind, "blocks", len(s), "total cycles/block median", int(np.median(ph[:, 7])))
for k, nm in enumerate(names):
    dtk = ph[:, k] - prev; prev = ph[:, k]
    print("  %-17s median %8d  p90 %8d  max %8d cycles" % (nm, np.median(dtk), np.percentile(dtk, 90), dtk.max()))
w = s[:, 8:].reshape(len(s), 32, 4)
print("  P7 per warp: total median of max %d, median %d | stage A median of max %d, median %d | long lists per block median %d" % (
    np.median(w[:, :, 0].max(1)), np.median(np.median(w[:, :, 0], 1)), np.median(w[:, :, 2].max(1)), np.median(np.median(w[:, :, 2], 1)), np.median(w[:, 0, 3])))
