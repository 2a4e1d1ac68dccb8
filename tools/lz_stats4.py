"""Phase cycles of lz77_v4_kernel (debug instantiation): python tools/lz_stats4.py [kind] [blocks per CTA]"""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from compression_algorithms_b200 import corpus, device as dv
ctx = dv.Context(0)
kind = int(sys.argv[1]) if len(sys.argv) > 1 else 0
mult = int(sys.argv[2]) if len(sys.argv) > 2 else 1
data = corpus.generate(148 * mult * 65536, kind, 5)
d = torch.from_numpy(data).to(ctx.device)
st, tok = dv.lz77_encode_debug(ctx, d, 1, 65536)
st, tok = dv.lz77_encode_debug(ctx, d, 1, 65536)
s = st.debug_stats.cpu().numpy().astype(np.int64)[-148:]
ph = s[:, :7]
ok = ph[:, 6] > 0          # blocks handed back to v2 carry v2's stamps or none
names = ["P0 load", "P1 bitmap", "P2 prefix+flags", "P3 rank", "clusters", "P5 parse", "P6 emit"]
prev = np.zeros(len(s), dtype=np.int64)
print("kind", kind, "blocks", len(s), "total cycles/block median", int(np.median(ph[:, 6])))
for k, nm in enumerate(names):
    dtk = ph[:, k] - prev; prev = ph[:, k]
    print("  %-16s median %8d  p90 %8d  max %8d cycles" % (nm, np.median(dtk), np.percentile(dtk, 90), dtk.max()))
x = s[:, 8:16]
for k, nm in enumerate(["scatter", "work list", "lane tiers (thread 0)", "final stage: clusters above 64 entries", "chunks", "warp/lane clusters", "team clusters (above 256)", "wait at chunk end (thread 0)"]):
    print("  %-26s median %8d  max %8d" % (nm, np.median(x[:, k]), x[:, k].max()))
mx = s[:, 16]; print("  slowest team cluster: median %d cycles (m %d), max %d cycles (m %d); entries in team clusters median %d" % (np.median(mx >> 14) * 64, np.median(mx & 0x3FFF), (mx >> 14).max() * 64, (mx[np.argmax(mx >> 14)] & 0x3FFF), np.median(s[:, 17])))
tt = s[:, 24:32] * 64; print("  per-team busy cycles: median of max %d, median of mean %d" % (np.median(tt.max(1)), np.median(tt.mean(1))))
for bi in range(3):
    print("   block", bi, "slowest", (mx[bi] >> 14) * 64, "m", mx[bi] & 0x3FFF, "team busy", list(tt[bi]))
tb = s[:, 40:45] * 64
print("  team clusters, cycles summed over the block's clusters (median): sort %d, phase-1 sweep on warp 0 %d, then waiting for the pipelined phase 2 %d, finds %d" % (np.median(tb[:, 0]), np.median(tb[:, 3]), np.median(tb[:, 2]), np.median(tb[:, 4])))
ff = s[:, 50:54]
print("  final stage (thread 0): list + bitmap + prefix %d, scatter %d, own work %d, wait for the others %d" % tuple(np.median(ff[:, k]) for k in range(4)))
tw = s[:, 45:48] * 64
print("  team clusters, cycles after entry summed over the block's clusters (median): release ranks done (warp 2) %d, phase-2 sweep starts (warp 1) %d, ends %d" % tuple(np.median(tw[:, k]) for k in range(3)))
