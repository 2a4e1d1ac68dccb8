"""lz77_v2_kernel (B200_LZ_V4=0), lz77_v4_kernel (B200_LZ_V4=1) and the default (a sample of the input decides) on the four
corpus kinds: python tools/v4_compare.py [MB]"""
import os, sys, torch
sys.path.insert(0, '.')
from compression_algorithms_b200 import corpus, device as dv
ctx = dv.Context(0)
n = (int(sys.argv[1]) if len(sys.argv) > 1 else 200) * 1_000_000
for kind, name in ((0, "enwik"), (1, "acgt"), (2, "skewed"), (3, "random")):
    d = torch.from_numpy(corpus.generate(n, kind, 7)).to(ctx.device)
    st = dv.lz77_alloc(ctx, n, 65536, 1)
    res = {}
    for mode in ("v2", "v4", "auto"):
        if mode == "v4": os.environ["B200_LZ_V4"] = "1"
        elif mode == "v2": os.environ["B200_LZ_V4"] = "0"
        else: os.environ.pop("B200_LZ_V4", None)
        for _ in range(2): dv.lz77_encode(ctx, d, 1, 65536, stream=st, sync=False)
        ctx.sync(); torch.cuda.synchronize()
        ts = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); dv.lz77_encode(ctx, d, 1, 65536, stream=st, sync=False); e1.record()
            ctx.sync(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        res[mode] = min(ts)
    os.environ.pop("B200_LZ_V4", None)
    print("%-7s %d MB: v2 %.2f ms (%.1f GB/s)  v4 %.2f ms (%.1f GB/s)  default %.2f ms (%.1f GB/s)" % (name, n // 1_000_000, res["v2"], n / res["v2"] / 1e6, res["v4"], n / res["v4"] / 1e6, res["auto"], n / res["auto"] / 1e6))
