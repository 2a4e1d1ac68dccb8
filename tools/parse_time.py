"""Device time of the parse kernel path (b200_lz77_encode_dev, deflate variant, 64 KiB blocks) on the 1 GB
enwik-shaped buffer of the headline config: python tools/parse_time.py [bytes] [reps]"""
import sys, torch
sys.path.insert(0, '.')
from compression_algorithms_b200 import corpus, device as dv
ctx = dv.Context(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
d = torch.from_numpy(corpus.generate(n, corpus.ENWIK, corpus.DEFAULT_SEED)).to(ctx.device)
st = dv.lz77_alloc(ctx, n, 65536, 1)
for _ in range(2):
    dv.lz77_encode(ctx, d, 1, 65536, stream=st, sync=False)
ctx.sync(); torch.cuda.synchronize()
ts = []
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); dv.lz77_encode(ctx, d, 1, 65536, stream=st, sync=False); e1.record()
    ctx.sync(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
print("encode_dev %d bytes: min %.2f ms, median %.2f ms" % (n, min(ts), sorted(ts)[len(ts) // 2]))
