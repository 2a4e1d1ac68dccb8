import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from compression_algorithms_b200 import corpus, device as dv
from oracle import bindings as ob
ctx = dv.Context(0)
n = 6_000_000
data = corpus.generate(n, 0, 7)
d = torch.from_numpy(data).to(ctx.device)
for variant in (0, 1):
    t0 = time.time()
    st = dv.lz77_encode(ctx, d, variant, 0)
    t1 = time.time()
    exp, sizes = ob.port_lz77_compress_blocks(data, n, variant)
    out = st.out[: st.total_bytes].cpu().numpy()
    nb = int(sizes[0]) if variant else int(sizes[0]) // 8
    print("variant", variant, "whole-buffer 6MB gpu_s %.2f" % (t1 - t0), "equal", np.array_equal(out[:nb], exp[0, :nb]), int(st.block_sizes[0]), int(sizes[0]))
    print(" roundtrip", torch.equal(dv.lz77_decode(ctx, st), d))
