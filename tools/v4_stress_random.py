"""lz77_v4_kernel on near-random input (the clusters on slot 0 / the table end are frequent there), repeated:
python tools/v4_stress_random.py [rounds]"""
import os, sys, numpy as np, torch
sys.path.insert(0, '.')
from compression_algorithms_b200 import corpus, device as dv
ctx = dv.Context(0)
rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 30
rng = np.random.default_rng(777)
bad = 0
for r in range(rounds):
    n = int(rng.integers(20_000_000, 40_000_000)); seed = int(rng.integers(1, 1 << 30))
    d = torch.from_numpy(corpus.generate(n, 3, seed)).to(ctx.device)
    os.environ["B200_LZ_V4"] = "0"
    a = dv.lz77_encode(ctx, d, 1, 65536)
    out_a = a.out[: a.total_bytes].clone(); sz_a = a.block_sizes.clone()
    os.environ["B200_LZ_V4"] = "1"
    try:
        b = dv.lz77_encode(ctx, d, 1, 65536)
    except Exception as e:
        print("FAILED round %d n %d seed %d: %s" % (r, n, seed, str(e)[-80:]), flush=True)
        sys.exit(1)
    ok = a.total_bytes == b.total_bytes and bool(torch.equal(sz_a, b.block_sizes)) and bool(torch.equal(out_a, b.out[: b.total_bytes]))
    bad += 0 if ok else 1
    if not ok: print("MISMATCH n %d seed %d" % (n, seed))
print("v4 == v2 on %d near-random inputs: %d mismatches" % (rounds, bad))
