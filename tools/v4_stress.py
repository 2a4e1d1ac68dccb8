"""lz77_v4_kernel against lz77_v2_kernel, byte for byte, over many seeds / corpus kinds / block sizes / ragged lengths:
python tools/v4_stress.py [rounds]"""
import os, sys, numpy as np, torch
sys.path.insert(0, '.')
from compression_algorithms_b200 import corpus, device as dv
ctx = dv.Context(0)
rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(12345)
tot = 0; bad = 0
for r in range(rounds):
    kind = int(rng.integers(0, 4))
    n = int(rng.integers(1, 48_000_000)) if kind != 2 else int(rng.integers(1, 3_000_000))
    block = int(rng.choice([65536, 65536, 65536, 32768, 16384, 40000, 4096]))
    seed = int(rng.integers(1, 1 << 30))
    data = corpus.generate(n, kind, seed)
    if rng.random() < 0.3:      # splice in runs and periodic stretches
        for _ in range(20):
            a = int(rng.integers(0, max(1, n - 70000))); L = int(rng.integers(10, 70000)); per = int(rng.integers(1, 300))
            data[a:a + L] = np.resize(data[a:a + per].copy(), min(L, n - a))
    d = torch.from_numpy(data).to(ctx.device)
    os.environ["B200_LZ_V4"] = "0"
    try:
        a = dv.lz77_encode(ctx, d, 1, block)
    except Exception:
        print("FAILED (lz77_v2_kernel call) round %d kind %d n %d block %d seed %d" % (r, kind, n, block, seed), flush=True)
        raise
    out_a = a.out[: a.total_bytes].clone(); sz_a = a.block_sizes.clone()
    os.environ["B200_LZ_V4"] = "1"
    try:
        b = dv.lz77_encode(ctx, d, 1, block)
    except Exception:
        print("FAILED round %d kind %d n %d block %d seed %d" % (r, kind, n, block, seed), flush=True)
        raise
    ok = a.total_bytes == b.total_bytes and bool(torch.equal(sz_a, b.block_sizes)) and bool(torch.equal(out_a, b.out[: b.total_bytes]))
    tot += n; bad += 0 if ok else 1
    if not ok: print("MISMATCH kind %d n %d block %d seed %d" % (kind, n, block, seed))
os.environ.pop("B200_LZ_V4", None)
print("v4 == v2 on %d inputs, %.2f GB in total: %d mismatches" % (rounds, tot / 1e9, bad))
