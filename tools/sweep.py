"""BASELINE.json configs[4]: block-size sweep (64 KiB - 4 MiB) x {low-entropy, enwik-shaped, near-random}
x {deflate, Huffman, FSE}: compression ratio and device-resident GB/s (compress / decompress), round trip
checked. Prints one JSON object; `python tools/sweep.py > profiles/r01_block_sweep.json` on a B200."""
import json
import sys

import torch

sys.path.insert(0, '.')
from compression_algorithms_b200 import corpus, device as dv

N = int(sys.argv[1]) if len(sys.argv) > 1 else 64 * 1024 * 1024
ctx = dv.Context(0)


def timed(fn, reps=2):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        r = fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps / 1e3, r


out = {"bytes": N, "note": "device-resident, CUDA events, 2 repetitions after a warm-up; deflate blocks above 64 KiB are simulated in slices by one CTA per block; deflate_entropy = the same tokens through the entropy stage (times are LZ77 + stage)", "rows": []}
for kind, kname in ((corpus.ACGT, "low-entropy (acgt)"), (corpus.ENWIK, "enwik-shaped"), (corpus.RANDOM, "near-random")):
    d = torch.from_numpy(corpus.generate(N, kind, 7)).to(ctx.device)
    dec = torch.empty_like(d)
    for bs in (1 << 16, 1 << 17, 1 << 18, 1 << 19, 1 << 20, 1 << 21, 1 << 22):
        row = {"input": kname, "block": bs}
        st = dv.lz77_alloc(ctx, N, bs, dv.LZ_DEFLATE)
        tc, _ = timed(lambda: dv.lz77_encode(ctx, d, dv.LZ_DEFLATE, bs, stream=st, sync=False), reps=1)
        st = dv.lz77_encode(ctx, d, dv.LZ_DEFLATE, bs, stream=st)
        td, _ = timed(lambda: dv.lz77_decode(ctx, st, out=dec), reps=1)
        row["deflate"] = {"ratio": N / st.total_bytes, "compress_gbps": N / 1e9 / tc, "decompress_gbps": N / 1e9 / td, "ok": bool(torch.equal(dec, d))}
        # the same blocks with the token entropy stage (the reference's TODO, deflate/lz77.c:279)
        ds = dv.deflate_alloc(ctx, N, bs, lz=st)
        tce, _ = timed(lambda: dv.dfl_encode(ctx, st, stream=ds, sync=False), reps=1)
        ds = dv.dfl_encode(ctx, st, stream=ds)
        tok = torch.empty_like(st.out)
        tde, _ = timed(lambda: dv.dfl_decode(ctx, ds, tok), reps=1)
        tok_ok = bool(torch.equal(tok[: st.total_bytes], st.out[: st.total_bytes]))
        row["deflate_entropy"] = {"ratio": N / (ds.total_words * 4.0), "compress_gbps": N / 1e9 / (tc + tce), "decompress_gbps": N / 1e9 / (td + tde), "ok": tok_ok}
        del ds, tok
        hs = dv.huffman_alloc(ctx, N, bs)
        tc, _ = timed(lambda: dv.huffman_encode(ctx, d, bs, stream=hs, sync=False))
        hs = dv.huffman_encode(ctx, d, bs, stream=hs)
        td, _ = timed(lambda: dv.huffman_decode(ctx, hs, out=dec))
        row["huffman"] = {"ratio": N / (hs.total_words * 4.0), "compress_gbps": N / 1e9 / tc, "decompress_gbps": N / 1e9 / td, "ok": bool(torch.equal(dec, d))}
        fs = dv.fse_alloc(ctx, N, bs, dv.DEFAULT_FSE_SEG)
        tc, _ = timed(lambda: dv.fse_encode(ctx, d, bs, dv.DEFAULT_FSE_SEG, stream=fs, sync=False))
        fs = dv.fse_encode(ctx, d, bs, dv.DEFAULT_FSE_SEG, stream=fs)
        td, _ = timed(lambda: dv.fse_decode(ctx, fs, out=dec, sync=False))
        row["fse"] = {"ratio": N / (fs.total_words * 8.0), "compress_gbps": N / 1e9 / tc, "decompress_gbps": N / 1e9 / td, "ok": bool(torch.equal(dec, d))}
        out["rows"].append(row)
        print(kname, bs, {k: (round(v["ratio"], 3), round(v["compress_gbps"], 2), round(v["decompress_gbps"], 2), v["ok"]) for k, v in row.items() if isinstance(v, dict)}, file=sys.stderr)
print(json.dumps(out))
