"""How many blocks of the 100 MB corpus lz77_v4_kernel hands back to lz77_v2_kernel: python tools/v4_handback_count.py"""
import os, sys, numpy as np, torch
os.environ['B200_LZ_V4'] = '1'
sys.path.insert(0, '.')
from compression_algorithms_b200 import corpus, device as dv
ctx = dv.Context(0)
n = 100_000_000
data = corpus.generate(n, 0, 20261018)
d = torch.from_numpy(data).to(ctx.device)
st, tok = dv.lz77_encode_debug(ctx, d, 1, 65536)
s = st.debug_stats.cpu().numpy().astype(np.int64)
v4 = (s[:, 12] == 2) | (s[:, 12] == 1)
print("blocks", len(s), "by v4", int(v4.sum()), "handed back", int((~v4).sum()))
bad = np.nonzero(~v4)[0][:10]
print("first handed-back blocks", bad)
