import sys, torch
sys.path.insert(0, '.')
from compression_algorithms_b200 import corpus, device as dv
ctx = dv.Context(0)
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 2
data = corpus.generate(nb * 65536, 0, 5)
d = torch.from_numpy(data).to(ctx.device)
for _ in range(2):
    st = dv.lz77_encode(ctx, d, 1, 65536)
torch.cuda.synchronize()
print("ok", st.total_bytes)
