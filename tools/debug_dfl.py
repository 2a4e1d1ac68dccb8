"""Debug helper: dump the entropy-stage side tables of a small input."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from compression_algorithms_b200 import device as dv

ctx = dv.Context(0)
data = np.arange(65536 * 2 + 1, dtype=np.uint32).view(np.uint8)[: 65536 * 2 + 1].copy()
st = dv.deflate_compress(ctx, torch.from_numpy(data).cuda(), 65536)
L = st.layout
off = st.lz.block_off.cpu().numpy()
print("tok_off", off, "sizes", st.lz.block_sizes.cpu().numpy())
print("block_bits", st.block_bits().cpu().numpy(), "block_word", st.block_word().cpu().numpy())
cb = st.side[L.off_chunk_bits: L.off_chunk_bits + 4 * L.nchunks].view(torch.int32).cpu().numpy()
cs = st.side[L.off_chunk_state: L.off_chunk_state + L.nchunks].cpu().numpy()
print("chunk_bits", cb.reshape(L.nblocks, -1))
print("chunk_state", cs.reshape(L.nblocks, -1))
print("meta", st.meta().cpu().numpy())
print("lens[2] nonzero", np.nonzero(st.lens().cpu().numpy()[2]), "freq[2]", np.nonzero(st.freq().cpu().numpy()[2]))
print("tok tail", st.lz.out[int(off[2]) - 4: int(off[3]) + 4].cpu().numpy())
