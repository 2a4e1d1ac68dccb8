import sys, numpy as np, torch
sys.path.insert(0, '.')
from compression_algorithms_b200 import corpus, device as dv
from oracle import bindings as ob
ctx = dv.Context(0)
for n in (3_000_000, 5_000_000, 20_000_000, 100_000_000):
    data = corpus.generate(n, 0)
    d = torch.from_numpy(data).to(ctx.device)
    for block in (0, 65536):
        st = dv.huffman_encode(ctx, d, block)
        dec = dv.huffman_decode(ctx, st).cpu().numpy()
        bad = np.nonzero(dec != data)[0]
        msg = "n=%d block=%d total_words=%d status=%d decode_mismatches=%d" % (n, block, st.total_words, st.worst_status, bad.size)
        if bad.size:
            msg += " first=%d (chunk %d sub %d) last=%d" % (bad[0], bad[0] // 4096, (bad[0] % 4096) // 256, bad[-1])
            ch = np.unique(bad // 4096)
            msg += " bad_chunks=%d first_chunks=%s" % (ch.size, ch[:8])
        if block == 0:
            e = ob.port_huffman_compress(data)
            nw = e["word_idx"] + (1 if e["bit_idx"] else 0)
            w = st.words[:st.total_words].cpu().numpy().view(np.uint32)
            tab = np.array_equal(st.lens()[0].cpu().numpy(), e["lens"]) and np.array_equal(st.codes()[0].cpu().numpy().view(np.uint32), e["codes"])
            msg += " | tables_ok=%s nw_ok=%s" % (tab, nw == st.total_words)
            if nw == st.total_words:
                wb = np.nonzero(w != e["words"])[0]
                msg += " word_mismatches=%d" % wb.size
                if wb.size: msg += " first_word=%d" % wb[0]
        print(msg, flush=True)
