"""GPU parity for FSE. Histogram + normalisation are PINNED (restatement of
main.zig:88-149, checked against the hand-derived known answer); the tANS stream is
checked against oracle/port/fse_port.c, which is this project's own completion of the
reference's unfinished encoder (PARITY UNPINNED)."""
import numpy as np
import pytest

from helpers import first_diff

pytestmark = pytest.mark.gpu


def _to_dev(ctx, a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a).copy()).to(ctx.device)


def _corpus(n, kind=0, seed=20261018):
    from compression_algorithms_b200 import corpus
    return corpus.generate(n, kind, seed)


def test_normalisation_known_answer(ctx):
    from compression_algorithms_b200 import device as dv
    data = np.frombuffer(b"nine times", dtype=np.uint8)
    st = dv.fse_normalize(ctx, _to_dev(ctx, data), 0)
    nm = st.norm()[0].cpu().numpy()
    assert {chr(i): int(nm[i]) for i in range(256) if nm[i]} == {" ": 24, "e": 62, "i": 49, "m": 24, "n": 49, "s": 24, "t": 24}


@pytest.mark.parametrize("kind", [0, 1, 2, 3])
@pytest.mark.parametrize("block", [0, 4096, 65536])
def test_histogram_and_normalisation_parity(ctx, ob, kind, block):
    from compression_algorithms_b200 import device as dv
    n = 700_001
    data = _corpus(n, kind, 13)
    st = dv.fse_normalize(ctx, _to_dev(ctx, data), block)
    freq = st.freq().cpu().numpy().view(np.uint32)
    norm = st.norm().cpu().numpy().view(np.uint16)
    bs = n if block == 0 else block
    for b in range((n + bs - 1) // bs):
        blk = data[b * bs: (b + 1) * bs]
        f = np.bincount(blk, minlength=256).astype(np.uint64)
        assert np.array_equal(freq[b], f), "histogram of block %d" % b
        assert np.array_equal(norm[b], ob.port_fse_normalize(f)), "normalisation of block %d" % b
        assert norm[b].sum() == 256


@pytest.mark.parametrize("kind", [0, 1, 3])
@pytest.mark.parametrize("block,seg", [(0, 1024), (65536, 1024), (65536, 256), (4096, 64), (262144, 4096)])
def test_stream_parity_and_roundtrip(ctx, ob, kind, block, seg):
    from compression_algorithms_b200 import device as dv
    n = 300_007
    data = _corpus(n, kind, 17)
    st = dv.fse_encode(ctx, _to_dev(ctx, data), block, seg)
    norm = st.norm().cpu().numpy().view(np.uint16)
    ttg = st.tt().cpu().numpy().view(np.uint32)
    bits = st.seg_bits().cpu().numpy().view(np.uint32)
    woff = st.seg_word().cpu().numpy()
    words = st.words[: st.total_words].cpu().numpy().view(np.uint64)
    bs = ((n + seg - 1) // seg * seg) if block == 0 else block
    spb = bs // seg
    checked = 0
    for b in range((n + bs - 1) // bs):
        tt, enc, cum = ob.port_fse_tables(norm[b].astype(np.uint64))
        assert np.array_equal(ttg[b], tt), "TT of block %d" % b
        for g in (0, 1, spb // 2, spb - 1):
            lo = b * bs + g * seg
            hi = min(lo + seg, (b + 1) * bs, n)
            if lo >= hi:
                continue
            import ctypes as C
            wbuf = np.zeros(seg // 8 + 4, dtype=np.uint64)
            f = ob._lib("oracle_port").port_fse_encode_stream
            f.restype = C.c_uint64
            segd = np.ascontiguousarray(data[lo:hi])
            tb = f(ob._p(segd, ob._u8p), C.c_uint64(hi - lo), ob._p(norm[b].astype(np.uint64), ob._u64p),
                   ob._p(enc, ob._u8p), ob._p(cum, ob._u16p), ob._p(wbuf, ob._u64p))
            gi = b * spb + g
            assert bits[gi] == tb, "bits of segment %d" % gi
            nw = (tb + 63) // 64
            assert int(woff[gi + 1] - woff[gi]) == nw
            assert np.array_equal(words[int(woff[gi]): int(woff[gi]) + nw], wbuf[:nw]), "words of segment %d" % gi
            checked += 1
    assert checked > 0
    dec = dv.fse_decode(ctx, st).cpu().numpy()
    assert first_diff(dec, data) == -1


@pytest.mark.parametrize("n", [1, 2, 63, 64, 65, 1023, 1025, 5000])
def test_tiny(ctx, n):
    from compression_algorithms_b200 import device as dv
    data = _corpus(n, 0, 3)
    st = dv.fse_encode(ctx, _to_dev(ctx, data), 0, 64)
    dec = dv.fse_decode(ctx, st).cpu().numpy()
    assert first_diff(dec, data) == -1


def test_full_size_roundtrip(ctx):
    """BASELINE.json configs[1]: FSE encode+decode of 100 MB on one B200."""
    import torch
    from compression_algorithms_b200 import device as dv
    n = 100_000_000
    data = torch.from_numpy(_corpus(n, 0)).to(ctx.device)
    st = dv.fse_encode(ctx, data, 65536, 1024)
    assert st.total_words * 8 < n
    sw = st.seg_word(); sb = st.seg_bits().to(torch.int64)
    assert torch.equal(sw[1:] - sw[:-1], (sb + 63) // 64)
    assert torch.equal(dv.fse_decode(ctx, st), data)
