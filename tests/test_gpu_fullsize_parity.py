"""GPU parity at BASELINE.json's full sizes (100 MB): the CUDA streams against the compiled reference
(oracle/_ref, when the prebuilt libraries travelled) or the oracle port, bit for bit -- words, code tables,
token bytes, block sizes -- not only round trips. The 1 GB headline buffer is checked the same way inside
bench.py (cpu_baseline.parity_ok, first 4096 blocks)."""
import numpy as np
import pytest

from helpers import first_diff, u32

pytestmark = pytest.mark.gpu

N = 100_000_000
BLOCK = 65536


def _to_dev(ctx, a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).to(ctx.device)


@pytest.fixture(scope="module")
def data100():
    from compression_algorithms_b200 import corpus
    return corpus.generate(N, corpus.ENWIK, corpus.DEFAULT_SEED)


def test_huffman_whole_buffer_100mb(ctx, ob, data100):
    """configs[0]: huffman_compress of the 100 MB buffer (huffman.c:288-328): tables and every word."""
    from compression_algorithms_b200 import device as dv
    exp = ob.ref_huffman_compress(data100) if ob.have_ref() else ob.port_huffman_compress(data100)
    st = dv.huffman_encode(ctx, _to_dev(ctx, data100), 0)
    assert st.worst_status == 0
    assert np.array_equal(st.lens()[0].cpu().numpy(), exp["lens"]), "code lengths differ"
    assert np.array_equal(u32(st.codes()[0]), exp["codes"]), "codes differ"
    nw = exp["word_idx"] + (1 if exp["bit_idx"] else 0)
    assert st.total_words == nw
    assert int(st.block_bits()[0].item()) == 32 * exp["word_idx"] + exp["bit_idx"]
    got = u32(st.words[:nw])
    d = first_diff(got, exp["words"])
    assert d == -1, "stream differs at word %d of %d" % (d, nw)
    dec = dv.huffman_decode(ctx, st).cpu().numpy()
    assert first_diff(dec, data100) == -1


def test_huffman_blocks_100mb(ctx, ob, data100):
    """the block-parallel mode: one huffman_compress per 64 KiB block, all 1526 tables and streams."""
    from compression_algorithms_b200 import device as dv
    if not ob.have_ref():
        pytest.skip("needs oracle/_ref (per-block reference harness)")
    words, wi, bi, codes, lens = ob.ref_huffman_compress_blocks(data100, BLOCK, threads=0)
    st = dv.huffman_encode(ctx, _to_dev(ctx, data100), BLOCK)
    assert st.worst_status == 0
    assert np.array_equal(st.lens().cpu().numpy(), lens), "code lengths differ"
    assert np.array_equal(u32(st.codes()).reshape(-1, 256), codes), "codes differ"
    bw = st.block_word().cpu().numpy()
    got = u32(st.words[: st.total_words])
    nb = len(wi)
    for b in range(nb):
        nwb = int(wi[b]) + (1 if bi[b] else 0)
        assert int(bw[b + 1] - bw[b]) == nwb, "block %d word count" % b
        assert np.array_equal(got[bw[b]: bw[b] + nwb], words[b, :nwb]), "block %d stream" % b


@pytest.mark.parametrize("variant", [1, 0])
def test_lz77_100mb(ctx, ob, data100, variant):
    """configs[2]: lz77_compress per 64 KiB block on a fresh table, deflate variant (byte tokens) and the standalone
    variant (bit tokens): every block's size and bytes."""
    from compression_algorithms_b200 import device as dv
    n = N if variant == 1 else 40_000_000    # the standalone reference clears a 24 MiB table per call: bounded sample
    data = data100[:n]
    st = dv.lz77_encode(ctx, _to_dev(ctx, data), variant, BLOCK)
    off = st.block_off.cpu().numpy().astype(np.int64)
    sizes = st.block_sizes.cpu().numpy().astype(np.int64)
    out = st.out[: st.total_bytes].cpu().numpy()
    if variant == 1:
        if ob.have_ref():
            blocks, exp_sizes = ob.ref_deflate_lz77_compress_blocks(data, BLOCK, persistent=False, threads=0)
        else:
            o, exp_sizes = ob.port_lz77_compress_blocks(data, BLOCK, 1, 0)
            blocks = [o[b, : int(exp_sizes[b])] for b in range(len(exp_sizes))]
        assert np.array_equal(sizes, np.asarray(exp_sizes, dtype=np.int64)), "block sizes differ"
        exp = np.concatenate(blocks)
        assert exp.size == st.total_bytes
        d = first_diff(out, exp)
        assert d == -1, "token stream differs at byte %d (block %d)" % (d, int(np.searchsorted(off, d, side="right")) - 1)
    else:
        if ob.have_ref():
            blocks, bits = ob.ref_lz77_compress_blocks(data, BLOCK, threads=0)
        else:
            o, bits = ob.port_lz77_compress_blocks(data, BLOCK, 0, 0)
            blocks = [o[b, : (int(bits[b]) + 7) // 8] for b in range(len(bits))]
        assert np.array_equal(sizes, np.asarray(bits, dtype=np.int64)), "bit counts differ"
        for b in range(len(bits)):
            nb = int(bits[b])
            got = out[off[b]: off[b] + (nb + 7) // 8].copy()
            exp = np.array(blocks[b], copy=True)
            if nb % 8:   # pad bits of the last byte are undefined in the reference (U3)
                mask = (1 << (nb % 8)) - 1
                got[-1] &= mask; exp[-1] &= mask
            assert np.array_equal(got, exp), "block %d differs" % b
    dec = dv.lz77_decode(ctx, st).cpu().numpy()
    assert first_diff(dec, data) == -1
