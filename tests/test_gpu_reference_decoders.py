"""GPU streams through the REFERENCE's own decoders (oracle/_ref: the unmodified reference C, compiled in the
build container and carried to the GPU box as a built .so): every stream the CUDA encoders produce must
decode to the input with `lz77_decompress` (algorithms/lz77/lz77.c:347-377) and `huffman_decompress`
(algorithms/huffman/huffman.c:330-364). The deflate variant has no reference decoder (deflate.c:78-79 is
empty, deflate/lz77.c:282-311 is broken), so its streams are compared byte for byte with the reference
ENCODER instead."""
import numpy as np
import pytest

from helpers import first_diff, u32

pytestmark = pytest.mark.gpu


def _to_dev(ctx, a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).to(ctx.device)


def _corpus(n, kind=0, seed=20261018):
    from compression_algorithms_b200 import corpus
    return corpus.generate(n, kind, seed)


@pytest.fixture(scope="module")
def ref(ob):
    if not ob.have_ref():
        pytest.skip("oracle/_ref not present (built from /root/reference in the build container)")
    return ob


@pytest.mark.parametrize("kind", [0, 1, 3])
def test_lz77_stream_through_reference_decoder(ctx, ref, kind):
    from compression_algorithms_b200 import device as dv
    block = 65536
    data = _corpus(3 * block + 4321, kind, 77)
    st = dv.lz77_encode(ctx, _to_dev(ctx, data), dv.LZ_STANDALONE, block)
    off = st.block_off.cpu().numpy()
    bits = st.block_sizes.cpu().numpy()
    out = st.out[: st.total_bytes].cpu().numpy()
    for b in range(len(bits)):
        want = data[b * block: (b + 1) * block]
        got, osz = ref.ref_lz77_decompress(out[int(off[b]): int(off[b + 1])], int(bits[b]), want.size)
        assert osz == want.size and first_diff(got[: want.size], want) == -1, "block %d" % b


@pytest.mark.parametrize("block", [0, 65536])
def test_huffman_stream_through_reference_decoder(ctx, ref, block):
    from compression_algorithms_b200 import device as dv
    data = _corpus(4 * 65536, 0, 78)
    st = dv.huffman_encode(ctx, _to_dev(ctx, data), block)
    assert st.worst_status == 0
    words = u32(st.words[: st.total_words])
    codes = st.codes().cpu().numpy().view(np.uint32)
    lens = st.lens().cpu().numpy()
    bw = st.block_word().cpu().numpy()
    bb = st.block_bits().cpu().numpy()
    bs = data.size if block == 0 else block
    for b in range(len(bb)):
        want = data[b * bs: (b + 1) * bs]
        nbits = int(bb[b])
        buffer_size = (nbits >> 5) * 4 + ((nbits & 31) // 8) + (1 if (nbits & 31) % 8 else 0)   # huffman.c:318-320
        got, cnt = ref.ref_huffman_decompress(words[int(bw[b]): int(bw[b + 1])], buffer_size, codes[b], lens[b], want.size)
        assert cnt >= want.size and first_diff(got[: want.size], want) == -1, "block %d" % b


def test_deflate_stream_equals_reference_encoder(ctx, ref):
    from compression_algorithms_b200 import device as dv
    block = 65536
    data = _corpus(3 * block + 99, 0, 79)
    st = dv.lz77_encode(ctx, _to_dev(ctx, data), dv.LZ_DEFLATE, block)
    off = st.block_off.cpu().numpy()
    out = st.out[: st.total_bytes].cpu().numpy()
    blocks, sizes = ref.ref_deflate_lz77_compress_blocks(data, block)
    for b in range(len(sizes)):
        assert int(off[b + 1] - off[b]) == int(sizes[b])
        assert first_diff(out[int(off[b]): int(off[b + 1])], blocks[b]) == -1, "block %d" % b
    # frequencies[286] of the entropy stage through the reference's own append_huffman_tree_* functions
    ds = dv.dfl_encode(ctx, st)
    freq = ds.freq().cpu().numpy()
    for b in range(len(sizes)):
        assert np.array_equal(freq[b].astype(np.uint32), ref.ref_deflate_token_frequencies(blocks[b])), "block %d" % b
