"""GPU parity: CUDA Huffman path vs the oracle (bit-exact), through the C-ABI."""
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


def _check_whole(ctx, ob, data):
    from compression_algorithms_b200 import device as dv
    data = np.frombuffer(bytes(data), dtype=np.uint8) if isinstance(data, (bytes, bytearray)) else data
    exp = ob.port_huffman_compress(data)
    st = dv.huffman_encode(ctx, _to_dev(ctx, data), 0)
    assert st.worst_status == 0
    codes = u32(st.codes()[0]); lens = st.lens()[0].cpu().numpy()
    assert np.array_equal(lens, exp["lens"]), "code lengths differ"
    assert np.array_equal(codes, exp["codes"]), "codes differ"
    nw = exp["word_idx"] + (1 if exp["bit_idx"] else 0)
    assert st.total_words == nw
    assert int(st.block_bits()[0].item()) == 32 * exp["word_idx"] + exp["bit_idx"]
    got = u32(st.words[:nw])
    assert first_diff(got, exp["words"]) == -1, "stream differs at word %d" % first_diff(got, exp["words"])
    dec = dv.huffman_decode(ctx, st).cpu().numpy()
    assert first_diff(dec, data) == -1, "decode differs at byte %d" % first_diff(dec, data)
    return st, exp


def test_nine_times(ctx, ob):
    """the reference's only fixture, algorithms/huffman/main.c:25-31"""
    st, exp = _check_whole(ctx, ob, b"nine times")
    assert u32(st.words[:1])[0] == 0x39A5EB30


@pytest.mark.parametrize("kind", [0, 1, 2, 3])
@pytest.mark.parametrize("n", [1, 2, 15, 16, 17, 4095, 4096, 4097, 65536, 100003, 1 << 20])
def test_whole_buffer_parity(ctx, ob, kind, n):
    data = _corpus(n, kind)
    if len(np.unique(data)) < 2:
        pytest.skip("single symbol: reference exits (U6)")
    _check_whole(ctx, ob, data)


@pytest.mark.parametrize("kind", [0, 1, 3])
@pytest.mark.parametrize("block", [4096, 65536, 262144])
def test_per_block_parity(ctx, ob, kind, block):
    from compression_algorithms_b200 import device as dv
    n = 1_000_003
    data = _corpus(n, kind, 7)
    st = dv.huffman_encode(ctx, _to_dev(ctx, data), block)
    assert st.worst_status == 0
    nblocks = (n + block - 1) // block
    words = u32(st.words[: st.total_words])
    bw = st.block_word().cpu().numpy()
    bb = st.block_bits().cpu().numpy()
    codes = u32(st.codes()); lens = st.lens().cpu().numpy()
    assert bw[nblocks] == st.total_words
    for b in range(nblocks):
        blk = data[b * block: (b + 1) * block]
        exp = ob.port_huffman_compress(blk)
        assert np.array_equal(lens[b], exp["lens"]) and np.array_equal(codes[b], exp["codes"]), "table of block %d" % b
        assert int(bb[b]) == 32 * exp["word_idx"] + exp["bit_idx"]
        nw = exp["word_idx"] + (1 if exp["bit_idx"] else 0)
        assert int(bw[b + 1] - bw[b]) == nw
        got = words[int(bw[b]): int(bw[b]) + nw]
        assert first_diff(got, exp["words"]) == -1, "block %d word %d" % (b, first_diff(got, exp["words"]))
    dec = dv.huffman_decode(ctx, st).cpu().numpy()
    assert first_diff(dec, data) == -1


def test_long_codes(ctx, ob):
    """Fibonacci-like counts force code lengths well above the 12-bit primary table."""
    fib = [1, 1]
    while len(fib) < 28:
        fib.append(fib[-1] + fib[-2])
    data = np.repeat(np.arange(len(fib), dtype=np.uint8) + 33, fib)
    rng = np.random.default_rng(3)
    rng.shuffle(data)
    st, exp = _check_whole(ctx, ob, data)
    assert exp["lens"].max() > 20


def test_single_symbol_is_flagged(ctx):
    """U6: the reference exit(1)s; the CUDA path reports status 1 instead of encoding."""
    from compression_algorithms_b200 import device as dv
    data = np.full(5000, 65, dtype=np.uint8)
    st = dv.huffman_encode(ctx, _to_dev(ctx, data), 0)
    assert st.worst_status == 1


def test_serial_decoder_matches_reference_count(ctx, ob):
    """U5: the reference decoder returns n + k symbols; the index-free GPU decoder
    reproduces that count and the bytes."""
    from compression_algorithms_b200 import device as dv
    for data in (b"nine times", bytes(_corpus(20000, 0, 5))):
        exp = ob.port_huffman_compress(data)
        ref_out, ref_n = ob.port_huffman_decompress(exp["words"], exp["buffer_size"], exp["codes"], exp["lens"], len(data))
        words = _to_dev(ctx, exp["words"].view(np.int32))
        codes = _to_dev(ctx, exp["codes"].view(np.int32))
        lens = _to_dev(ctx, exp["lens"])
        out, cnt = dv.huffman_decode_serial(ctx, words, exp["buffer_size"], codes, lens, len(data) + 64)
        assert cnt == ref_n
        assert bytes(out[: min(cnt, len(data) + 64)].cpu().numpy()) == bytes(ref_out)
    assert ref_n >= len(data)


def test_golden_checksums(ctx, ob):
    """SURVEY.md §4.3 checksum vectors produced by the compiled reference."""
    from helpers import fnv1a64, lcg_bytes
    from compression_algorithms_b200 import device as dv
    want = {"A": (12500, 0, 0xF0780B1AF82ACB11), "B": (29761, 11, 0x8A5F6D48D34665D4)}
    for mode, (wi, bi, h) in want.items():
        data = np.frombuffer(lcg_bytes(mode), dtype=np.uint8)
        st = dv.huffman_encode(ctx, _to_dev(ctx, data), 0)
        bits = int(st.block_bits()[0].item())
        assert (bits // 32, bits % 32) == (wi, bi)
        assert fnv1a64(u32(st.words[: st.total_words]).tobytes()) == h


def test_full_size_roundtrip(ctx):
    """100 MB (BASELINE.json configs[0] size): size-independent property -- decode(encode(x)) == x,
    and the per-block streams tile the output exactly."""
    import torch
    from compression_algorithms_b200 import device as dv
    n = 100_000_000
    data = torch.from_numpy(_corpus(n, 0)).to(ctx.device)
    for block in (0, 65536):
        st = dv.huffman_encode(ctx, data, block)
        assert st.worst_status == 0
        bw = st.block_word(); bb = st.block_bits()
        assert torch.all((bb + 31) // 32 == bw[1:] - bw[:-1])
        dec = dv.huffman_decode(ctx, st)
        assert torch.equal(dec, data)


@pytest.mark.parametrize("kind", [0, 1, 3])
@pytest.mark.parametrize("cuts", [(500_001,), (1, 4097, 4097, 700_000), (0, 333_333, 1_000_003)])
def test_one_table_over_shards(ctx, ob, kind, cuts):
    """SURVEY.md §8e whole-buffer mode over ranks, with the ranks played by one GPU: shard histograms summed
    (the all-reduce), every shard packed with the table of the sum, streams spliced at the prefix of the bit
    counts == huffman_compress on the whole buffer (huffman.c:288-328); every shard decodes on its own.
    Cuts at odd byte positions, repeated cuts (empty shards) and a 1-byte shard included."""
    import torch
    from compression_algorithms_b200 import device as dv, sharding
    n = 1_000_003
    data = _corpus(n, kind, 11)
    exp = ob.port_huffman_compress(data)
    edges = [0] + list(cuts) + [n]
    shards = [_to_dev(ctx, data[a:b]) for a, b in zip(edges[:-1], edges[1:])]
    freq = torch.stack([dv.huffman_histogram(ctx, s) for s in shards]).sum(0)
    assert np.array_equal(freq.cpu().numpy(), np.bincount(data, minlength=256))
    enc = [dv.huffman_encode_with_freq(ctx, s, freq) for s in shards]
    for st, _ in enc:
        assert st.worst_status == 0
        assert np.array_equal(u32(st.codes()[0]), exp["codes"]) and np.array_equal(st.lens()[0].cpu().numpy(), exp["lens"])
    bits = [b for _, b in enc]
    off = np.concatenate([[0], np.cumsum(bits)])
    assert off[-1] == 32 * exp["word_idx"] + exp["bit_idx"]
    nwords, nbytes = sharding.whole_stream_size(int(off[-1]))
    assert nbytes == exp["buffer_size"] and nwords == len(exp["words"])
    out = torch.zeros(nwords, dtype=torch.int32, device=ctx.device)
    for (st, b), o in reversed(list(zip(enc, off[:-1]))):      # any order: boundary words are OR-ed atomically
        dv.huffman_splice(ctx, out, int(o), st.words, b)
    ctx.sync()
    got = u32(out)
    assert first_diff(got, exp["words"]) == -1, "spliced stream differs at word %d" % first_diff(got, exp["words"])
    for (st, _), s in zip(enc, shards):
        if s.numel():
            assert torch.equal(dv.huffman_decode(ctx, st), s)


def test_splice_rejects_small_destination(ctx):
    import torch
    from compression_algorithms_b200 import device as dv
    src = torch.full((4,), -1, dtype=torch.int32, device=ctx.device)
    dst = torch.zeros(4, dtype=torch.int32, device=ctx.device)
    with pytest.raises(RuntimeError):
        dv.huffman_splice(ctx, dst, 1, src, 128)
    dv.huffman_splice(ctx, dst, 3, src, 100)
    ctx.sync()
    assert u32(dst).tolist() == [0x1FFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFE000000]


def test_one_table_argument_errors(ctx):
    """misaligned input pointer and an empty shard through the sharded whole-buffer entry points"""
    import torch
    from compression_algorithms_b200 import device as dv
    buf = torch.zeros(4096 + 16, dtype=torch.uint8, device=ctx.device)
    with pytest.raises(RuntimeError):
        dv.huffman_histogram(ctx, buf[1:4097])          # d_in must be 16-byte aligned
    freq = dv.huffman_histogram(ctx, torch.tensor(list(b"abracadabra" * 100), dtype=torch.uint8, device=ctx.device))
    assert int(freq.sum().item()) == 1100 and int(freq[ord("a")].item()) == 500
    st, bits = dv.huffman_encode_with_freq(ctx, torch.empty(0, dtype=torch.uint8, device=ctx.device), freq)
    assert bits == 0 and st.total_words == 0 and st.worst_status == 0
    assert int(st.lens()[0][ord("a")].item()) == 1    # the table of the global histogram is there all the same
