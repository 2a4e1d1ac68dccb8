"""GPU parity: the entropy stage of the deflate token stream (the reference's TODO at
algorithms/deflate/lz77.c:279) vs the oracle (oracle/port/deflate_huff_port.c), through the C-ABI.

frequencies[286], code tables, per-block bit counts and the packed words must be bit-exact with
the oracle on the same block segmentation; GPU decode must give the byte tokens and the input back.
"""
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


def _check(ctx, ob, data, block, full_parity=True):
    import torch
    from compression_algorithms_b200 import device as dv
    data = np.frombuffer(bytes(data), dtype=np.uint8) if isinstance(data, (bytes, bytearray)) else data
    n = data.size
    d = _to_dev(ctx, data)
    st = dv.deflate_compress(ctx, d, block)
    assert st.worst_status == 0
    bs = n if (block == 0 or block > n) else block
    nblocks = (n + bs - 1) // bs
    off = st.lz.block_off.cpu().numpy()
    tok = st.lz.out[: int(off[nblocks])].cpu().numpy()
    words = u32(st.words[: st.total_words])
    bw = st.block_word().cpu().numpy()
    bb = st.block_bits().cpu().numpy()
    assert int(bw[nblocks]) == st.total_words
    if full_parity:
        freq = st.freq().cpu().numpy()
        codes = st.codes().cpu().numpy().view(np.uint32)
        lens = st.lens().cpu().numpy()
        for b in range(nblocks):
            t = tok[int(off[b]): int(off[b + 1])]
            e = ob.port_dfl_encode(t)
            assert np.array_equal(freq[b].astype(np.uint64), e["freq"]), "frequencies differ in block %d" % b
            assert np.array_equal(lens[b], e["lens"]), "code lengths differ in block %d" % b
            assert np.array_equal(codes[b], e["codes"]), "codes differ in block %d" % b
            assert int(bb[b]) == e["bits"], "block %d: %d bits, oracle %d" % (b, bb[b], e["bits"])
            got = words[int(bw[b]): int(bw[b]) + (e["bits"] + 31) // 32]
            dd = first_diff(got, e["words"])
            assert dd == -1, "block %d: packed words differ at word %d" % (b, dd)
            assert int(bw[b + 1] - bw[b]) == (e["bits"] + 31) // 32
    # decode: tokens, then the input
    tok2 = dv.dfl_decode(ctx, st, torch.zeros_like(st.lz.out))[: int(off[nblocks])].cpu().numpy()
    dd = first_diff(tok2, tok)
    assert dd == -1, "decoded tokens differ at byte %d" % dd
    scratch = torch.zeros_like(st.lz.out)
    out = dv.deflate_decompress(ctx, st, tokens=scratch).cpu().numpy()
    dd = first_diff(out, data)
    assert dd == -1, "decompress differs at byte %d" % dd
    return st


def test_enwik_blocks(ctx, ob):
    st = _check(ctx, ob, _corpus(5 * 65536 + 1234), 65536)
    assert st.total_words * 4 < 0.7 * st.n      # it actually compresses (raw tokens are ~1.26 n)


def test_enwik_small_blocks(ctx, ob):
    _check(ctx, ob, _corpus(100_000, seed=7), 4096)


def test_whole_buffer_and_large_blocks(ctx, ob):
    data = _corpus(300_000, seed=3)
    _check(ctx, ob, data, 0)
    _check(ctx, ob, data, 131072)


@pytest.mark.parametrize("kind", [1, 2, 3])
def test_other_inputs(ctx, ob, kind):
    _check(ctx, ob, _corpus(200_000, kind, 11), 65536)


def test_periodic_long_head_runs(ctx, ob):
    # period 300: every match has offset 300 = 0x012C, so the match tail's first byte is 1 too and the
    # units with "first byte == 1" form one run over the whole block (the tail-parity rule across
    # threads, chunks and the per-block carry)
    rng = np.random.default_rng(5)
    base = rng.integers(1, 255, 300, dtype=np.uint8)
    data = np.tile(base, 700)[:200_000]
    _check(ctx, ob, data, 65536)
    base = rng.integers(1, 255, 257, dtype=np.uint8)
    _check(ctx, ob, np.tile(base, 600)[:150_001], 0)


def test_tiny_and_degenerate(ctx, ob):
    _check(ctx, ob, b"a", 65536)                       # one literal: one distinct symbol -> 1-bit code
    _check(ctx, ob, b"ab", 65536)
    _check(ctx, ob, b"abcde", 65536)
    _check(ctx, ob, b"aaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaa", 65536)
    _check(ctx, ob, bytes([7]) * 5000, 4096)           # literal + matches of one class per block
    _check(ctx, ob, np.arange(65536 * 2 + 1, dtype=np.uint32).view(np.uint8)[: 65536 * 2 + 1], 65536)


def test_entropy_stage_on_given_tokens(ctx, ob):
    """dfl_encode on the tokens of a separate lz77_encode call gives the same stream as deflate_compress"""
    from compression_algorithms_b200 import device as dv
    data = _corpus(3 * 65536 + 17, seed=9)
    d = _to_dev(ctx, data)
    lz = dv.lz77_encode(ctx, d, dv.LZ_DEFLATE, 65536)
    a = dv.dfl_encode(ctx, lz)
    b = dv.deflate_compress(ctx, d, 65536)
    assert a.total_words == b.total_words
    assert np.array_equal(u32(a.words[: a.total_words]), u32(b.words[: b.total_words]))


def test_roundtrip_32mb(ctx, ob):
    """size-independent properties at a larger size: round trip, and the stream is the concatenation of
    per-block streams whose bit counts match the oracle on sampled blocks"""
    from compression_algorithms_b200 import device as dv
    n = 32 * 1024 * 1024
    data = _corpus(n)
    d = _to_dev(ctx, data)
    st = dv.deflate_compress(ctx, d, 65536)
    assert st.worst_status == 0
    out = dv.deflate_decompress(ctx, st)
    assert bool((out == d).all())
    off = st.lz.block_off.cpu().numpy()
    bb = st.block_bits().cpu().numpy()
    for b in (0, 17, 255, 511):
        t = st.lz.out[int(off[b]): int(off[b + 1])].cpu().numpy()
        assert ob.port_dfl_encode(t)["bits"] == int(bb[b])


def test_host_buffer_entry_points(ctx, ob):
    """b200_deflate_compress_host / _decompress_host: same words as the device-resident path, round trip"""
    import ctypes as C
    from compression_algorithms_b200 import _lib, device as dv
    data = _corpus(5 * 65536 + 999, seed=21)
    n = data.size
    lib = _lib.core()
    L = dv.dfl_layout(n, 65536)
    cap = int(lib.b200_dfl_max_words(n, 65536))
    words = np.zeros(cap, dtype=np.uint32)
    side = np.zeros(L.bytes, dtype=np.uint8)
    tw, ws = C.c_uint64(0), C.c_uint32(0)
    _lib.check(lib.b200_deflate_compress_host(ctx.handle, data.ctypes.data, n, 65536, words.ctypes.data, cap,
                                              side.ctypes.data, side.size, C.byref(tw), C.byref(ws)))
    st = dv.deflate_compress(ctx, _to_dev(ctx, data), 65536)
    assert tw.value == st.total_words and ws.value == 0
    assert np.array_equal(words[: tw.value], u32(st.words[: st.total_words]))
    out = np.zeros(n, dtype=np.uint8)
    _lib.check(lib.b200_deflate_decompress_host(ctx.handle, words.ctypes.data, tw.value, side.ctypes.data, side.size,
                                                n, 65536, out.ctypes.data))
    assert first_diff(out, data) == -1


def test_error_paths(ctx, ob):
    """capacity errors are reported, not written past: side buffer too small, word buffer too small"""
    import ctypes as C
    import torch
    from compression_algorithms_b200 import _lib, device as dv
    data = _corpus(2 * 65536 + 5, seed=4)
    d = _to_dev(ctx, data)
    lz = dv.lz77_encode(ctx, d, dv.LZ_DEFLATE, 65536)
    st = dv.deflate_alloc(ctx, lz.n, 65536, lz)
    lib = _lib.core()
    tw, ws = C.c_uint64(0), C.c_uint32(0)
    rc = lib.b200_dfl_encode_dev(ctx.handle, C.c_void_p(lz.out.data_ptr()), lz.out.numel(), C.c_void_p(lz.block_off.data_ptr()),
                                 C.c_void_p(lz.block_sizes.data_ptr()), lz.n, 65536, C.c_void_p(st.words.data_ptr()), st.words.numel(),
                                 C.c_void_p(st.side.data_ptr()), st.side.numel() - 8, C.byref(tw), C.byref(ws))
    assert rc == 3 and b"side buffer" in lib.b200_last_error()
    guard = torch.full((64,), 0x5A, dtype=torch.int32, device=ctx.device)
    small = torch.cat([torch.zeros(100, dtype=torch.int32, device=ctx.device), guard])
    rc = lib.b200_dfl_encode_dev(ctx.handle, C.c_void_p(lz.out.data_ptr()), lz.out.numel(), C.c_void_p(lz.block_off.data_ptr()),
                                 C.c_void_p(lz.block_sizes.data_ptr()), lz.n, 65536, C.c_void_p(small.data_ptr()), 100,
                                 C.c_void_p(st.side.data_ptr()), st.side.numel(), C.byref(tw), C.byref(ws))
    assert rc == 3 and tw.value > 100            # needed size is reported
    assert bool((small[100:] == 0x5A).all())     # nothing written past the capacity
    # a good call afterwards still works (no sticky state)
    st = dv.dfl_encode(ctx, lz, stream=st)
    assert torch.equal(dv.deflate_decompress(ctx, st, tokens=torch.zeros_like(lz.out)), d)
