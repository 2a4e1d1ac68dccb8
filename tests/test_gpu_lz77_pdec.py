"""GPU parity of the token-parallel LZ77 decoder (csrc/lz77_pdec.cu): chunk tables -> composition -> pointer emission ->
pointer-chain resolution. It is the default from 128 KiB (standalone) / 1 MiB (deflate variant) blocks on and for one block =
the whole buffer; B200_LZ_PDEC=1 forces it for every block size. Checked against the input (round trip of streams that are
themselves bit-exact with the reference encoder) and against the oracle's byte-serial decoder on hand-made streams."""
import numpy as np
import pytest

from helpers import first_diff
from test_gpu_lz77 import _corpus, _random_token_block, _to_dev

pytestmark = pytest.mark.gpu


def _roundtrip(ctx, data, variant, block):
    import torch
    from compression_algorithms_b200 import device as dv
    d = _to_dev(ctx, data)
    st = dv.lz77_encode(ctx, d, variant, block)
    out = torch.full((data.size,), 0xEE, dtype=torch.uint8, device=ctx.device)
    got = dv.lz77_decode(ctx, st, out=out).cpu().numpy()
    diff = first_diff(got, data)
    assert diff == -1, "decode differs at byte %d" % diff


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("kind", [0, 1, 3])
@pytest.mark.parametrize("block", [4096, 65536, 1 << 20, 0])
def test_roundtrip_forced(ctx, monkeypatch, variant, kind, block):
    monkeypatch.setenv("B200_LZ_PDEC", "1")
    _roundtrip(ctx, _corpus(3_000_017, kind, 31), variant, block)


@pytest.mark.parametrize("variant", [0, 1])
def test_default_dispatch_large_blocks(ctx, variant):
    """no environment override: 4 MiB blocks and the whole buffer as one block take the token-parallel path"""
    data = _corpus(9_000_001, 0, 5)
    _roundtrip(ctx, data, variant, 4 << 20)
    _roundtrip(ctx, data, variant, 0)


@pytest.mark.parametrize("variant", [0, 1])
def test_long_pointer_chains(ctx, monkeypatch, variant):
    """constant / two-symbol / short-period input: every byte copies the one before it, chains as long as the block"""
    monkeypatch.setenv("B200_LZ_PDEC", "1")
    _roundtrip(ctx, np.zeros(300_000, dtype=np.uint8), variant, 0)
    _roundtrip(ctx, np.full(200_000, 0x41, dtype=np.uint8), variant, 65536)
    _roundtrip(ctx, _corpus(400_000, 2, 5), variant, 0)
    _roundtrip(ctx, np.tile(np.arange(7, dtype=np.uint8), 60_000), variant, 0)


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 31, 32, 33, 255, 256, 257, 1000])
def test_tiny(ctx, monkeypatch, variant, n):
    monkeypatch.setenv("B200_LZ_PDEC", "1")
    _roundtrip(ctx, _corpus(n, 0, 3), variant, 0)
    if n >= 33:
        _roundtrip(ctx, _corpus(n, 1, 3), variant, 16)


@pytest.mark.parametrize("style", ["mixed", "matches", "rle", "far"])
def test_handmade_deflate_streams(ctx, ob, style, monkeypatch):
    import torch
    from compression_algorithms_b200 import device as dv
    monkeypatch.setenv("B200_LZ_PDEC", "1")
    rng = np.random.default_rng({"mixed": 11, "matches": 12, "rle": 13, "far": 14}[style])
    for block, blens in ((4096, [4096] * 6 + [777]), (40000, [40000] * 3 + [5]), (33, [33])):
        nn = sum(blens)
        toks = [_random_token_block(rng, L, style) for L in blens]
        stream = np.frombuffer(b"".join(toks), dtype=np.uint8)
        sizes = np.array([len(t) for t in toks], dtype=np.int64)
        off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
        want = np.concatenate([ob.port_deflate_lz77_decompress(np.frombuffer(t, dtype=np.uint8), L)[:L] for t, L in zip(toks, blens)])
        st = dv.Lz77Stream(variant=dv.LZ_DEFLATE, out=_to_dev(ctx, np.concatenate([stream, np.zeros(64, dtype=np.uint8)])),
                           block_sizes=_to_dev(ctx, sizes), block_off=_to_dev(ctx, off), n=nn, block_size=block, total_bytes=int(off[-1]))
        got = dv.lz77_decode(ctx, st, out=torch.zeros(nn, dtype=torch.uint8, device=ctx.device)).cpu().numpy()
        d = first_diff(got, want)
        assert d == -1, "%s, block %d: differs at byte %d" % (style, block, d)


@pytest.mark.parametrize("variant,block", [(0, 65536), (1, 1 << 20), (0, 1 << 18)])
def test_host_pipeline_two_streams(ctx, monkeypatch, variant, block):
    """b200_lz77_decompress_host decodes consecutive chunks on two kernel streams: the decoder's scratch arrays (pointer
    words, chunk tables) must come from one bank per stream"""
    import ctypes as C
    import torch
    from compression_algorithms_b200 import _lib, device as dv
    monkeypatch.setenv("B200_LZ_PDEC", "1")
    monkeypatch.setenv("B200_LZ_CHUNKS", "12")
    lib = _lib.core()
    n = 40 * (1 << 20) + 4321
    data = _corpus(n, 0, 41)
    st = dv.lz77_encode(ctx, _to_dev(ctx, data), variant, block)
    stream = st.out[: st.total_bytes].cpu().numpy()
    off = st.block_off.cpu().numpy().astype(np.uint64)
    sz = st.block_sizes.cpu().numpy().astype(np.uint64)
    out = np.empty(n, dtype=np.uint8)
    _lib.check(lib.b200_lz77_decompress_host(ctx.handle, variant, stream.ctypes.data_as(C.c_void_p), C.c_uint64(stream.size),
                                             off.ctypes.data_as(C.c_void_p), sz.ctypes.data_as(C.c_void_p), C.c_uint64(n),
                                             C.c_uint64(block), out.ctypes.data_as(C.c_void_p)))
    d = first_diff(out, data)
    assert d == -1, "host pipeline decode differs at byte %d" % d
