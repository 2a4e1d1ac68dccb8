"""GPU parity: CUDA LZ77 (both reference variants) vs the oracle, through the C-ABI.

Token streams must be bit-exact with the reference's lz77_compress run per block
on a fresh table (SURVEY.md §8a parity contracts), and decode must give the input back.
"""
import binascii

import numpy as np
import pytest

from helpers import first_diff, fnv1a64, lcg_bytes

pytestmark = pytest.mark.gpu


def _to_dev(ctx, a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).to(ctx.device)


def _corpus(n, kind=0, seed=20261018):
    from compression_algorithms_b200 import corpus
    return corpus.generate(n, kind, seed)


def _as_np(data):
    return np.frombuffer(bytes(data), dtype=np.uint8) if isinstance(data, (bytes, bytearray)) else data


def _encode_check(ctx, ob, data, variant, block):
    """returns (stream, per-block list of token bytes from the GPU)"""
    from compression_algorithms_b200 import device as dv
    data = _as_np(data)
    n = data.size
    st = dv.lz77_encode(ctx, _to_dev(ctx, data), variant, block)
    bs = n if (block == 0 or block > n) else block
    nblocks = (n + bs - 1) // bs
    exp_out, exp_sizes = ob.port_lz77_compress_blocks(data, bs, variant)
    sizes = st.block_sizes.cpu().numpy().astype(np.uint64)
    off = st.block_off.cpu().numpy()
    out = st.out[: st.total_bytes].cpu().numpy()
    bad = np.nonzero(sizes != exp_sizes)[0]
    assert bad.size == 0, "block sizes differ first at block %d: got %d want %d" % (bad[0], sizes[bad[0]], exp_sizes[bad[0]])
    for b in range(nblocks):
        nbytes = int(exp_sizes[b]) // 8 + 1 if variant == 0 else int(exp_sizes[b])
        assert int(off[b + 1] - off[b]) == nbytes
        got = out[int(off[b]): int(off[b]) + nbytes]
        want = exp_out[b, :nbytes].copy()
        if variant == 0 and exp_sizes[b] % 8 == 0:
            got = got[:-1]; want = want[:-1]   # the extra byte of bit_index/8+1 is undefined in the reference (U3)
        d = first_diff(got, want)
        assert d == -1, "block %d differs at byte %d" % (b, d)
    assert int(off[nblocks]) == st.total_bytes
    dec = dv.lz77_decode(ctx, st).cpu().numpy()
    d = first_diff(dec, data)
    assert d == -1, "decode differs at byte %d" % d
    return st


def test_kat_standalone(ctx, ob):
    """SURVEY.md §4.3 known answer produced by the compiled reference."""
    from compression_algorithms_b200 import device as dv
    inp = b"abc" * 10 + b"_the quick brown fox the quick brown fox!"
    st = _encode_check(ctx, ob, inp, 0, 0)
    assert int(st.block_sizes[0].item()) == 301
    got = bytes(st.out[:38].cpu().numpy())
    want = binascii.unhexlify("c288193b007c09807da183a60c883875d28c5903428c9c3777dc8030f3060f480ae053002404")
    assert got[:37] == want[:37] and (got[37] & 0x1F) == (want[37] & 0x1F)


def test_kat_deflate(ctx, ob):
    inp = b"abc" * 16 + b"_the quick brown fox the quick brown fox!"
    st = _encode_check(ctx, ob, inp, 1, 0)
    want = binascii.unhexlify(
        "0061006200630103001f0121000e005f00740068006500200071007500690063006b002000620072006f0077006e00200066006f00780020011400130021")
    assert bytes(st.out[: st.total_bytes].cpu().numpy()) == want


@pytest.mark.parametrize("variant", [0, 1])
def test_golden_checksums(ctx, ob, variant):
    """ring wrap / eviction vectors of SURVEY.md §4.3 (200 000 bytes, one block)"""
    want = {0: {"A": (881309, 0x4ECEC64A28BFCA73), "B": (1697723, 0xEFC63237F46EA7C1)},
            1: {"A": (185176, 0xE62329D39182A5B3), "B": (359204, 0x80E2EA207690B525)}}[variant]
    for mode, (size, h) in want.items():
        data = np.frombuffer(lcg_bytes(mode), dtype=np.uint8)
        st = _encode_check(ctx, ob, data, variant, 0)
        assert int(st.block_sizes[0].item()) == size
        if variant == 1:
            assert fnv1a64(st.out[:size].cpu().numpy().tobytes()) == h
        else:
            raw = bytearray(st.out[: size // 8 + 1].cpu().numpy().tobytes())
            raw[-1] &= (1 << (size % 8)) - 1
            assert fnv1a64(bytes(raw)) == h


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("kind", [0, 1, 3])
@pytest.mark.parametrize("block", [65536, 262144])
def test_block_parity(ctx, ob, variant, kind, block):
    _encode_check(ctx, ob, _corpus(1_000_003, kind, 11), variant, block)


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 31, 32, 33, 1000])
def test_tiny_and_ragged(ctx, ob, variant, n):
    _encode_check(ctx, ob, _corpus(n, 0, 3), variant, 0)
    if n >= 33:
        _encode_check(ctx, ob, _corpus(n, 1, 3), variant, 16)   # many ragged 16-byte blocks


@pytest.mark.parametrize("variant", [0, 1])
def test_skewed_long_chains(ctx, ob, variant):
    """2-symbol skewed input: probe chains thousands of slots long"""
    _encode_check(ctx, ob, _corpus(40_000, 2, 5), variant, 0)


@pytest.mark.parametrize("variant", [0, 1])
def test_zero_bytes_and_overshoot(ctx, ob, variant):
    """U1: data containing 0x00 lets a match run past the end of the block"""
    rng = np.random.default_rng(9)
    data = rng.integers(0, 3, size=70000, dtype=np.uint8)
    data[-40:] = 0
    _encode_check(ctx, ob, data, variant, 65536)
    _encode_check(ctx, ob, data[:5000], variant, 0)


@pytest.mark.parametrize("variant", [0, 1])
def test_slot0_exception(ctx, ob, variant):
    """U10: pattern 0x01021578 hashes to slot 0; the reference clears slot 0 early at
    insert W-1 and keeps evicting its occupants early afterwards."""
    assert ob.port_lz77_hash(0x01021578) == 0
    rng = np.random.default_rng(4)
    n = 200_000
    data = rng.integers(97, 123, size=n, dtype=np.uint8)
    pat = np.array([0x78, 0x15, 0x02, 0x01], dtype=np.uint8)
    pos = 0
    while pos + 4 < n:
        data[pos: pos + 4] = pat
        pos += int(rng.integers(800, 3000))
    _encode_check(ctx, ob, data, variant, 0)
    _encode_check(ctx, ob, data, variant, 65536)


def test_full_size_roundtrip(ctx):
    """BASELINE.json configs[2] size (100 MB, 64 KiB blocks): round trip + size bookkeeping."""
    import torch
    from compression_algorithms_b200 import device as dv
    n = 100_000_000
    data = torch.from_numpy(_corpus(n, 0)).to(ctx.device)
    for variant in (1, 0):
        st = dv.lz77_encode(ctx, data, variant, 65536)
        off = st.block_off
        sizes = st.block_sizes
        per = sizes // 8 + 1 if variant == 0 else sizes
        assert torch.equal(off[1:] - off[:-1], per)
        assert int(off[-1].item()) == st.total_bytes
        dec = dv.lz77_decode(ctx, st)
        assert torch.equal(dec, data)


@pytest.mark.parametrize("variant", [0, 1])
def test_hbm_table_path_still_exact(ctx, ob, variant, monkeypatch):
    """blocks <= 64 KiB normally take the shared-memory match finder; force the HBM-table
    kernel (the path used for larger blocks) on the same input"""
    monkeypatch.setenv("B200_LZ_FORCE_V1", "1")
    _encode_check(ctx, ob, _corpus(300_000, 0, 21), variant, 65536)


@pytest.mark.parametrize("variant", [0, 1])
def test_match_finder_candidates(ctx, ob, variant):
    """token candidates of the shared-memory match finder at every token start == the
    reference's find() result + extension (via the oracle port's recorded find results)"""
    from compression_algorithms_b200 import device as dv
    data = _corpus(140_000, 0, 23)
    st, tok = dv.lz77_encode_debug(ctx, _to_dev(ctx, data), variant, 65536)
    tok = tok.cpu().numpy().view(np.uint32)
    W, MAXLEN = (32768, 31) if variant else (16384, 15)
    for b in range(3):
        blk = data[b * 65536: (b + 1) * 65536]
        F = (ob.port_deflate_lz77_compress(blk, want_F=True)[1] if variant else ob.port_lz77_compress(blk, want_F=True)[2])
        pad = np.concatenate([blk, np.zeros(64, dtype=np.uint8)])
        for p in np.nonzero(F != 0xFFFFFFFE)[0]:
            m = int(F[p])
            rej = m == 0xFFFFFFFF or ((p - m >= W - 1) if variant else (p - m == W))
            want = 0
            if not rej:
                l = 4
                while l < MAXLEN and pad[m + l] == pad[p + l]:
                    l += 1
                want = (p - m) | (l << 16)
            assert int(tok[b, p]) == want, "block %d position %d" % (b, p)


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("n,block", [(700 * 65536 + 321, 65536), (1000, 65536), (300_000, 0)])
def test_host_pipeline_matches_device(ctx, variant, n, block):
    """b200_lz77_compress_host / _decompress_host (chunked, three streams: H2D, kernels, D2H overlapped)
    give the same stream, sizes and offsets as the device-resident call, and round-trip."""
    import ctypes as C
    import torch
    from compression_algorithms_b200 import _lib, corpus, device as dv
    lib = _lib.core()
    data = corpus.generate(n, 0, 3)
    nb = 1 if block == 0 else (n + block - 1) // block
    cap = int(lib.b200_lz77_max_bytes(variant, n, block))
    h_in = torch.from_numpy(data.copy()).pin_memory()
    h_out = torch.empty(cap, dtype=torch.uint8).pin_memory()
    sizes = np.zeros(nb, dtype=np.uint64); off = np.zeros(nb + 1, dtype=np.uint64); tot = C.c_uint64(0)
    _lib.check(lib.b200_lz77_compress_host(ctx.handle, variant, h_in.data_ptr(), n, block, h_out.data_ptr(), cap,
                                           sizes.ctypes.data, off.ctypes.data, C.byref(tot)))
    st = dv.lz77_encode(ctx, h_in.to(ctx.device), variant, block)
    assert st.total_bytes == tot.value
    assert np.array_equal(st.out[: st.total_bytes].cpu().numpy(), h_out.numpy()[: tot.value])
    assert np.array_equal(st.block_off.cpu().numpy().astype(np.uint64), off)
    assert np.array_equal(st.block_sizes.cpu().numpy().astype(np.uint64), sizes)
    h_dec = torch.empty(n, dtype=torch.uint8).pin_memory()
    _lib.check(lib.b200_lz77_decompress_host(ctx.handle, variant, h_out.data_ptr(), tot.value, off.ctypes.data, sizes.ctypes.data,
                                             n, block, h_dec.data_ptr()))
    assert np.array_equal(h_dec.numpy(), data)


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("kind", [0, 1])
@pytest.mark.parametrize("block", [98304, 131072, 1 << 20, 0])
def test_large_blocks_in_slices(ctx, ob, variant, kind, block):
    """blocks above 65536 bytes run through the shared-memory emulation in slices (carried entries,
    parse carry, bit-offset carry of the standalone stream); block 0 = the reference's whole-buffer call"""
    _encode_check(ctx, ob, _corpus(1_500_000, kind, 13), variant, block)


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("block", [131072, 0])
def test_slot0_exception_across_slices(ctx, ob, variant, block):
    """U10 with the slot-0 cluster, its clear queue and wiped entries handed from slice to slice"""
    rng = np.random.default_rng(5)
    n = 600_000
    data = rng.integers(97, 123, size=n, dtype=np.uint8)
    pat = np.array([0x78, 0x15, 0x02, 0x01], dtype=np.uint8)
    pos = 0
    while pos + 4 < n:
        data[pos: pos + 4] = pat
        pos += int(rng.integers(700, 3000))
    _encode_check(ctx, ob, data, variant, block)


def test_whole_buffer_6mb(ctx, ob):
    """the drop-in's own call shape (one table sliding over the whole buffer) at a size with ~180 slices"""
    _encode_check(ctx, ob, _corpus(6_000_000, 0, 7), 1, 0)


def _random_token_block(rng, out_len, style):
    """A VALID deflate-variant token stream (write_literal / write_length_distance, deflate/lz77.c:176-197)
    producing exactly out_len bytes, built directly (not by the match finder) so that offsets, lengths and
    token alignments the parser rarely produces are covered: overlapping copies (offset < length, offset 1),
    offsets whose high byte is non-zero (long runs of units with a non-zero first byte), length 0 .. 255."""
    tok = bytearray()
    o = 0
    while o < out_len:
        r = rng.random()
        if o == 0 or r < (0.15 if style == "matches" else 0.93 if style == "sparse" else 0.995 if style == "nearlit" else 0.6):
            tok += bytes((0, int(rng.integers(0, 256)))); o += 1
            continue
        if style in ("sparse", "nearlit"):
            # ("nearlit": stream above 1.9 x block, the 64-units-per-step kernel)
            # literal-heavy (stream above 1.1 x block: the decoder's byte-per-lane instantiation) with short matches that
            # read what the same 32-unit step has just produced: offsets 1 .. 40, overlapping copies included
            off = int(rng.integers(1, min(o, 40) + 1))
            ln = min(int(rng.integers(0, 12)), out_len - o)
            tok += bytes((1, off & 0xFF, off >> 8, ln)); o += ln
            continue
        if style == "rle":
            off = int(rng.integers(1, min(o, 4) + 1))
        elif style == "far":
            off = int(rng.integers(max(1, min(o, 256)), min(o, 32766) + 1))
        else:
            off = int(rng.integers(1, min(o, 32766) + 1))
        ln = int(rng.integers(0, 256)) if rng.random() < 0.1 else int(rng.integers(4, 32))
        ln = min(ln, out_len - o)
        tok += bytes((1, off & 0xFF, off >> 8, ln)); o += ln
    return bytes(tok)


@pytest.mark.parametrize("style", ["mixed", "matches", "rle", "far", "sparse", "nearlit"])
def test_decoder_on_handmade_streams(ctx, ob, style, monkeypatch):
    """both deflate-variant decoders (token-parallel units, token-serial) against the oracle's byte-serial decoder"""
    import torch
    from compression_algorithms_b200 import device as dv
    rng = np.random.default_rng({"mixed": 1, "matches": 2, "rle": 3, "far": 4, "sparse": 5, "nearlit": 6}[style])
    block = 4096
    lens = [block] * 6 + [1, 2, 33, 777]
    n = sum(lens)
    # the decoder derives block lengths from (n, block): all blocks full except the last -> lay the ragged ones out as separate calls
    for blens in ([block] * 6 + [777], [1], [2], [33]):
        nn = sum(blens)
        toks = [_random_token_block(rng, L, style) for L in blens]
        stream = np.frombuffer(b"".join(toks), dtype=np.uint8)
        sizes = np.array([len(t) for t in toks], dtype=np.int64)
        off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
        want = np.concatenate([ob.port_deflate_lz77_decompress(np.frombuffer(t, dtype=np.uint8), L)[:L] for t, L in zip(toks, blens)])
        pad = np.zeros(64, dtype=np.uint8)
        st = dv.Lz77Stream(variant=dv.LZ_DEFLATE, out=_to_dev(ctx, np.concatenate([stream, pad])), block_sizes=_to_dev(ctx, sizes),
                           block_off=_to_dev(ctx, off), n=nn, block_size=block, total_bytes=int(off[-1]))
        for serial in ("0", "1"):
            monkeypatch.setenv("B200_LZ_DEC_SERIAL", serial)
            got = dv.lz77_decode(ctx, st, out=torch.zeros(nn, dtype=torch.uint8, device=ctx.device)).cpu().numpy()
            d = first_diff(got, want)
            assert d == -1, "%s decoder (%s) differs at byte %d" % ("serial" if serial == "1" else "unit", style, d)


def test_decoder_mixed_block_kinds(ctx):
    """one call whose 64 KiB blocks alternate between literal-heavy (text, random) and match-heavy (acgt, two-symbol) content:
    the two instantiations of lz77_decode_units_kernel share the blocks by stream size and together restore the input"""
    import torch
    from compression_algorithms_b200 import device as dv
    parts = [_corpus(65536, kind, 40 + i) for i, kind in enumerate([0, 1, 3, 2, 1, 0, 2, 3, 0, 0, 1])] + [_corpus(12345, 0, 77)]
    data = np.concatenate(parts)
    d = _to_dev(ctx, data)
    st = dv.lz77_encode(ctx, d, dv.LZ_DEFLATE, 65536)
    sizes = st.block_sizes.cpu().numpy()
    assert (sizes[:11] * 10 >= 65536 * 11).any() and (sizes[:11] * 10 < 65536 * 11).any()      # both kinds of block are present
    got = dv.lz77_decode(ctx, st, out=torch.full((data.size,), 0xEE, dtype=torch.uint8, device=ctx.device)).cpu().numpy()
    assert first_diff(got, data) == -1
