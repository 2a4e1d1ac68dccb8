"""GPU parity of the two-phase match finder (csrc/lz77_v4.cu; deflate variant, blocks of at most 65536 bytes; taken by
default for text-like input, forced here with B200_LZ_V4=1): same oracle comparisons as tests/test_gpu_lz77.py, plus byte
equality with lz77_v2_kernel (B200_LZ_V4=0).
Blocks it hands back (a cluster on slot 0 / the table end, a cluster above 16383 entries) run through lz77_v2_kernel,
so the skewed / slot-0 cases below exercise the hand-back list."""
import numpy as np
import pytest

from test_gpu_lz77 import _corpus, _encode_check, _to_dev

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _v4(monkeypatch):
    monkeypatch.setenv("B200_LZ_V4", "1")


@pytest.mark.parametrize("kind", [0, 1, 2, 3])
@pytest.mark.parametrize("block", [65536, 32768, 4096, 50000])
def test_block_parity(ctx, ob, kind, block):
    _encode_check(ctx, ob, _corpus(1_000_003, kind, 11), 1, block)


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 31, 32, 33, 1000, 65535, 65536, 65537])
def test_tiny_and_ragged(ctx, ob, n):
    _encode_check(ctx, ob, _corpus(n, 0, 3), 1, 65536)
    if n >= 33:
        _encode_check(ctx, ob, _corpus(min(n, 5000), 1, 3), 1, 16)


def test_single_chain_blocks(ctx, ob):
    """constant and two-symbol input: one cluster holds (nearly) every entry of a block"""
    _encode_check(ctx, ob, np.zeros(200_000, dtype=np.uint8), 1, 65536)
    _encode_check(ctx, ob, np.full(70_000, 0x41, dtype=np.uint8), 1, 65536)
    _encode_check(ctx, ob, _corpus(300_000, 2, 5), 1, 65536)


def test_zero_bytes_and_overshoot(ctx, ob):
    rng = np.random.default_rng(9)
    data = rng.integers(0, 3, size=70000, dtype=np.uint8)
    data[-40:] = 0
    _encode_check(ctx, ob, data, 1, 65536)


def test_slot0_pattern_is_handed_back(ctx, ob):
    """pattern 0x01021578 hashes to slot 0 (U10): such blocks go through the default kernel"""
    rng = np.random.default_rng(4)
    n = 200_000
    data = rng.integers(97, 123, size=n, dtype=np.uint8)
    pat = np.array([0x78, 0x15, 0x02, 0x01], dtype=np.uint8)
    for p in rng.integers(0, n - 8, size=3000):
        data[p:p + 4] = pat
    _encode_check(ctx, ob, data, 1, 65536)


def test_periodic_inputs(ctx, ob):
    """short periods: a handful of long chains whose clusters merge"""
    for period in (1, 2, 3, 5, 7, 64, 255, 1000):
        base = np.random.default_rng(period).integers(0, 256, size=period, dtype=np.uint8)
        _encode_check(ctx, ob, np.tile(base, 140_000 // period + 1)[:140_000], 1, 65536)


def test_equals_default_kernel(ctx, monkeypatch):
    from compression_algorithms_b200 import device as dv
    data = _corpus(40 * 65536 + 777, 0, 77)
    d = _to_dev(ctx, data)
    a = dv.lz77_encode(ctx, d, 1, 65536)
    out_a = a.out[: a.total_bytes].cpu().numpy().copy(); sz_a = a.block_sizes.cpu().numpy().copy()
    monkeypatch.setenv("B200_LZ_V4", "0")
    b = dv.lz77_encode(ctx, d, 1, 65536)
    assert a.total_bytes == b.total_bytes
    assert np.array_equal(sz_a, b.block_sizes.cpu().numpy())
    assert np.array_equal(out_a, b.out[: b.total_bytes].cpu().numpy())


def test_match_finder_candidates(ctx, ob):
    from compression_algorithms_b200 import device as dv
    data = _corpus(140_000, 0, 23)
    st, tok = dv.lz77_encode_debug(ctx, _to_dev(ctx, data), 1, 65536)
    tok = tok.cpu().numpy().view(np.uint32)
    for b in range(2):
        blk = data[b * 65536: (b + 1) * 65536]
        F = ob.port_deflate_lz77_compress(blk, want_F=True)[1]
        pad = np.concatenate([blk, np.zeros(64, dtype=np.uint8)])
        for p in np.nonzero(F != 0xFFFFFFFE)[0]:
            m = int(F[p])
            want = 0
            if not (m == 0xFFFFFFFF or p - m >= 32767):
                l = 4
                while l < 31 and pad[m + l] == pad[p + l]:
                    l += 1
                want = (p - m) | (l << 16)
            assert int(tok[b, p]) == want, "block %d position %d" % (b, p)


@pytest.mark.parametrize("kind,expect_v4", [(0, True), (1, True), (2, False), (3, True)])
def test_default_choice_by_sample(ctx, ob, monkeypatch, kind, expect_v4):
    """no override: a byte-entropy sample of the input picks the kernel (text-like -> v4); the stream is the oracle's either way.
    The debug statistics tell which kernel ran: lz77_v4_kernel leaves its cluster counters behind stamp 8."""
    from compression_algorithms_b200 import device as dv
    monkeypatch.delenv("B200_LZ_V4")
    monkeypatch.setenv("B200_LZ_V4_MIN_BLOCKS", "1")    # (by default only calls of 64+ blocks are sampled at all)
    data = _corpus(6 * 65536, kind, 9)
    _encode_check(ctx, ob, data, 1, 65536)
    st, tok = dv.lz77_encode_debug(ctx, _to_dev(ctx, data), 1, 65536)
    stats = st.debug_stats.cpu().numpy()
    ran_v4 = bool((stats[:, 12] == 1).any())       # v4: the lane stage's pass counter
    assert ran_v4 == expect_v4


@pytest.mark.parametrize("copies", [1, 3, 20])
def test_small_special_clusters_inside_v4(ctx, ob, copies):
    """a few occurrences of the pattern that hashes to slot 0 per block (U10: the early clear) and near-random filler whose
    probes reach the table end: clusters of a few entries on slot 0 / the last slot are simulated by the kernel itself"""
    rng = np.random.default_rng(100 + copies)
    n = 12 * 65536
    data = rng.integers(0, 256, size=n, dtype=np.uint8)
    pat = np.array([0x78, 0x15, 0x02, 0x01], dtype=np.uint8)      # hashes to slot 0
    for b in range(12):
        for p in rng.integers(b * 65536, (b + 1) * 65536 - 8, size=copies):
            data[p:p + 4] = pat
    _encode_check(ctx, ob, data, 1, 65536)


@pytest.mark.parametrize("kind", [0, 3])
@pytest.mark.parametrize("last", [58395, 60470, 65535, 65473, 64001])
def test_ragged_last_block_in_the_top_chunks(ctx, ob, kind, last):
    """a last block whose length is no multiple of 64 and lies in the parse's last super-chunks: the exit of its final 64-byte
    chunk must not come from shared memory nobody wrote (found by tools/v4_stress_random.py: it indexed the super-chunk table out
    of range, an illegal address in 1 of ~3 encodes of such a block; same code in lz77_v2_kernel). Repeated: the stale bytes varied."""
    data = _corpus(2 * 65536 + last, kind, 1000 + last)
    for _ in range(6):
        _encode_check(ctx, ob, data, 1, 65536)
