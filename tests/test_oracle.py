"""CPU tests: the oracle port (oracle/port) against (a) the committed golden vectors
produced by the compiled reference, (b) the known answers recorded in SURVEY.md §4.3,
and (c) the compiled reference itself when oracle/_ref is present."""
import json
import os

import numpy as np
import pytest

from helpers import fnv1a64, lcg_bytes

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_vectors.json")


def _inputs():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(os.path.dirname(GOLDEN), "make_golden.py"))
    return spec


@pytest.fixture(scope="module")
def golden():
    with open(GOLDEN) as f:
        return json.load(f)


@pytest.fixture(scope="module")
def cases():
    from compression_algorithms_b200 import corpus
    out = {"nine_times": b"nine times",
           "abc10_fox": b"abc" * 10 + b"_the quick brown fox the quick brown fox!",
           "abc16_fox": b"abc" * 16 + b"_the quick brown fox the quick brown fox!",
           "lcg_A": lcg_bytes("A"), "lcg_B": lcg_bytes("B")}
    for kind, name in ((0, "enwik"), (1, "acgt"), (2, "skewed"), (3, "random")):
        n = 30000 if kind == 2 else 300000
        out["%s_%d_seed5" % (name, n)] = bytes(corpus.generate(n, kind, 5))
    rng = np.random.default_rng(4)
    d = rng.integers(97, 123, size=200000, dtype=np.uint8)
    pos = 0
    while pos + 4 < d.size:
        d[pos: pos + 4] = (0x78, 0x15, 0x02, 0x01)
        pos += int(rng.integers(800, 3000))
    out["slot0_pattern"] = bytes(d)
    return out


def test_hash_known_answers(ob, golden):
    for k, v in golden["hash"].items():
        assert ob.port_lz77_hash(int(k, 16)) == v
    assert ob.port_lz77_hash(0x64636261) == 210155


def test_inputs_reproduce(cases, golden):
    for name, data in cases.items():
        assert "%016x" % fnv1a64(data) == golden["cases"][name]["input_fnv"], name


def test_lz77_port_vs_golden(ob, cases, golden):
    for name, data in cases.items():
        g = golden["cases"][name]
        s, bits = ob.port_lz77_compress(data)
        assert bits == g["lz77"]["bit_index"], name
        assert "%016x" % fnv1a64(bytes(s)) == g["lz77"]["fnv"], name
        t = ob.port_deflate_lz77_compress(data)
        assert t.size == g["deflate"]["bytes"], name
        assert "%016x" % fnv1a64(bytes(t)) == g["deflate"]["fnv"], name
        if "hex" in g["lz77"]:
            assert bytes(s).hex() == g["lz77"]["hex"]
            assert bytes(t).hex() == g["deflate"]["hex"]
        # decoders: bit tokens per the reference decoder, byte tokens pinned by round trip
        assert bytes(ob.port_lz77_decompress(s, len(data))[: len(data)]) == data
        assert bytes(ob.port_deflate_lz77_decompress(t, len(data))[: len(data)]) == data


def test_lz77_port_blocks_vs_golden(ob, cases, golden):
    for name, data in cases.items():
        g = golden["cases"][name]
        if "lz77_blocks_65536" not in g:
            continue
        out, sizes = ob.port_lz77_compress_blocks(data, 65536, 0)
        assert [int(x) for x in sizes] == g["lz77_blocks_65536"]["bits"], name
        cat = b"".join(bytes(out[b, : (int(sizes[b]) + 7) // 8]) for b in range(len(sizes)))
        assert "%016x" % fnv1a64(cat) == g["lz77_blocks_65536"]["fnv"], name
        out, sizes = ob.port_lz77_compress_blocks(data, 65536, 1)
        assert [int(x) for x in sizes] == g["deflate_blocks_65536"]["bytes"], name
        cat = b"".join(bytes(out[b, : int(sizes[b])]) for b in range(len(sizes)))
        assert "%016x" % fnv1a64(cat) == g["deflate_blocks_65536"]["fnv"], name


def test_huffman_port_vs_golden(ob, cases, golden):
    for name, data in cases.items():
        g = golden["cases"][name].get("huffman")
        if not g:
            continue
        h = ob.port_huffman_compress(data)
        assert (h["word_idx"], h["bit_idx"], h["buffer_size"]) == (g["word_idx"], g["bit_idx"], g["buffer_size"]), name
        assert "%016x" % fnv1a64(h["words"].tobytes()) == g["words_fnv"], name
        assert bytes(h["lens"]).hex() == g["lens"], name
        assert "%016x" % fnv1a64(h["codes"].tobytes()) == g["codes_fnv"], name
        dec, cnt = ob.port_huffman_decompress(h["words"], h["buffer_size"], h["codes"], h["lens"], len(data))
        assert cnt == g["decoder_count"], name
        assert bytes(dec[: len(data)]) == data


def test_survey_known_answers(ob):
    """SURVEY.md §4.3"""
    h = ob.port_huffman_compress(b"nine times")
    assert h["words"][0] == 0x39A5EB30 and (h["word_idx"], h["bit_idx"], h["buffer_size"]) == (0, 28, 4)
    table = {chr(s): (int(h["lens"][s]), int(h["codes"][s])) for s in range(256) if h["lens"][s]}
    assert table == {" ": (3, 4), "e": (3, 6), "i": (3, 7), "m": (3, 2), "n": (2, 0), "s": (3, 3), "t": (3, 5)}
    dec, cnt = ob.port_huffman_decompress(h["words"], 4, h["codes"], h["lens"], 10)
    assert cnt == 12 and bytes(dec[:12]) == b"nine timesnn"
    s, bits = ob.port_lz77_compress(b"abc" * 10 + b"_the quick brown fox the quick brown fox!")
    assert bits == 301
    assert bytes(s).hex()[:74] == "c288193b007c09807da183a60c883875d28c5903428c9c3777dc8030f3060f480ae0530024"
    for mode, (b0, h0, b1, h1, wi, bi, h2) in {
            "A": (881309, "4ecec64a28bfca73", 185176, "e62329d39182a5b3", 12500, 0, "f0780b1af82acb11"),
            "B": (1697723, "efc63237f46ea7c1", 359204, "80e2ea207690b525", 29761, 11, "8a5f6d48d34665d4")}.items():
        d = lcg_bytes(mode)
        s, bits = ob.port_lz77_compress(d)
        assert (bits, "%016x" % fnv1a64(bytes(s))) == (b0, h0)
        t = ob.port_deflate_lz77_compress(d)
        assert (t.size, "%016x" % fnv1a64(bytes(t))) == (b1, h1)
        hh = ob.port_huffman_compress(d)
        assert (hh["word_idx"], hh["bit_idx"], "%016x" % fnv1a64(hh["words"].tobytes())) == (wi, bi, h2)


def test_fse_normalisation_known_answer(ob):
    """hand-derived from algorithms/fse/src/main.zig:106-149"""
    f = np.zeros(256, dtype=np.uint64)
    for ch in b"nine times":
        f[ch] += 1
    nm = ob.port_fse_normalize(f)
    assert {chr(i): int(nm[i]) for i in range(256) if nm[i]} == {" ": 24, "e": 62, "i": 49, "m": 24, "n": 49, "s": 24, "t": 24}
    assert nm.sum() == 256


def test_fse_port_roundtrip(ob):
    from compression_algorithms_b200 import corpus
    for data in (b"a", b"ab", b"nine times", bytes(corpus.generate(70000, 0, 2)), bytes(corpus.generate(5000, 3, 2)),
                 bytes(corpus.generate(5000, 1, 2)), b"z" * 300):
        w, norm, tb, sz = ob.port_fse_compress(data)
        assert norm.sum() == 256
        assert sz == 8 * (tb // 64) + (tb % 64) // 8   # main.zig:67
        assert int(w[0]) & 0xFF == data[-1]            # last byte raw in the first 8 bits (main.zig:55-56)
        out, rc = ob.port_fse_decompress(w, tb, len(data), norm)
        assert rc == 0 and bytes(out) == data


@pytest.mark.parametrize("kind", [0, 1, 3])
def test_port_vs_compiled_reference(ob, kind):
    """Only where oracle/_ref exists (it is built from /root/reference in the build container
    and travels to the GPU box as a built .so)."""
    if not ob.have_ref():
        pytest.skip("oracle/_ref not built")
    from compression_algorithms_b200 import corpus
    d = corpus.generate(400000, kind, 99)
    rs, rb = ob.ref_lz77_compress_blocks(d, 65536)
    po, pb = ob.port_lz77_compress_blocks(d, 65536, 0)
    assert np.array_equal(rb, pb)
    assert all(np.array_equal(rs[b], po[b, : len(rs[b])]) for b in range(len(rs)))
    ds, dn = ob.ref_deflate_lz77_compress_blocks(d, 65536)
    qo, qn = ob.port_lz77_compress_blocks(d, 65536, 1)
    assert np.array_equal(dn, qn)
    assert all(np.array_equal(ds[b], qo[b, : len(ds[b])]) for b in range(len(ds)))
    r = ob.ref_huffman_compress(d)
    p = ob.port_huffman_compress(d)
    assert all(np.array_equal(r[k], p[k]) for k in ("words", "codes", "lens"))
    # the reference decoder accepts the port's stream
    out, cnt = ob.ref_huffman_decompress(p["words"], p["buffer_size"], p["codes"], p["lens"], d.size)
    assert bytes(out[: d.size]) == bytes(d)
    # the reference LZ77 decoder accepts the port's stream
    s, bits = ob.port_lz77_compress(d[:100000])
    out, osz = ob.ref_lz77_decompress(s, bits, 100000)
    assert bytes(out[:100000]) == bytes(d[:100000])


def test_huffman_heap_ties(ob):
    """all-equal and near-equal frequencies: the tie-break rules of the reference heap"""
    if not ob.have_ref():
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(1)
    for t in range(60):
        k = int(rng.integers(2, 257))
        syms = rng.choice(256, k, replace=False)
        f = np.zeros(256, dtype=np.int64)
        f[syms] = 1 if t % 3 == 0 else (rng.integers(1, 5, k) if t % 3 == 1 else rng.integers(1, 100000, k))
        data = np.repeat(np.arange(256, dtype=np.uint8), f)
        c1, l1 = ob.ref_huffman_tables(data)
        c2, l2, _ = ob.port_huffman_build(f)
        assert np.array_equal(c1, c2) and np.array_equal(l1, l2)


# ----------------------------------------------------------------- deflate token entropy stage
def _token_pieces(tok, codes, lens):
    """(value, length) list of the entropy stage, built independently of the port in plain Python"""
    vals, ls = [], []
    i = 0
    while i + 1 < len(tok):
        if tok[i] == 1:
            off = int(tok[i + 1]) | int(tok[i + 2]) << 8
            k = 16 - off.bit_length()
            vals.append(int(codes[256 + k])); ls.append(int(lens[256 + k]))
            if k < 15:
                vals.append(off & ((1 << (15 - k)) - 1)); ls.append(15 - k)
            vals.append(int(tok[i + 3])); ls.append(5)
            i += 4
        else:
            vals.append(int(codes[tok[i + 1]])); ls.append(int(lens[tok[i + 1]]))
            i += 2
    return vals, ls


def test_deflate_frequencies_vs_golden(ob, cases, golden):
    """frequencies[286] (deflate/lz77.c:206,231,273) — golden values come from the reference's own
    append_huffman_tree_literal/_pair; the entropy-coded stream is pinned to the port's spec vector"""
    for name, data in cases.items():
        g = golden["cases"][name]
        tok = ob.port_deflate_lz77_compress(data)
        tok = tok[0] if isinstance(tok, tuple) else tok
        fr = ob.port_dfl_frequencies(tok)
        want = np.zeros(286, dtype=np.uint64)
        for k, v in g["deflate_freq"]["nonzero"].items():
            want[int(k)] = v
        assert np.array_equal(fr, want), name
        e = ob.port_dfl_encode(tok)
        gp = g["deflate_huff_port"]
        assert e["bits"] == gp["bits"] and "%016x" % fnv1a64(e["words"].tobytes()) == gp["words_fnv"], name
        assert bytes(e["lens"]).hex() == gp["lens"], name
        back, used = ob.port_dfl_decode(e["words"], e["codes"], e["lens"], len(tok))
        assert used == e["bits"] and np.array_equal(back, tok), name


def test_deflate_entropy_stage_vs_compiled_reference(ob):
    """the pieces the reference does define: frequencies through its append_* functions, the packed
    words through its write_bits (deflate/huffman.c:18-48), and for a literal-only stream the code
    table of algorithms/huffman (same heap rule over 256 symbols)"""
    if not ob.have_ref():
        pytest.skip("oracle/_ref not built")
    from compression_algorithms_b200 import corpus
    for kind in (0, 1, 3):
        d = corpus.generate(70000, kind, 31)
        tok = ob.ref_deflate_lz77_compress(d)
        e = ob.port_dfl_encode(tok)
        assert np.array_equal(ob.ref_deflate_token_frequencies(tok).astype(np.uint64), e["freq"])
        vals, ls = _token_pieces(tok, e["codes"], e["lens"])
        w, bits = ob.ref_deflate_write_bits(vals, ls)
        assert bits == e["bits"] and np.array_equal(w, e["words"])
    # literal-only token stream over bytes: the 286-symbol build must agree with the reference's
    # 256-symbol tables (symbols 256.. are absent)
    d = corpus.generate(20000, 0, 2)
    tok = np.zeros(2 * d.size, dtype=np.uint8)
    tok[1::2] = d
    tok[0::2] = 0
    e = ob.port_dfl_encode(tok)
    c1, l1 = ob.ref_huffman_tables(d)
    assert np.array_equal(e["codes"][:256], c1) and np.array_equal(e["lens"][:256], l1)
    r = ob.ref_huffman_compress(d)
    assert np.array_equal(e["words"], r["words"])


def test_deflate_entropy_stage_degenerate(ob):
    for data in (b"a", b"ab", b"aaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaa", bytes([7]) * 5000):
        tok = ob.port_deflate_lz77_compress(data)
        tok = tok[0] if isinstance(tok, tuple) else tok
        e = ob.port_dfl_encode(tok)
        back, used = ob.port_dfl_decode(e["words"], e["codes"], e["lens"], len(tok))
        assert used == e["bits"] and np.array_equal(back, tok)
        if e["distinct"] == 1:
            assert e["lens"].max() == 1 and e["bits"] >= 1


def test_zig_huffman_port_format(ob):
    """structure of the restated Zig Huffman file (parity unpinned): pre-order tree dump, CompressedSize header with the
    last-block bit, whole-byte payload; an input of exactly 4 MiB ends with an empty last chunk"""
    from compression_algorithms_b200 import corpus
    n = (1 << 22) + 1000
    data = corpus.generate(n, corpus.ENWIK, 4)
    c = ob.port_zig_huffman_compress(data)
    # first node = the root: value 0, freq = the whole 4 MiB read buffer
    assert c[0] == 0 and int(np.frombuffer(c[1:5].tobytes(), np.uint32)[0]) == 1 << 22
    dec = ob.port_zig_huffman_decompress(c, 2 << 22)
    assert abs(int(dec.size) - n) <= 4 and np.array_equal(dec[: (1 << 22) - 8], data[: (1 << 22) - 8])
    full = ob.port_zig_huffman_compress(corpus.generate(1 << 22, corpus.ENWIK, 4))
    assert int(np.frombuffer(full[-4:].tobytes(), np.uint32)[0]) == 1          # last header: last_block = 1, 0 bytes
