"""Generates tests/golden/reference_vectors.json from the COMPILED, UNMODIFIED reference
(oracle/_ref, built by `make -C oracle ref` from /root/reference). Run in the build
container only; the JSON is committed so the checks travel to the GPU box.

    python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import fnv1a64, lcg_bytes  # noqa: E402
from oracle import bindings as ob  # noqa: E402
from compression_algorithms_b200 import corpus  # noqa: E402


def inputs():
    yield "nine_times", b"nine times"
    yield "abc10_fox", b"abc" * 10 + b"_the quick brown fox the quick brown fox!"
    yield "abc16_fox", b"abc" * 16 + b"_the quick brown fox the quick brown fox!"
    yield "lcg_A", lcg_bytes("A")
    yield "lcg_B", lcg_bytes("B")
    for kind, name in ((0, "enwik"), (1, "acgt"), (2, "skewed"), (3, "random")):
        n = 30000 if kind == 2 else 300000
        yield "%s_%d_seed5" % (name, n), bytes(corpus.generate(n, kind, 5))
    # slot-0 exception (U10): a pattern with hash == 0 every ~1-3 KB
    rng = np.random.default_rng(4)
    d = rng.integers(97, 123, size=200000, dtype=np.uint8)
    pos = 0
    while pos + 4 < d.size:
        d[pos: pos + 4] = (0x78, 0x15, 0x02, 0x01)
        pos += int(rng.integers(800, 3000))
    yield "slot0_pattern", bytes(d)


def main():
    assert ob.have_ref(), "build oracle/_ref first: make -C oracle ref"
    vec = {"hash": {hex(x): ob.ref_lz77_hash(x) for x in (0, 0x64636261, 0xFFFFFFFF, 0x20656874, 0x01021578)}}
    cases = {}
    for name, data in inputs():
        c = {"n": len(data), "input_fnv": "%016x" % fnv1a64(data)}
        s, bits = ob.ref_lz77_compress(data)
        c["lz77"] = {"bit_index": bits, "fnv": "%016x" % fnv1a64(bytes(s))}
        t = ob.ref_deflate_lz77_compress(data)
        c["deflate"] = {"bytes": int(t.size), "fnv": "%016x" % fnv1a64(bytes(t))}
        # frequencies[286] of the token stream through the reference's own append_huffman_tree_literal/_pair
        # (deflate/lz77.c:231,273; deflate/huffman.c:49-62): pinned. The entropy-coded stream is the
        # oracle port's (the reference stops at its TODO, lz77.c:279): "port" marks it as unpinned.
        fr = ob.ref_deflate_token_frequencies(t)
        e = ob.port_dfl_encode(t)
        c["deflate_freq"] = {"nonzero": {str(i): int(fr[i]) for i in np.nonzero(fr)[0]}}
        c["deflate_huff_port"] = {"bits": int(e["bits"]), "words_fnv": "%016x" % fnv1a64(e["words"].tobytes()),
                                  "lens": bytes(e["lens"]).hex()}
        if len(data) <= 100:
            c["lz77"]["hex"] = bytes(s).hex()
            c["deflate"]["hex"] = bytes(t).hex()
        for blk in (65536,):
            if len(data) > blk:
                ss, bb = ob.ref_lz77_compress_blocks(data, blk)
                c["lz77_blocks_%d" % blk] = {"bits": [int(x) for x in bb], "fnv": "%016x" % fnv1a64(b"".join(bytes(x) for x in ss))}
                tt, nn = ob.ref_deflate_lz77_compress_blocks(data, blk)
                c["deflate_blocks_%d" % blk] = {"bytes": [int(x) for x in nn], "fnv": "%016x" % fnv1a64(b"".join(bytes(x) for x in tt))}
        if len(set(data)) >= 2:
            h = ob.ref_huffman_compress(data)
            dec, cnt = ob.ref_huffman_decompress(h["words"], h["buffer_size"], h["codes"], h["lens"], len(data))
            c["huffman"] = {"word_idx": h["word_idx"], "bit_idx": h["bit_idx"], "buffer_size": h["buffer_size"],
                            "words_fnv": "%016x" % fnv1a64(h["words"].tobytes()),
                            "lens": bytes(h["lens"]).hex(), "codes_fnv": "%016x" % fnv1a64(h["codes"].tobytes()),
                            "decoder_count": int(cnt)}
        cases[name] = c
    vec["cases"] = cases
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_vectors.json")
    with open(out, "w") as f:
        json.dump(vec, f, indent=1, sort_keys=True)
    print("wrote", out, len(cases), "cases")


if __name__ == "__main__":
    main()
