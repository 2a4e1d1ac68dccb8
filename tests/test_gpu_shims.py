"""Drop-in boundary on the GPU: the reference-named shims (libb200_{lz77,huffman,deflate,fse}.so)
called the way the reference's drivers call them, compared with the oracle; plus C drivers
shaped like the reference mains compiled with gcc against include/*.h."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "compression_algorithms_b200")


def _corpus(n, kind=0, seed=11):
    from compression_algorithms_b200 import corpus
    return corpus.generate(n, kind, seed)


def _shim(name):
    C.CDLL(os.path.join(PKG, "libb200comp.so"), mode=C.RTLD_GLOBAL)
    return C.CDLL(os.path.join(PKG, "libb200_%s.so" % name))


class BitStream(C.Structure):
    _fields_ = [("data", C.POINTER(C.c_uint8)), ("bit_index", C.c_uint64)]


class BitWriter(C.Structure):
    _fields_ = [("buffer", C.POINTER(C.c_uint32)), ("bit_idx", C.c_uint64), ("word_idx", C.c_uint64), ("buffer_size", C.c_uint64)]


class Node(C.Structure):
    pass


Node._fields_ = [("value", C.c_uint8), ("frequency", C.c_uint32), ("left", C.POINTER(Node)), ("right", C.POINTER(Node))]


def test_lz77_shim_whole_buffer(ctx, ob):
    lib = _shim("lz77")
    lib.lz77_compress.restype = C.POINTER(BitStream)
    lib.lz77_compress.argtypes = [C.c_char_p, C.c_uint64]
    lib.lz77_decompress.restype = C.POINTER(C.c_uint8)
    lib.lz77_decompress.argtypes = [C.POINTER(BitStream), C.c_uint64, C.POINTER(C.c_uint64)]
    lib.hash.restype = C.c_uint32
    assert lib.hash(C.c_uint32(0x64636261)) == 210155
    data = _corpus(300_000)
    st = lib.lz77_compress(data.tobytes(), data.size)
    bits = st.contents.bit_index
    exp, ebits = ob.port_lz77_compress(data)
    assert bits == ebits
    got = np.ctypeslib.as_array(st.contents.data, shape=(bits // 8 + 1,))
    assert np.array_equal(got[: bits // 8], exp[: bits // 8])
    if bits % 8:
        m = (1 << (bits % 8)) - 1
        assert (got[bits // 8] & m) == (exp[bits // 8] & m)
    n_out = C.c_uint64(0)
    dec = lib.lz77_decompress(st, data.size, C.byref(n_out))
    assert n_out.value == data.size and st.contents.bit_index == 0   # lz77.c:356 resets the read position
    assert np.array_equal(np.ctypeslib.as_array(dec, shape=(data.size,)), data)


def _gather(lib, root):
    codes = (C.c_uint32 * 256)()
    lens = (C.c_uint8 * 256)()
    lib.gather_codes(C.byref(root), 0, 0, codes, lens)
    return np.frombuffer(codes, dtype=np.uint32).copy(), np.frombuffer(lens, dtype=np.uint8).copy()


@pytest.mark.parametrize("n", [10, 5000, 1_000_003])
def test_huffman_shim(ctx, ob, n):
    lib = _shim("huffman")
    lib.huffman_compress.restype = Node
    lib.huffman_compress.argtypes = [C.c_char_p, C.c_uint64, C.POINTER(BitWriter)]
    lib.huffman_decompress.argtypes = [C.POINTER(BitWriter), C.POINTER(Node), C.c_char_p, C.POINTER(C.c_uint64)]
    lib.gather_codes.argtypes = [C.POINTER(Node), C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint8)]
    data = np.frombuffer(b"nine times", dtype=np.uint8) if n == 10 else _corpus(n)
    e = ob.port_huffman_compress(data)
    w = BitWriter()
    root = lib.huffman_compress(data.tobytes(), data.size, C.byref(w))
    assert (w.word_idx, w.bit_idx, w.buffer_size) == (e["word_idx"], e["bit_idx"], e["buffer_size"])
    nw = w.word_idx + (1 if w.bit_idx else 0)
    assert np.array_equal(np.ctypeslib.as_array(w.buffer, shape=(nw,)), e["words"])
    codes, lens = _gather(lib, root)
    assert np.array_equal(codes, e["codes"]) and np.array_equal(lens, e["lens"])
    assert root.frequency == data.size
    # decode: the reference's symbol count (n + symbols out of the pad bits), first n bytes equal
    _, ecount = ob.port_huffman_decompress(e["words"], e["buffer_size"], e["codes"], e["lens"], data.size)
    out = C.create_string_buffer(data.size + 64)
    cnt = C.c_uint64(data.size + 64)
    lib.huffman_decompress(C.byref(w), C.byref(root), out, C.byref(cnt))
    assert cnt.value == ecount
    assert out.raw[: data.size] == data.tobytes()
    if n == 10:
        assert out.raw[:12] == b"nine timesnn"


def test_huffman_shim_foreign_stream_and_given_codes(ctx, ob):
    """a stream the library did not produce (built from the oracle's words) goes through the
    serial decoder; _huffman_compress packs with the caller's codes; build_huffman_tree alone."""
    lib = _shim("huffman")
    lib.huffman_decompress.argtypes = [C.POINTER(BitWriter), C.POINTER(Node), C.c_char_p, C.POINTER(C.c_uint64)]
    lib.gather_codes.argtypes = [C.POINTER(Node), C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint8)]
    lib.build_huffman_tree.argtypes = [C.c_char_p, C.c_uint64, C.POINTER(C.POINTER(Node))]
    lib._huffman_compress.argtypes = [C.c_char_p, C.c_uint64, C.POINTER(C.c_uint32), C.POINTER(C.c_uint8), C.POINTER(BitWriter)]
    lib.init_bitwriter.argtypes = [C.POINTER(BitWriter), C.c_uint64]
    data = _corpus(20_000, 0, 3)
    e = ob.port_huffman_compress(data)
    rootp = C.POINTER(Node)()
    lib.build_huffman_tree(data.tobytes(), data.size, C.byref(rootp))
    codes, lens = _gather(lib, rootp.contents)
    assert np.array_equal(codes, e["codes"]) and np.array_equal(lens, e["lens"])
    # _huffman_compress into a fresh writer, then a second call appending at a bit offset
    w = BitWriter()
    lib.init_bitwriter(C.byref(w), 2 * data.size)
    cc = (C.c_uint32 * 256)(*codes.tolist()); ll = (C.c_uint8 * 256)(*lens.tolist())
    lib._huffman_compress(data.tobytes(), data.size, cc, ll, C.byref(w))
    assert (w.word_idx, w.bit_idx) == (e["word_idx"], e["bit_idx"])
    nw = w.word_idx + (1 if w.bit_idx else 0)
    assert np.array_equal(np.ctypeslib.as_array(w.buffer, shape=(nw,)), e["words"])
    lib._huffman_compress(data.tobytes(), data.size, cc, ll, C.byref(w))
    e2 = ob.port_huffman_compress(np.concatenate([data, data]))
    if np.array_equal(e2["codes"], e["codes"]):   # doubling every count keeps the tree
        nw2 = w.word_idx + (1 if w.bit_idx else 0)
        assert (w.word_idx, w.bit_idx) == (e2["word_idx"], e2["bit_idx"])
        assert np.array_equal(np.ctypeslib.as_array(w.buffer, shape=(nw2,)), e2["words"])
    # foreign stream -> serial decoder
    words = np.ascontiguousarray(e["words"])
    fw = BitWriter(words.ctypes.data_as(C.POINTER(C.c_uint32)), e["bit_idx"], e["word_idx"], e["buffer_size"])
    _, ecount = ob.port_huffman_decompress(e["words"], e["buffer_size"], e["codes"], e["lens"], data.size)
    out = C.create_string_buffer(data.size + 64)
    cnt = C.c_uint64(data.size + 64)
    lib.huffman_decompress(C.byref(fw), rootp, out, C.byref(cnt))
    assert cnt.value == ecount and out.raw[: data.size] == data.tobytes()


def test_deflate_shim(ctx, ob, tmp_path):
    lib = _shim("deflate")
    lib.lz77_compress.argtypes = [C.c_char_p, C.c_uint64, C.c_char_p, C.POINTER(C.c_uint64), C.c_void_p]
    lib.lz77_decompress.argtypes = [C.c_char_p, C.c_uint64, C.c_char_p, C.POINTER(C.c_uint64)]
    # one block, KAT of SURVEY.md §4.3
    kat = b"abc" * 16 + b"_the quick brown fox the quick brown fox!"
    out = C.create_string_buffer(2 * len(kat))
    n_out = C.c_uint64(0)
    lib.lz77_compress(kat, len(kat), out, C.byref(n_out), None)
    assert out.raw[: n_out.value].hex() == ("0061006200630103001f0121000e005f00740068006500200071007500690063006b00200062"
                                            "0072006f0077006e00200066006f00780020011400130021")
    back = C.create_string_buffer(len(kat) + 64)
    n_back = C.c_uint64(len(kat) + 64)
    lib.lz77_decompress(out.raw[: n_out.value], n_out.value, back, C.byref(n_back))
    assert n_back.value == len(kat) and back.raw[: len(kat)] == kat
    # one 64 KiB block against the oracle
    data = _corpus(65536, 0, 5)
    out = C.create_string_buffer(2 * data.size)
    lib.lz77_compress(data.tobytes(), data.size, out, C.byref(n_out), None)
    exp = ob.port_deflate_lz77_compress(data)
    assert n_out.value == exp.size and out.raw[: exp.size] == exp.tobytes()


def _build_driver(name, tmp_path):
    exe = str(tmp_path / name)
    src = os.path.join(ROOT, "tests", "drivers", name + "_driver.c")
    cmd = ["gcc", "-O2", "-I", os.path.join(ROOT, "include"), src, "-o", exe, "-L", PKG, "-lb200_" + name, "-lb200comp",
           "-Wl,-rpath," + PKG]
    subprocess.run(cmd, check=True)
    return exe


def test_c_drivers(ctx, ob, tmp_path):
    """the reference's own driver flow, in C, against the shims"""
    data = _corpus(1_500_000, 0, 9)
    path = tmp_path / "enwik_synth"
    path.write_bytes(data.tobytes())
    # huffman
    r = subprocess.run([_build_driver("huffman", tmp_path), str(path), str(tmp_path / "h.words")], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 0 and "SUCCESS" in r.stdout, r.stdout + r.stderr
    e = ob.port_huffman_compress(data)
    assert np.array_equal(np.fromfile(tmp_path / "h.words", dtype=np.uint32), e["words"])
    assert "Compressed size: %d" % e["buffer_size"] in r.stdout
    # lz77 (whole buffer, one sequential stream: keep it small)
    small = tmp_path / "small"
    small.write_bytes(data[:200_000].tobytes())
    r = subprocess.run([_build_driver("lz77", tmp_path), str(small), str(tmp_path / "l.bits")], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 0 and "SUCCESS" in r.stdout, r.stdout + r.stderr
    exp, bits = ob.port_lz77_compress(data[:200_000])
    got = np.fromfile(tmp_path / "l.bits", dtype=np.uint8)
    assert np.array_equal(got[: bits // 8], exp[: bits // 8])
    # deflate: file in, <name>.deflate out = concatenation of the per-block oracle streams
    r = subprocess.run([_build_driver("deflate", tmp_path), str(path)], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 0, r.stdout + r.stderr
    blocks, sizes = ob.port_lz77_compress_blocks(data, 65536, 1)
    expect = np.concatenate([blocks[b, : int(sizes[b])] for b in range(len(sizes))])
    got = np.fromfile(tmp_path / "enwik_synth.deflate", dtype=np.uint8)
    assert np.array_equal(got, expect)
    assert np.array_equal(np.fromfile(tmp_path / "enwik_synth.deflate.out", dtype=np.uint8), data)
    # without the index the block boundaries are recovered from the token flags
    os.remove(tmp_path / "enwik_synth.deflate.idx")
    os.remove(tmp_path / "enwik_synth.deflate.out")
    lib = _shim("deflate")
    lib.decompress.argtypes = [C.c_void_p, C.c_char_p]
    lib.decompress(None, str(tmp_path / "enwik_synth.deflate").encode())
    assert np.array_equal(np.fromfile(tmp_path / "enwik_synth.deflate.out", dtype=np.uint8), data)
    # fse
    r = subprocess.run([_build_driver("fse", tmp_path), str(path)], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 0 and "SUCCESS" in r.stdout and "Normalised sum: 256" in r.stdout, r.stdout + r.stderr


def test_fse_shim_tables(ctx, ob):
    lib = _shim("fse")
    sz = C.c_size_t * 256
    lib.fse_build_frequency_table.argtypes = [C.c_char_p, C.c_size_t, sz]
    lib.fse_normalize_frequency_table.argtypes = [sz]
    lib.fse_build_transition_table.argtypes = [sz, C.c_uint32 * 256]
    data = np.frombuffer(b"nine times", dtype=np.uint8)
    freq = sz()
    lib.fse_build_frequency_table(data.tobytes(), data.size, freq)
    assert list(freq) == np.bincount(data, minlength=256).tolist()
    lib.fse_normalize_frequency_table(freq)
    got = {chr(s): freq[s] for s in range(256) if freq[s]}
    assert got == {" ": 24, "e": 62, "i": 49, "m": 24, "n": 49, "s": 24, "t": 24}
    big = _corpus(200_000, 0, 2)
    lib.fse_build_frequency_table(big.tobytes(), big.size, freq)
    lib.fse_normalize_frequency_table(freq)
    exp = ob.port_fse_normalize(np.bincount(big, minlength=256))
    assert list(freq) == exp.tolist()
    tt = (C.c_uint32 * 256)()
    lib.fse_build_transition_table(freq, tt)
    ett, _, _ = ob.port_fse_tables(exp)
    assert list(tt) == ett.tolist()
