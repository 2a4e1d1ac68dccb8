"""TEST INFRASTRUCTURE ONLY -- ctypes bindings for the oracle libraries.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module. Nothing under compression_algorithms_b200/ does.

Two families:
  * ``ref_*``  -> oracle/_ref/*.so : the UNMODIFIED reference C compiled from
    /root/reference by oracle/Makefile (kind "reference").
  * ``port_*`` -> oracle/_build/liboracle_port.so : the CPU restatement under
    oracle/port/ (kind "port"), validated against ``ref_*`` and the golden vectors.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_u8p = C.POINTER(C.c_uint8)
_u16p = C.POINTER(C.c_uint16)
_u32p = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)
_i32p = C.POINTER(C.c_int32)
_f64p = C.POINTER(C.c_double)


def _p(a, t):
    return a.ctypes.data_as(t)


def _load(path):
    if not os.path.exists(path):
        return None
    return C.CDLL(path, mode=C.RTLD_LOCAL)


_cache = {}


def _lib(name):
    if name not in _cache:
        sub = "_build" if name == "oracle_port" else "_ref"
        _cache[name] = _load(os.path.join(_HERE, sub, "lib%s.so" % name))
    return _cache[name]


def have_ref():
    return all(_lib(n) is not None for n in ("lz77_ref", "huffman_ref", "deflate_ref"))


def have_port():
    return _lib("oracle_port") is not None


def _as_u8(data):
    if isinstance(data, (bytes, bytearray)):
        return np.frombuffer(bytes(data), dtype=np.uint8)
    return np.ascontiguousarray(data, dtype=np.uint8)


# --------------------------------------------------------------------------- #
# reference (oracle/_ref)
# --------------------------------------------------------------------------- #
def ref_lz77_hash(x):
    f = _lib("lz77_ref").orc_ref_lz77_hash
    f.restype = C.c_uint32
    return f(C.c_uint32(x))


def ref_lz77_compress(data):
    """-> (stream bytes [ceil(bits/8), tail masked], bit_index)"""
    d = _as_u8(data)
    out = np.zeros(2 * d.size + 16, dtype=np.uint8)
    bits = C.c_uint64(0)
    rc = _lib("lz77_ref").orc_ref_lz77_compress(_p(d, _u8p), C.c_uint64(d.size), _p(out, _u8p), C.byref(bits))
    if rc:
        raise RuntimeError("reference lz77_compress exited with %d" % rc)
    return out[: (bits.value + 7) // 8].copy(), bits.value


def ref_lz77_decompress(stream, bit_index, size):
    s = _as_u8(stream)
    out = np.zeros(size + 64, dtype=np.uint8)
    osz = C.c_uint64(0)
    rc = _lib("lz77_ref").orc_ref_lz77_decompress(_p(s, _u8p), C.c_uint64(bit_index), C.c_uint64(size), _p(out, _u8p), C.byref(osz))
    if rc:
        raise RuntimeError("reference lz77_decompress exited with %d" % rc)
    return out[: min(osz.value, size + 64)].copy(), osz.value


def ref_lz77_compress_blocks(data, block, threads=0):
    """-> (list of per-block streams, bit_index array)"""
    d = _as_u8(data)
    nb = (d.size + block - 1) // block
    stride = 2 * block + 16
    out = np.zeros(nb * stride, dtype=np.uint8)
    bits = np.zeros(nb, dtype=np.uint64)
    rc = _lib("lz77_ref").orc_ref_lz77_compress_blocks(
        _p(d, _u8p), C.c_uint64(d.size), C.c_uint64(block), _p(out, _u8p), C.c_uint64(stride), _p(bits, _u64p), C.c_int(threads))
    if rc:
        raise RuntimeError("reference lz77_compress exited with %d" % rc)
    return [out[b * stride: b * stride + (int(bits[b]) + 7) // 8] for b in range(nb)], bits


def ref_huffman_tables(data):
    d = _as_u8(data)
    codes = np.zeros(256, dtype=np.uint32)
    lens = np.zeros(256, dtype=np.uint8)
    rc = _lib("huffman_ref").orc_ref_huffman_tables(_p(d, _u8p), C.c_uint64(d.size), _p(codes, _u32p), _p(lens, _u8p))
    if rc:
        raise RuntimeError("reference huffman exited with %d" % rc)
    return codes, lens


def ref_huffman_compress(data):
    """-> dict(words, word_idx, bit_idx, buffer_size, codes, lens)"""
    d = _as_u8(data)
    words = np.zeros(d.size // 4 + 4, dtype=np.uint32)
    wi, bi, bs = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
    codes = np.zeros(256, dtype=np.uint32)
    lens = np.zeros(256, dtype=np.uint8)
    rc = _lib("huffman_ref").orc_ref_huffman_compress(
        _p(d, _u8p), C.c_uint64(d.size), _p(words, _u32p), C.byref(wi), C.byref(bi), C.byref(bs), _p(codes, _u32p), _p(lens, _u8p))
    if rc:
        raise RuntimeError("reference huffman_compress exited with %d" % rc)
    nw = wi.value + (1 if bi.value else 0)
    return dict(words=words[:nw].copy(), word_idx=wi.value, bit_idx=bi.value, buffer_size=bs.value, codes=codes, lens=lens)


def ref_huffman_decompress(words, buffer_size, codes, lens, expect):
    w = np.ascontiguousarray(words, dtype=np.uint32)
    cap = expect + 64
    out = np.zeros(cap, dtype=np.uint8)
    osz = C.c_uint64(0)
    codes = np.ascontiguousarray(codes, dtype=np.uint32)
    lens = np.ascontiguousarray(lens, dtype=np.uint8)
    rc = _lib("huffman_ref").orc_ref_huffman_decompress(
        _p(w, _u32p), C.c_uint64(w.size), C.c_uint64(buffer_size), _p(codes, _u32p), _p(lens, _u8p),
        _p(out, _u8p), C.c_uint64(cap), C.byref(osz))
    if rc:
        raise RuntimeError("reference huffman_decompress exited with %d" % rc)
    return out[: min(osz.value, cap)].copy(), osz.value


def ref_huffman_compress_blocks(data, block, threads=0):
    d = _as_u8(data)
    nb = (d.size + block - 1) // block
    stride = block // 4 + 4
    words = np.zeros(nb * stride, dtype=np.uint32)
    wi = np.zeros(nb, dtype=np.uint64)
    bi = np.zeros(nb, dtype=np.uint64)
    codes = np.zeros(nb * 256, dtype=np.uint32)
    lens = np.zeros(nb * 256, dtype=np.uint8)
    rc = _lib("huffman_ref").orc_ref_huffman_compress_blocks(
        _p(d, _u8p), C.c_uint64(d.size), C.c_uint64(block), _p(words, _u32p), C.c_uint64(stride),
        _p(wi, _u64p), _p(bi, _u64p), _p(codes, _u32p), _p(lens, _u8p), C.c_int(threads))
    if rc:
        raise RuntimeError("reference huffman_compress exited with %d" % rc)
    return words.reshape(nb, stride), wi, bi, codes.reshape(nb, 256), lens.reshape(nb, 256)


def ref_huffman_time(data):
    d = _as_u8(data)
    tc, td = C.c_double(0), C.c_double(0)
    cb, mm = C.c_uint64(0), C.c_uint64(0)
    rc = _lib("huffman_ref").orc_ref_huffman_time(_p(d, _u8p), C.c_uint64(d.size), C.byref(tc), C.byref(td), C.byref(cb), C.byref(mm))
    if rc:
        raise RuntimeError("reference huffman exited with %d" % rc)
    return dict(t_comp=tc.value, t_decomp=td.value, comp_bytes=cb.value, mismatches=mm.value)


def ref_deflate_lz77_compress(data):
    d = _as_u8(data)
    out = np.zeros(2 * d.size + 16, dtype=np.uint8)
    n = C.c_uint64(0)
    rc = _lib("deflate_ref").orc_ref_deflate_lz77_compress(_p(d, _u8p), C.c_uint64(d.size), _p(out, _u8p), C.byref(n))
    if rc:
        raise RuntimeError("reference deflate lz77_compress exited with %d" % rc)
    return out[: n.value].copy()


def ref_deflate_lz77_compress_blocks(data, block, persistent=False, threads=0):
    d = _as_u8(data)
    nb = (d.size + block - 1) // block
    stride = 2 * block + 16
    out = np.zeros(nb * stride, dtype=np.uint8)
    sizes = np.zeros(nb, dtype=np.uint64)
    rc = _lib("deflate_ref").orc_ref_deflate_lz77_compress_blocks(
        _p(d, _u8p), C.c_uint64(d.size), C.c_uint64(block), _p(out, _u8p), C.c_uint64(stride), _p(sizes, _u64p),
        C.c_int(1 if persistent else 0), C.c_int(threads))
    if rc:
        raise RuntimeError("reference deflate lz77_compress exited with %d" % rc)
    return [out[b * stride: b * stride + int(sizes[b])] for b in range(nb)], sizes


def ref_deflate_token_frequencies(tokens):
    """frequencies[286] through the reference's own append_huffman_tree_literal/_pair."""
    t = _as_u8(tokens)
    fr = np.zeros(286, dtype=np.uint32)
    _lib("deflate_ref").orc_ref_deflate_token_frequencies(_p(t, _u8p), C.c_uint64(t.size), _p(fr, _u32p))
    return fr


def ref_deflate_write_bits(values, lengths):
    """The reference's BitWriter (deflate/huffman.c:9-48) fed with (value, length) pairs -> (words, bits)."""
    v = np.ascontiguousarray(values, dtype=np.uint32)
    ln = np.ascontiguousarray(lengths, dtype=np.uint8)
    cap = int(ln.astype(np.uint64).sum()) // 32 + 4
    words = np.zeros(cap, dtype=np.uint32)
    f = _lib("deflate_ref").orc_ref_deflate_write_bits
    f.restype = C.c_uint64
    bits = f(_p(v, _u32p), _p(ln, _u8p), C.c_uint64(v.size), _p(words, _u32p), C.c_uint64(cap))
    return words[: (bits + 31) // 32].copy(), int(bits)


def ref_threads():
    f = _lib("deflate_ref").orc_ref_deflate_threads
    return int(f())


# --------------------------------------------------------------------------- #
# port (oracle/_build/liboracle_port.so)
# --------------------------------------------------------------------------- #
def port_lz77_hash(x):
    f = _lib("oracle_port").port_lz77_hash
    f.restype = C.c_uint32
    return f(C.c_uint32(x))


def port_lz77_compress(data, want_F=False):
    d = _as_u8(data)
    out = np.zeros(2 * d.size + 16, dtype=np.uint8)
    bits = C.c_uint64(0)
    F = np.full(max(d.size, 1), 0xFFFFFFFE, dtype=np.uint32) if want_F else None
    _lib("oracle_port").port_lz77_compress(_p(d, _u8p), C.c_uint64(d.size), _p(out, _u8p), C.byref(bits),
                                           _p(F, _u32p) if want_F else None)
    res = out[: (bits.value + 7) // 8].copy(), bits.value
    return res + (F,) if want_F else res


def port_deflate_lz77_compress(data, want_F=False):
    d = _as_u8(data)
    out = np.zeros(2 * d.size + 16, dtype=np.uint8)
    n = C.c_uint64(0)
    F = np.full(max(d.size, 1), 0xFFFFFFFE, dtype=np.uint32) if want_F else None
    _lib("oracle_port").port_deflate_lz77_compress(_p(d, _u8p), C.c_uint64(d.size), _p(out, _u8p), C.byref(n),
                                                   _p(F, _u32p) if want_F else None)
    res = out[: n.value].copy()
    return (res, F) if want_F else res


def port_lz77_compress_blocks(data, block, variant, threads=0):
    """variant 0: algorithms/lz77 (sizes = bits), 1: algorithms/deflate (sizes = bytes).
    -> (out[nblocks, stride] u8, sizes u64)"""
    d = _as_u8(data)
    nb = (d.size + block - 1) // block
    stride = 2 * block + 16
    out = np.zeros(nb * stride, dtype=np.uint8)
    sizes = np.zeros(nb, dtype=np.uint64)
    _lib("oracle_port").port_lz77_compress_blocks(
        _p(d, _u8p), C.c_uint64(d.size), C.c_uint64(block), C.c_int(variant), _p(out, _u8p), C.c_uint64(stride),
        _p(sizes, _u64p), C.c_int(threads))
    return out.reshape(nb, stride), sizes


def port_lz77_decompress(stream, size):
    s = np.concatenate([_as_u8(stream), np.zeros(8, dtype=np.uint8)])
    out = np.zeros(size + 64, dtype=np.uint8)
    f = _lib("oracle_port").port_lz77_decompress
    f.restype = C.c_uint64
    n = f(_p(s, _u8p), C.c_uint64(size), _p(out, _u8p))
    return out[:n].copy()


def port_deflate_lz77_decompress(tokens, max_out):
    t = _as_u8(tokens)
    out = np.zeros(max_out + 64, dtype=np.uint8)
    f = _lib("oracle_port").port_deflate_lz77_decompress
    f.restype = C.c_uint64
    n = f(_p(t, _u8p), C.c_uint64(t.size), _p(out, _u8p))
    return out[:n].copy()


def port_huffman_build(freq):
    fr = np.ascontiguousarray(freq, dtype=np.uint64)
    codes = np.zeros(256, dtype=np.uint32)
    lens = np.zeros(256, dtype=np.uint8)
    nodes = np.zeros(511 * 3, dtype=np.int32)
    root = C.c_int(0)
    distinct = _lib("oracle_port").port_huffman_build(_p(fr, _u64p), _p(codes, _u32p), _p(lens, _u8p), _p(nodes, _i32p), C.byref(root))
    return codes, lens, distinct


def port_huffman_compress(data):
    d = _as_u8(data)
    words = np.zeros(d.size // 4 + 4, dtype=np.uint32)
    wi, bi, bs = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
    codes = np.zeros(256, dtype=np.uint32)
    lens = np.zeros(256, dtype=np.uint8)
    rc = _lib("oracle_port").port_huffman_compress(
        _p(d, _u8p), C.c_uint64(d.size), _p(words, _u32p), C.byref(wi), C.byref(bi), C.byref(bs), _p(codes, _u32p), _p(lens, _u8p))
    if rc:
        raise RuntimeError("huffman: fewer than 2 distinct symbols (reference exit(1), U6)")
    nw = wi.value + (1 if bi.value else 0)
    return dict(words=words[:nw].copy(), word_idx=wi.value, bit_idx=bi.value, buffer_size=bs.value, codes=codes, lens=lens)


def port_huffman_encode(data, codes, lens):
    """_huffman_compress with a given table (huffman.c:267-285) -> (u32 words, bits)."""
    d = _as_u8(data)
    words = np.zeros(d.size + 2, dtype=np.uint32)
    f = _lib("oracle_port").port_huffman_encode
    f.restype = C.c_uint64
    bits = f(_p(d, _u8p), C.c_uint64(d.size), _p(np.ascontiguousarray(codes, dtype=np.uint32), _u32p),
             _p(np.ascontiguousarray(lens, dtype=np.uint8), _u8p), _p(words, _u32p))
    return words[: (bits + 31) // 32].copy(), int(bits)


def port_huffman_decompress(words, buffer_size, codes, lens, expect):
    w = np.ascontiguousarray(words, dtype=np.uint32)
    cap = expect + 64
    out = np.zeros(cap, dtype=np.uint8)
    codes = np.ascontiguousarray(codes, dtype=np.uint32)
    lens = np.ascontiguousarray(lens, dtype=np.uint8)
    f = _lib("oracle_port").port_huffman_decompress
    f.restype = C.c_uint64
    n = f(_p(w, _u32p), C.c_uint64(w.size), C.c_uint64(buffer_size), _p(codes, _u32p), _p(lens, _u8p), _p(out, _u8p), C.c_uint64(cap))
    return out[: min(n, cap)].copy(), n


def port_dfl_frequencies(tokens):
    t = _as_u8(tokens)
    fr = np.zeros(286, dtype=np.uint64)
    _lib("oracle_port").port_dfl_frequencies(_p(t, _u8p), C.c_uint64(t.size), _p(fr, _u64p))
    return fr


def port_dfl_build(freq):
    fr = np.ascontiguousarray(freq, dtype=np.uint64)
    codes = np.zeros(286, dtype=np.uint32)
    lens = np.zeros(286, dtype=np.uint8)
    distinct = _lib("oracle_port").port_dfl_build(_p(fr, _u64p), _p(codes, _u32p), _p(lens, _u8p))
    return codes, lens, distinct


def port_dfl_encode(tokens):
    """One block of byte tokens -> dict(freq, codes, lens, words, bits) of the entropy stage."""
    t = _as_u8(tokens)
    fr = port_dfl_frequencies(t)
    codes, lens, distinct = port_dfl_build(fr)
    words = np.zeros(t.size // 2 + 8, dtype=np.uint32)   # <= 32 bits per 2-byte unit
    f = _lib("oracle_port").port_dfl_encode
    f.restype = C.c_uint64
    bits = int(f(_p(t, _u8p), C.c_uint64(t.size), _p(codes, _u32p), _p(lens, _u8p), _p(words, _u32p)))
    return dict(freq=fr, codes=codes, lens=lens, distinct=distinct, words=words[: (bits + 31) // 32].copy(), bits=bits)


def port_dfl_decode(words, codes, lens, nbytes):
    w = np.ascontiguousarray(words, dtype=np.uint32)
    codes = np.ascontiguousarray(codes, dtype=np.uint32)
    lens = np.ascontiguousarray(lens, dtype=np.uint8)
    out = np.zeros(nbytes + 8, dtype=np.uint8)
    used = C.c_uint64(0)
    rc = _lib("oracle_port").port_dfl_decode(_p(w, _u32p), C.c_uint64(w.size), _p(codes, _u32p), _p(lens, _u8p),
                                             C.c_uint64(nbytes), _p(out, _u8p), C.byref(used))
    if rc:
        raise RuntimeError("deflate entropy stage: corrupt stream")
    return out[:nbytes].copy(), used.value


def port_fse_normalize(freq):
    fr = np.array(freq, dtype=np.uint64)
    _lib("oracle_port").port_fse_normalize(_p(fr, _u64p))
    return fr


def port_fse_tables(norm):
    nm = np.ascontiguousarray(norm, dtype=np.uint64)
    tt = np.zeros(256, dtype=np.uint32)
    enc = np.zeros(256, dtype=np.uint8)
    cum = np.zeros(257, dtype=np.uint16)
    _lib("oracle_port").port_fse_build_tables(_p(nm, _u64p), _p(tt, _u32p), _p(enc, _u8p), _p(cum, _u16p))
    return tt, enc, cum


def port_fse_compress(data):
    """-> (words u64, norm u64[256], total_bits, size_bytes per main.zig:67)"""
    d = _as_u8(data)
    words = np.zeros((16 + 8 * d.size) // 64 + 2, dtype=np.uint64)
    norm = np.zeros(256, dtype=np.uint64)
    tb = C.c_uint64(0)
    f = _lib("oracle_port").port_fse_compress
    f.restype = C.c_uint64
    sz = f(_p(d, _u8p), C.c_uint64(d.size), _p(words, _u64p), _p(norm, _u64p), C.byref(tb))
    return words[: (tb.value + 63) // 64].copy(), norm, tb.value, sz


def port_fse_decompress(words, total_bits, n, norm):
    w = np.concatenate([np.ascontiguousarray(words, dtype=np.uint64), np.zeros(2, dtype=np.uint64)])
    nm = np.ascontiguousarray(norm, dtype=np.uint64)
    out = np.zeros(max(n, 1), dtype=np.uint8)
    rc = _lib("oracle_port").port_fse_decompress(_p(w, _u64p), C.c_uint64(total_bits), C.c_uint64(n), _p(nm, _u64p), _p(out, _u8p))
    return out[:n].copy(), rc


def port_lz77_decompress_blocks(stream, off, block, n, variant, threads=0):
    s = np.concatenate([_as_u8(stream), np.zeros(64, dtype=np.uint8)])
    off = np.ascontiguousarray(off, dtype=np.uint64)
    out = np.zeros(n + 64, dtype=np.uint8)
    f = _lib("oracle_port").port_lz77_decompress_blocks
    f.restype = C.c_uint64
    bad = f(_p(s, _u8p), _p(off, _u64p), C.c_uint64(off.size - 1), C.c_uint64(block), C.c_uint64(n), C.c_int(variant),
            _p(out, _u8p), C.c_int(threads))
    return out[:n], int(bad)


def port_zig_huffman_compress(data):
    """Zig Huffman file format (parity unpinned) -> bytes, or None where the reference's behaviour is undefined."""
    d = _as_u8(data)
    out = np.zeros(d.size + d.size // 4 + 8192 * (d.size // (1 << 22) + 2), dtype=np.uint8)
    f = _lib("oracle_port").port_zig_huffman_compress
    f.restype = C.c_uint64
    n = f(_p(d, _u8p), C.c_uint64(d.size), _p(out, _u8p))
    return out[:n].copy() if n else None


def port_zig_huffman_decompress(stream, max_out):
    s = _as_u8(stream)
    out = np.zeros(max_out + 64, dtype=np.uint8)
    f = _lib("oracle_port").port_zig_huffman_decompress
    f.restype = C.c_uint64
    n = f(_p(s, _u8p), C.c_uint64(s.size), _p(out, _u8p), C.c_uint64(out.size))
    if n == 0xFFFFFFFFFFFFFFFF:
        raise RuntimeError("zig huffman: corrupt stream")
    return out[: min(n, out.size)].copy()


def port_threads():
    import os
    return os.cpu_count() or 1
