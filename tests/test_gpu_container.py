"""GPU: the self-describing containers: one stream = header + serialized tables + decode index +
payload; written in one process, decoded in a fresh one with nothing else; payload words bit-exact with the oracle;
corrupt containers are rejected with B200_ERR_FORMAT before any kernel trusts them."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _corpus(n, kind=0, seed=11):
    from compression_algorithms_b200 import corpus
    return corpus.generate(n, kind, seed)


def _compress(ctx, codec, data, block):
    from compression_algorithms_b200 import _lib
    lib = _lib.core()
    n = data.size
    cap = int((lib.b200_huffman_container_max_bytes if codec == 1 else lib.b200_deflate_container_max_bytes)(n, block))
    out = np.zeros(cap, dtype=np.uint8)
    tot = C.c_uint64(0)
    fn = lib.b200_huffman_compress_container_host if codec == 1 else lib.b200_deflate_compress_container_host
    _lib.check(fn(ctx.handle, data.ctypes.data, n, block, out.ctypes.data, cap, C.byref(tot)))
    return out[: tot.value].copy()


def _decompress(ctx, codec, cont, cap=None):
    from compression_algorithms_b200 import _lib
    lib = _lib.core()
    cd, n, bs = C.c_uint32(0), C.c_uint64(0), C.c_uint64(0)
    rc = lib.b200_container_info(cont.ctypes.data, cont.size, C.byref(cd), C.byref(n), C.byref(bs))
    if rc:
        return rc, None
    out = np.zeros(max(n.value if cap is None else cap, 1) + 64, dtype=np.uint8)
    got = C.c_uint64(0)
    fn = lib.b200_huffman_decompress_container_host if codec == 1 else lib.b200_deflate_decompress_container_host
    rc = fn(ctx.handle, cont.ctypes.data, cont.size, out.ctypes.data, n.value if cap is None else cap, C.byref(got))
    return rc, out[: got.value]


@pytest.mark.parametrize("codec", [1, 2])
@pytest.mark.parametrize("n,block", [(1_000_003, 65536), (1_000_003, 0), (70_000, 65536), (8200, 4096), (300, 0), (2, 0)])
def test_roundtrip_in_memory(ctx, codec, n, block):
    data = _corpus(n).copy()
    data[-1] = data[-2] ^ 1          # every table scope needs two distinct symbols (the reference exits otherwise, U6)
    cont = _compress(ctx, codec, data, block)
    assert cont[:8].tobytes() == b"B200CONT"
    rc, dec = _decompress(ctx, codec, cont)
    assert rc == 0 and np.array_equal(dec, data)


def test_huffman_payload_is_the_reference_stream(ctx, ob):
    """the words inside the container are huffman_compress's words (here: the oracle port), the stored table is the histogram"""
    data = _corpus(500_000, 0, 3)
    cont = _compress(ctx, 1, data, 0)
    hdr = cont[:64].view(np.uint64)
    assert int(hdr[1]) == 1 | (1 << 32) and int(hdr[2]) == data.size and int(hdr[3]) == 0 and int(hdr[4]) == 1
    exp = ob.port_huffman_compress(data)
    nw = exp["word_idx"] + (1 if exp["bit_idx"] else 0)
    assert int(hdr[6]) == nw
    assert np.array_equal(cont[64: 64 + 1024].view(np.uint32), np.bincount(data, minlength=256).astype(np.uint32))
    assert np.array_equal(cont[cont.size - ((nw * 4 + 7) // 8) * 8:][: nw * 4].view(np.uint32), exp["words"])


def test_deflate_container_is_smaller_than_tokens(ctx):
    data = _corpus(4_000_000, 0, 5)
    cont = _compress(ctx, 2, data, 65536)
    assert cont.size < data.size * 0.70      # raw byte tokens are ~1.26 x the input; the whole container ~0.63 x


@pytest.mark.parametrize("codec", [1, 2])
def test_corrupt_containers_are_rejected(ctx, codec):
    data = _corpus(200_000, 0, 9)
    good = _compress(ctx, codec, data, 65536)
    bad = good.copy(); bad[0] ^= 1                                  # magic
    assert _decompress(ctx, codec, bad)[0] == 5
    bad = good.copy(); bad[8] = 9                                   # version
    assert _decompress(ctx, codec, bad)[0] == 5
    assert _decompress(ctx, codec, good[: good.size - 4096].copy())[0] == 5      # truncated
    bad = good.copy(); bad[32:40] = np.frombuffer(np.uint64(77).tobytes(), np.uint8)   # block count
    assert _decompress(ctx, codec, bad)[0] == 5
    # a chunk bit count that does not fit its chunk / the stream
    nblocks = 4                                                     # ceil(200000 / 65536)
    o_cbits = 64 + nblocks * (256 if codec == 1 else 288) * 4 + (0 if codec == 1 else nblocks * 8)
    bad = good.copy(); bad[o_cbits: o_cbits + 4] = np.frombuffer(np.uint32(0x7FFFFFFF).tobytes(), np.uint8)
    assert _decompress(ctx, codec, bad)[0] == 5
    bad = good.copy(); bad[o_cbits: o_cbits + 4] = np.frombuffer(np.uint32(8).tobytes(), np.uint8)
    assert _decompress(ctx, codec, bad)[0] == 5
    # wrong codec for the call, too small an output buffer
    assert _decompress(ctx, 3 - codec, good)[0] == 5
    assert _decompress(ctx, codec, good, cap=1000)[0] == 3
    rc, dec = _decompress(ctx, codec, good)
    assert rc == 0 and np.array_equal(dec, data)                    # the context survived all of the above


_CHILD = r"""
import ctypes as C, sys
lib = C.CDLL(sys.argv[1])
fn = getattr(lib, sys.argv[2]); fn.restype = C.c_uint64
args = [a.encode() for a in sys.argv[3:5]]
if sys.argv[2] == "huffman_compress_file":
    fn.argtypes = [C.c_char_p, C.c_char_p, C.c_uint64]; print(fn(args[0], args[1], int(sys.argv[5])))
else:
    fn.argtypes = [C.c_char_p, C.c_char_p]; print(fn(args[0], args[1]))
"""


@pytest.mark.parametrize("shim,enc,dec,extra", [("huffman", "huffman_compress_file", "huffman_decompress_file", ["0"]),
                                                ("huffman", "huffman_compress_file", "huffman_decompress_file", ["65536"]),
                                                ("deflate", "compress_entropy", "decompress_entropy", [])])
def test_file_written_by_the_shim_decodes_in_a_fresh_process(ctx, tmp_path, shim, enc, dec, extra):
    data = _corpus(3_000_017, 0, 21)
    src, packed, back = tmp_path / "in.bin", tmp_path / "in.b200", tmp_path / "out.bin"
    data.tofile(src)
    lib = os.path.join(ROOT, "compression_algorithms_b200", "libb200_%s.so" % shim)
    r = subprocess.run([sys.executable, "-c", _CHILD, lib, enc, str(src), str(packed)] + extra, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert int(r.stdout.strip().splitlines()[-1]) == os.path.getsize(packed) < data.size
    r = subprocess.run([sys.executable, "-c", _CHILD, lib, dec, str(packed), str(back)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert int(r.stdout.strip().splitlines()[-1]) == data.size
    assert np.array_equal(np.fromfile(back, dtype=np.uint8), data)


def test_huffman_shim_keeps_several_streams_alive(ctx):
    """the decode index of EVERY stream the drop-in produced stays reachable (no 'last stream only' global):
    compress A, compress B, decode A, decode B"""
    lib = C.CDLL(os.path.join(ROOT, "compression_algorithms_b200", "libb200_huffman.so"))

    class BitWriter(C.Structure):
        _fields_ = [("buffer", C.POINTER(C.c_uint32)), ("bit_idx", C.c_uint64), ("word_idx", C.c_uint64), ("buffer_size", C.c_uint64)]

    class Node(C.Structure):
        pass
    Node._fields_ = [("value", C.c_uint8), ("frequency", C.c_uint32), ("left", C.POINTER(Node)), ("right", C.POINTER(Node))]
    lib.huffman_compress.restype = Node
    lib.huffman_compress.argtypes = [C.c_char_p, C.c_uint64, C.POINTER(BitWriter)]
    lib.huffman_decompress.argtypes = [C.POINTER(BitWriter), C.POINTER(Node), C.c_char_p, C.POINTER(C.c_uint64)]
    datas = [_corpus(300_000, 0, 1), _corpus(200_000, 1, 2), _corpus(123_457, 3, 3)]
    streams = []
    for d in datas:
        w = BitWriter()
        root = lib.huffman_compress(d.tobytes(), d.size, C.byref(w))
        streams.append((w, root))
    for d, (w, root) in zip(datas, streams):
        out = C.create_string_buffer(d.size + 64)
        n = C.c_uint64(d.size + 64)
        lib.huffman_decompress(C.byref(w), C.byref(root), out, C.byref(n))
        assert d.size <= n.value <= d.size + 8
        assert np.array_equal(np.frombuffer(out.raw[: d.size], dtype=np.uint8), d)
