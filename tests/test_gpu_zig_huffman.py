"""GPU: the Zig-Huffman-compatible chunked mode against its CPU restatement
(oracle/port/zig_huffman_port.c; PARITY UNPINNED -- no Zig toolchain exists here): the file bytes
(pre-order tree dumps, CompressedSize headers, MSB-first byte payload) and the decoder's output."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
CH = 1 << 22


def _gpu_compress(ctx, data):
    from compression_algorithms_b200 import _lib
    lib = _lib.core()
    cap = int(lib.b200_zig_huffman_max_bytes(data.size))
    out = np.zeros(cap, dtype=np.uint8)
    tot = C.c_uint64(0)
    rc = lib.b200_zig_huffman_compress_host(ctx.handle, data.ctypes.data, data.size, out.ctypes.data, cap, C.byref(tot))
    return rc, out[: tot.value].copy()


def _gpu_decompress(ctx, stream, max_out):
    from compression_algorithms_b200 import _lib
    lib = _lib.core()
    out = np.zeros(max_out + 64, dtype=np.uint8)
    n = C.c_uint64(0)
    rc = lib.b200_zig_huffman_decompress_host(ctx.handle, stream.ctypes.data, stream.size, out.ctypes.data, out.size, C.byref(n))
    return rc, out[: n.value].copy()


@pytest.mark.parametrize("kind", [0, 3])
@pytest.mark.parametrize("n", [100_000, CH - 1, CH, CH + 12_345, 2 * CH, 3 * CH + 1])
def test_file_bytes_and_decoder_equal_the_port(ctx, ob, kind, n):
    from compression_algorithms_b200 import corpus
    data = corpus.generate(n, kind, 17)
    exp = ob.port_zig_huffman_compress(data)
    assert exp is not None
    rc, got = _gpu_compress(ctx, data)
    assert rc == 0
    assert got.size == exp.size and np.array_equal(got, exp), "file differs at byte %d" % int(np.nonzero(got[: min(got.size, exp.size)] != exp[: min(got.size, exp.size)])[0][:1].sum())
    nchunks = n // CH + 1                                   # an exact multiple ends with an empty chunk (the read that finds the end)
    dec_exp = ob.port_zig_huffman_decompress(exp, nchunks * CH)
    rc, dec = _gpu_decompress(ctx, got, nchunks * CH)
    assert rc == 0 and np.array_equal(dec, dec_exp)
    # the format drops the partial last byte of every chunk (main.zig:523), so later chunks may start a symbol early:
    # the first chunk's prefix must come back, the rest is pinned by the equality with the port above
    hi = min(CH, n)
    assert np.array_equal(dec[: hi - 8], data[: hi - 8]) and abs(int(dec.size) - n) <= 2 * nchunks


def test_undefined_inputs_are_refused(ctx):
    """one distinct symbol in a chunk: the reference shifts a u32 by 32 bits (main.zig:222) -> B200_ERR_DOMAIN"""
    rc, _ = _gpu_compress(ctx, np.zeros(CH, dtype=np.uint8))
    assert rc == 4


def test_corrupt_file_is_rejected(ctx):
    from compression_algorithms_b200 import corpus
    data = corpus.generate(200_000, 0, 5)
    rc, good = _gpu_compress(ctx, data)
    assert rc == 0
    assert _gpu_decompress(ctx, good[: good.size // 2].copy(), CH)[0] == 5
    bad = good.copy(); bad[0:4] = 0xFF
    assert _gpu_decompress(ctx, bad, CH)[0] == 5
