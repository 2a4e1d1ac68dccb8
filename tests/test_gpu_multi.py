"""GPU, needs >= 2 devices (skipped on a one-GPU box; run with `gpurun --gpus 2`): the C multi-GPU entry points
(csrc/multi.cu: one host thread per GPU, ncclAllGather of the shard sizes) give byte for byte the single-device
stream, and the deflate drop-in's compress()/decompress() use every GPU."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ndev():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("variant", [1, 0])
@pytest.mark.parametrize("n", [40_000_003, 65536 * 3 + 5, 1000])
def test_multi_equals_single(ctx, variant, n):
    if _ndev() < 2:
        pytest.skip("needs two GPUs")
    from compression_algorithms_b200 import _lib, corpus
    lib = _lib.core()
    G = min(_ndev(), 8)
    m = C.c_void_p()
    _lib.check(lib.b200_multi_create(C.byref(m), None, G))
    try:
        data = corpus.generate(n, 0, 13)
        nb = (n + 65535) // 65536
        cap = int(lib.b200_lz77_max_bytes(variant, n, 65536))
        outs = []
        for multi in (False, True):
            out = np.zeros(cap, dtype=np.uint8); sizes = np.zeros(nb, dtype=np.uint64); off = np.zeros(nb + 1, dtype=np.uint64); tot = C.c_uint64(0)
            if multi:
                _lib.check(lib.b200_lz77_compress_multi_host(m, variant, data.ctypes.data, n, 65536, out.ctypes.data, cap, sizes.ctypes.data, off.ctypes.data, C.byref(tot)))
            else:
                _lib.check(lib.b200_lz77_compress_host(ctx.handle, variant, data.ctypes.data, n, 65536, out.ctypes.data, cap, sizes.ctypes.data, off.ctypes.data, C.byref(tot)))
            outs.append((out[: tot.value].copy(), sizes, off, tot.value))
        (a, sa, oa, ta), (b, sb, ob_, tb) = outs
        assert ta == tb and np.array_equal(sa, sb) and np.array_equal(oa, ob_)
        if variant == 1:
            assert np.array_equal(a, b)
        else:   # bit tokens: pad bits of every block's last byte are undefined (U3); compare whole bytes per block
            for k in range(nb):
                nbits = int(sa[k])
                assert np.array_equal(a[int(oa[k]): int(oa[k]) + nbits // 8], b[int(oa[k]): int(oa[k]) + nbits // 8])
        assert int(lib.b200_multi_allgathers(m)) >= 1          # the one collective ran
        dec = np.zeros(n + 64, dtype=np.uint8)
        _lib.check(lib.b200_lz77_decompress_multi_host(m, variant, b.ctypes.data, tb, ob_.ctypes.data, sb.ctypes.data, n, 65536, dec.ctypes.data))
        assert np.array_equal(dec[:n], data)
    finally:
        lib.b200_multi_destroy(m)


def test_deflate_dropin_uses_every_gpu(ctx, ob, tmp_path):
    if _ndev() < 2:
        pytest.skip("needs two GPUs")
    from compression_algorithms_b200 import corpus
    data = corpus.generate(64 * 65536 * _ndev() + 777, 0, 3)
    path = tmp_path / "enwik_synth"
    data.tofile(path)
    child = ("import ctypes as C, sys\nlib = C.CDLL(sys.argv[1])\nclass S(C.Structure):\n    _fields_=[('t',C.c_void_p),('h',C.c_void_p),('n',C.c_char_p)]\n"
             "lib.compress.restype = S; lib.compress.argtypes=[C.c_char_p]\ns = lib.compress(sys.argv[2].encode())\n"
             "lib.decompress.argtypes=[C.c_void_p, C.c_char_p]\nlib.decompress(None, s.n)\n")
    r = subprocess.run([sys.executable, "-c", child, os.path.join(ROOT, "compression_algorithms_b200", "libb200_deflate.so"), str(path)],
                       capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 0, r.stdout + r.stderr
    blocks, sizes = ob.port_lz77_compress_blocks(data, 65536, 1)
    expect = np.concatenate([blocks[b, : int(sizes[b])] for b in range(len(sizes))])
    assert np.array_equal(np.fromfile(tmp_path / "enwik_synth.deflate", dtype=np.uint8), expect)
    assert np.array_equal(np.fromfile(tmp_path / "enwik_synth.deflate.out", dtype=np.uint8), data)
