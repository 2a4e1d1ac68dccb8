"""CPU tests: the C-ABI libraries load and export every symbol include/*.h declares
(no compute calls without a GPU), and compute calls fail loudly without a device."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "compression_algorithms_b200")

HEADER_TO_LIB = {
    "b200comp.h": "libb200comp.so",
    "b200_huffman.h": "libb200_huffman.so",
    "b200_lz77.h": "libb200_lz77.so",
    "b200_deflate.h": "libb200_deflate.so",
    "b200_fse.h": "libb200_fse.so",
}


def _declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//.*", "", src)
    src = re.sub(r"#.*", "", src)
    src = re.sub(r"typedef\s+struct\s*\w*\s*\{.*?\}\s*\w+\s*;", "", src, flags=re.S)
    names = re.findall(r"\b([A-Za-z_]\w*)\s*\([^;{}]*\)\s*;", src)
    return sorted(set(n for n in names if n not in ("defined",)))


@pytest.mark.parametrize("header", sorted(HEADER_TO_LIB))
def test_exports(header):
    hp = os.path.join(ROOT, "include", header)
    if not os.path.exists(hp):
        pytest.skip("%s not written yet" % header)
    lib = os.path.join(PKG, HEADER_TO_LIB[header])
    assert os.path.exists(lib), "%s not built: python -m compression_algorithms_b200.build" % lib
    C.CDLL(os.path.join(PKG, "libb200comp.so"), mode=C.RTLD_GLOBAL)
    h = C.CDLL(lib)
    names = _declared(header)
    assert names, "no prototypes parsed from %s" % header
    missing = [n for n in names if not hasattr(h, n)]
    assert not missing, "%s does not export %s" % (HEADER_TO_LIB[header], missing)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from compression_algorithms_b200 import _lib
    ctx = C.c_void_p()
    rc = _lib.core().b200_ctx_create(C.byref(ctx), 0, None)
    assert rc != 0
    assert b"no CPU fallback" in _lib.core().b200_last_error() or rc == 1
    from compression_algorithms_b200.device import Context
    with pytest.raises(RuntimeError):
        Context(0)


def test_layout_is_pure_host():
    from compression_algorithms_b200 import device as dv
    L = dv.huffman_layout(1_000_000, 65536)
    assert L.nblocks == 16 and L.nchunks == 245 and L.chunks_per_block == 16
    L = dv.huffman_layout(1_000_000, 0)
    assert L.nblocks == 1 and L.nchunks == 245
    with pytest.raises(RuntimeError):
        dv.huffman_layout(1_000_000, 1000)


def test_corpus_deterministic():
    from compression_algorithms_b200 import corpus
    a = corpus.generate(3_000_000, corpus.ENWIK, 1)
    b = corpus.generate(3_000_000, corpus.ENWIK, 1)
    assert (a == b).all() and a.min() > 0
    assert (corpus.generate(1 << 20, corpus.ENWIK, 1) == a[: 1 << 20]).all()   # prefix property
    assert len(set(bytes(corpus.generate(10000, corpus.ACGT, 1)))) == 4
