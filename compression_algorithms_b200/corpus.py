"""Seeded synthetic inputs (no network for enwik8/9): SURVEY.md §8d."""
import ctypes as C

import numpy as np

from . import _lib

ENWIK = 0        # enwik-shaped Zipf-word text with wiki/XML markup and some UTF-8
ACGT = 1         # 4-symbol i.i.d.
SKEWED = 2       # 2 symbols, P = 0.95
RANDOM = 3       # uniform bytes 1..255
DEFAULT_SEED = 20261018


def generate(n, kind=ENWIK, seed=DEFAULT_SEED, out=None):
    """Returns a numpy uint8 array of n bytes (or fills `out`, e.g. a pinned tensor's numpy view)."""
    a = np.empty(n, dtype=np.uint8) if out is None else out
    assert a.dtype == np.uint8 and a.size >= n and a.flags["C_CONTIGUOUS"]
    rc = _lib.corpus().b200_corpus_generate(a.ctypes.data_as(C.c_void_p), n, kind, seed)
    if rc:
        raise ValueError("b200_corpus_generate failed: %d" % rc)
    return a[:n]


def generate_range(start, length, kind=ENWIK, seed=DEFAULT_SEED, out=None):
    """Bytes [start, start + length) of the buffer generate(N, kind, seed) gives for any N >= start + length:
    every rank of a multi-GPU run materialises only its own shard of one global buffer."""
    a = np.empty(length, dtype=np.uint8) if out is None else out
    assert a.dtype == np.uint8 and a.size >= length and a.flags["C_CONTIGUOUS"]
    rc = _lib.corpus().b200_corpus_generate_range(a.ctypes.data_as(C.c_void_p), start, length, kind, seed)
    if rc:
        raise ValueError("b200_corpus_generate_range failed: %d" % rc)
    return a[:length]
