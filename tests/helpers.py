import numpy as np


def fnv1a64(b):
    h = 0xCBF29CE484222325
    for x in bytes(b):
        h = ((h ^ x) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return h


def lcg_bytes(mode, n=200000):
    """generator of SURVEY.md §4.3 'checksum vectors'"""
    s = 12345
    out = bytearray(n)
    for i in range(n):
        s = (s * 1664525 + 1013904223) & 0xFFFFFFFF
        r = s >> 24
        out[i] = b"acgt"[r & 3] if mode == "A" else ord("a") + r % 26
    return bytes(out)


def u32(t):
    """torch int32 tensor -> numpy uint32"""
    return t.detach().cpu().numpy().view(np.uint32)


def first_diff(a, b):
    n = min(len(a), len(b))
    d = np.nonzero(np.asarray(a[:n]) != np.asarray(b[:n]))[0]
    return int(d[0]) if d.size else (n if len(a) != len(b) else -1)
