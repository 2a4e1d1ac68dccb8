"""Wall time of the standalone-LZ77 drop-in entry points (libb200_lz77.so: lz77_compress / lz77_decompress on a whole buffer,
what the reference's algorithms/lz77/main.c calls): python tools/dropin_time.py [MB]"""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, '.')
from compression_algorithms_b200 import corpus
n = (int(sys.argv[1]) if len(sys.argv) > 1 else 32) * 1_000_000
here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "compression_algorithms_b200")
lib = C.CDLL(os.path.join(here, "libb200_lz77.so"))
class BitStream(C.Structure):
    _fields_ = [("data", C.POINTER(C.c_uint8)), ("bit_index", C.c_uint64)]
lib.lz77_compress.restype = C.POINTER(BitStream); lib.lz77_compress.argtypes = [C.c_char_p, C.c_uint64]
lib.lz77_decompress.restype = C.POINTER(C.c_uint8); lib.lz77_decompress.argtypes = [C.POINTER(BitStream), C.c_uint64, C.POINTER(C.c_uint64)]
data = corpus.generate(n, 0, 7)
raw = data.tobytes()
lib.lz77_compress(raw[:100000], 100000)   # warm-up: context, kernels
for rep in range(2):
    t0 = time.time(); st = lib.lz77_compress(raw, n); t1 = time.time()
    bits = st.contents.bit_index
    out_n = C.c_uint64(0)
    t2 = time.time(); dec = lib.lz77_decompress(st, n, C.byref(out_n)); t3 = time.time()
    st.contents.bit_index = bits
    ok = bool(np.array_equal(np.ctypeslib.as_array(dec, shape=(n,)), data))
    print("drop-in lz77 %d MB: lz77_compress %.3f s (%.3f GB/s), lz77_decompress %.4f s (%.2f GB/s), ok %s" % (n // 1_000_000, t1 - t0, n / (t1 - t0) / 1e9, t3 - t2, n / (t3 - t2) / 1e9, ok))
