"""ctypes loader for the in-tree native libraries. Fails loudly: there is no
Python or CPU fallback for any codec."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)
vp = C.c_void_p


class HuffLayout(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "bytes", "nblocks", "nchunks", "chunks_per_block", "off_freq", "off_codes", "off_lens", "off_tree",
        "off_meta", "off_block_bits", "off_block_word", "off_chunk_bits", "off_chunk_off", "off_sub_off")]


class DflLayout(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "bytes", "nblocks", "nchunks", "chunks_per_block", "off_freq", "off_codes", "off_lens", "off_tree", "off_meta",
        "off_tok_off", "off_tok_sizes", "off_block_bits", "off_block_word", "off_chunk_state", "off_chunk_bits",
        "off_chunk_off", "off_sub_off")]


class FseLayout(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "bytes", "nblocks", "nsegs", "segs_per_block", "off_freq", "off_norm", "off_tt", "off_seg_bits", "off_seg_word")]


_core = None
_corpus = None


def _path(name):
    return os.path.join(_HERE, name)


def core():
    """libb200comp.so (CUDA). Raises if it has not been built."""
    global _core
    if _core is None:
        p = _path("libb200comp.so")
        if not os.path.exists(p):
            raise RuntimeError(
                "%s is missing: run `python -m compression_algorithms_b200.build` (there is no CPU fallback)" % p)
        lib = C.CDLL(p, mode=C.RTLD_GLOBAL)
        lib.b200_last_error.restype = C.c_char_p
        lib.b200_ctx_launches.restype = C.c_uint64
        lib.b200_ctx_launches.argtypes = [vp]
        lib.b200_ctx_create.argtypes = [C.POINTER(vp), C.c_int, vp]
        lib.b200_ctx_destroy.argtypes = [vp]
        lib.b200_ctx_sync.argtypes = [vp]
        lib.b200_huffman_layout.argtypes = [C.c_uint64, C.c_uint64, C.POINTER(HuffLayout)]
        lib.b200_huffman_max_words.restype = C.c_uint64
        lib.b200_huffman_max_words.argtypes = [C.c_uint64, C.c_uint64]
        lib.b200_huffman_encode_dev.argtypes = [vp, vp, C.c_uint64, C.c_uint64, vp, C.c_uint64, vp, C.c_uint64, u64p, u32p]
        lib.b200_huffman_tables_dev.argtypes = [vp, vp, C.c_uint64, C.c_uint64, vp, C.c_uint64]
        lib.b200_huffman_histogram_dev.argtypes = [vp, vp, C.c_uint64, vp]
        lib.b200_huffman_encode_with_freq_dev.argtypes = [vp, vp, C.c_uint64, vp, vp, C.c_uint64, vp, C.c_uint64, u64p, u64p, u32p]
        lib.b200_huffman_splice_dev.argtypes = [vp, vp, C.c_uint64, C.c_uint64, vp, C.c_uint64]
        lib.b200_huffman_decode_dev.argtypes = [vp, vp, C.c_uint64, vp, C.c_uint64, C.c_uint64, C.c_uint64, vp]
        lib.b200_huffman_decode_serial_dev.argtypes = [vp, vp, C.c_uint64, C.c_uint64, vp, vp, vp, C.c_uint64, u64p]
        lib.b200_lz77_block_stride.restype = C.c_uint64
        lib.b200_lz77_block_stride.argtypes = [C.c_uint64]
        lib.b200_lz77_encode_dev.argtypes = [vp, C.c_int, vp, C.c_uint64, C.c_uint64, vp, C.c_uint64, vp, vp, u64p]
        lib.b200_lz77_decode_dev.argtypes = [vp, C.c_int, vp, vp, vp, C.c_uint64, C.c_uint64, vp]
        lib.b200_lz77_encode_debug_dev.argtypes = [vp, C.c_int, vp, C.c_uint64, C.c_uint64, vp, C.c_uint64, vp, vp, u64p, vp]
        lib.b200_ctx_set_timing.argtypes = [vp, C.c_int]
        lib.b200_ctx_timing_count.argtypes = [vp]
        lib.b200_ctx_timing_get.argtypes = [vp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_float)]
        lib.b200_lz77_max_bytes.restype = C.c_uint64
        lib.b200_lz77_max_bytes.argtypes = [C.c_int, C.c_uint64, C.c_uint64]
        lib.b200_lz77_compress_host.argtypes = [vp, C.c_int, vp, C.c_uint64, C.c_uint64, vp, C.c_uint64, vp, vp, u64p]
        lib.b200_lz77_decompress_host.argtypes = [vp, C.c_int, vp, C.c_uint64, vp, vp, C.c_uint64, C.c_uint64, vp]
        lib.b200_huffman_compress_host.argtypes = [vp, vp, C.c_uint64, C.c_uint64, vp, C.c_uint64, vp, C.c_uint64, u64p, u32p]
        lib.b200_huffman_decompress_host.argtypes = [vp, vp, C.c_uint64, vp, C.c_uint64, C.c_uint64, C.c_uint64, vp]
        lib.b200_huffman_decompress_serial_host.argtypes = [vp, vp, C.c_uint64, C.c_uint64, vp, vp, vp, C.c_uint64, u64p]
        lib.b200_dfl_layout_for.argtypes = [C.c_uint64, C.c_uint64, C.POINTER(DflLayout)]
        lib.b200_dfl_max_words.restype = C.c_uint64
        lib.b200_dfl_max_words.argtypes = [C.c_uint64, C.c_uint64]
        lib.b200_dfl_encode_dev.argtypes = [vp, vp, C.c_uint64, vp, vp, C.c_uint64, C.c_uint64, vp, C.c_uint64, vp, C.c_uint64, u64p, u32p]
        lib.b200_dfl_decode_dev.argtypes = [vp, vp, C.c_uint64, vp, C.c_uint64, C.c_uint64, C.c_uint64, vp]
        lib.b200_deflate_compress_dev.argtypes = [vp, vp, C.c_uint64, C.c_uint64, vp, C.c_uint64, vp, vp, vp, C.c_uint64, vp, C.c_uint64, u64p, u32p]
        lib.b200_deflate_decompress_dev.argtypes = [vp, vp, C.c_uint64, vp, C.c_uint64, C.c_uint64, C.c_uint64, vp, vp]
        lib.b200_deflate_compress_host.argtypes = [vp, vp, C.c_uint64, C.c_uint64, vp, C.c_uint64, vp, C.c_uint64, u64p, u32p]
        lib.b200_deflate_decompress_host.argtypes = [vp, vp, C.c_uint64, vp, C.c_uint64, C.c_uint64, C.c_uint64, vp]
        for name in ("b200_huffman_container_max_bytes", "b200_deflate_container_max_bytes"):
            getattr(lib, name).restype = C.c_uint64
            getattr(lib, name).argtypes = [C.c_uint64, C.c_uint64]
        lib.b200_container_info.argtypes = [vp, C.c_uint64, u32p, u64p, u64p]
        for name in ("b200_huffman_compress_container_host", "b200_deflate_compress_container_host"):
            getattr(lib, name).argtypes = [vp, vp, C.c_uint64, C.c_uint64, vp, C.c_uint64, u64p]
        for name in ("b200_huffman_decompress_container_host", "b200_deflate_decompress_container_host"):
            getattr(lib, name).argtypes = [vp, vp, C.c_uint64, vp, C.c_uint64, u64p]
        lib.b200_multi_create.argtypes = [C.POINTER(vp), C.POINTER(C.c_int), C.c_int]
        lib.b200_multi_destroy.argtypes = [vp]
        lib.b200_multi_device_count.argtypes = [vp]
        lib.b200_multi_allgathers.restype = C.c_uint64
        lib.b200_multi_allgathers.argtypes = [vp]
        lib.b200_multi_launches.restype = C.c_uint64
        lib.b200_multi_launches.argtypes = [vp]
        lib.b200_lz77_compress_multi_host.argtypes = [vp, C.c_int, vp, C.c_uint64, C.c_uint64, vp, C.c_uint64, vp, vp, u64p]
        lib.b200_lz77_decompress_multi_host.argtypes = [vp, C.c_int, vp, C.c_uint64, vp, vp, C.c_uint64, C.c_uint64, vp]
        lib.b200_zig_huffman_max_bytes.restype = C.c_uint64
        lib.b200_zig_huffman_max_bytes.argtypes = [C.c_uint64]
        lib.b200_zig_huffman_compress_host.argtypes = [vp, vp, C.c_uint64, vp, C.c_uint64, u64p]
        lib.b200_zig_huffman_decompress_host.argtypes = [vp, vp, C.c_uint64, vp, C.c_uint64, u64p]
        if hasattr(lib, "b200_fse_layout_for"):
            lib.b200_fse_layout_for.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.POINTER(FseLayout)]
            lib.b200_fse_max_words.restype = C.c_uint64
            lib.b200_fse_max_words.argtypes = [C.c_uint64, C.c_uint64]
            lib.b200_fse_encode_dev.argtypes = [vp, vp, C.c_uint64, C.c_uint64, C.c_uint64, vp, C.c_uint64, vp, C.c_uint64, u64p]
            lib.b200_fse_decode_dev.argtypes = [vp, vp, vp, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, vp, u32p]
            lib.b200_fse_normalize_dev.argtypes = [vp, vp, C.c_uint64, C.c_uint64, vp, C.c_uint64]
            lib.b200_fse_container_max_words.restype = C.c_uint64
            lib.b200_fse_container_max_words.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64]
            lib.b200_fse_compress_host.argtypes = [vp, vp, C.c_uint64, C.c_uint64, C.c_uint64, vp, C.c_uint64, u64p]
            lib.b200_fse_decompress_host.argtypes = [vp, vp, C.c_uint64, vp, C.c_uint64, u64p]
        _core = lib
    return _core


def corpus():
    """libb200corpus.so (host only)."""
    global _corpus
    if _corpus is None:
        p = _path("libb200corpus.so")
        if not os.path.exists(p):
            raise RuntimeError("%s is missing: run `python -m compression_algorithms_b200.build`" % p)
        lib = C.CDLL(p)
        lib.b200_corpus_generate.argtypes = [vp, C.c_uint64, C.c_int, C.c_uint64]
        lib.b200_corpus_generate_range.argtypes = [vp, C.c_uint64, C.c_uint64, C.c_int, C.c_uint64]
        _corpus = lib
    return _corpus


def check(rc):
    if rc != 0:
        raise RuntimeError("b200comp error %d: %s" % (rc, core().b200_last_error().decode("utf8", "replace")))
