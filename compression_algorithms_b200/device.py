"""Device-resident Python mirror of include/b200comp.h.

torch is used for device memory and streams only; every codec call goes through the
C-ABI of libb200comp.so (ctypes). There is no fallback path: without the library or
without a CUDA device these functions raise.
"""
import ctypes as C
from dataclasses import dataclass, field

import numpy as np
import torch

from . import _lib

LZ_STANDALONE = 0  # algorithms/lz77    : W=2^14, MAX_LEN 15, LSB-first bit tokens
LZ_DEFLATE = 1     # algorithms/deflate : W=2^15, MAX_LEN 31, byte tokens
DEFAULT_BLOCK = 65536  # BUFFER_SIZE, /root/reference/algorithms/deflate/deflate.h:8


class Context:
    """A b200_ctx bound to torch's current CUDA stream on `device`."""

    def __init__(self, device=0):
        if not torch.cuda.is_available():
            raise RuntimeError("compression_algorithms_b200 needs a CUDA device (no CPU fallback)")
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        self._h = C.c_void_p()
        # torch's default stream has handle 0; the C-ABI reads NULL as "make a private
        # stream", so name the legacy default stream explicitly (cudaStreamLegacy == 0x1)
        stream = torch.cuda.current_stream(self.device).cuda_stream or 1
        _lib.check(_lib.core().b200_ctx_create(C.byref(self._h), device, C.c_void_p(stream)))

    @property
    def handle(self):
        return self._h

    def sync(self):
        _lib.check(_lib.core().b200_ctx_sync(self._h))

    def set_timing(self, enable=True):
        _lib.check(_lib.core().b200_ctx_set_timing(self._h, 1 if enable else 0))

    def timings(self):
        """[(kind, ms)] of the dominant kernels launched since set_timing(); kinds:
        0 LZ77 parse, 1 LZ77 decode, 2 Huffman encode, 3 Huffman decode, 4 FSE encode, 5 FSE decode"""
        out = []
        for i in range(_lib.core().b200_ctx_timing_count(self._h)):
            k, ms = C.c_int(0), C.c_float(0)
            _lib.check(_lib.core().b200_ctx_timing_get(self._h, i, C.byref(k), C.byref(ms)))
            out.append((k.value, ms.value))
        return out

    @property
    def launches(self):
        return int(_lib.core().b200_ctx_launches(self._h))

    def close(self):
        if self._h:
            _lib.core().b200_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _ptr(t):
    return C.c_void_p(t.data_ptr())


def _check_u8(t):
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.uint8 and t.is_contiguous()):
        raise TypeError("expected a contiguous CUDA uint8 tensor")


# ------------------------------------------------------------------ Huffman
@dataclass
class HuffmanStream:
    words: torch.Tensor        # uint32 view as int32 storage [capacity]
    side: torch.Tensor         # uint8 side buffer (tables + decode index)
    layout: _lib.HuffLayout
    n: int
    block_size: int            # 0 = one table for the whole buffer
    total_words: int = -1
    worst_status: int = 0

    def _side_view(self, off, count, dtype):
        nbytes = count * torch.empty((), dtype=dtype).element_size()
        return self.side[off: off + nbytes].view(dtype)

    def codes(self):
        return self._side_view(self.layout.off_codes, self.layout.nblocks * 256, torch.int32).view(-1, 256)

    def lens(self):
        return self.side[self.layout.off_lens: self.layout.off_lens + self.layout.nblocks * 256].view(-1, 256)

    def freq(self):
        return self._side_view(self.layout.off_freq, self.layout.nblocks * 256, torch.int32).view(-1, 256)

    def meta(self):
        return self._side_view(self.layout.off_meta, self.layout.nblocks * 4, torch.int32).view(-1, 4)

    def block_bits(self):
        return self._side_view(self.layout.off_block_bits, self.layout.nblocks, torch.int64)

    def block_word(self):
        return self._side_view(self.layout.off_block_word, self.layout.nblocks + 1, torch.int64)


def huffman_layout(n, block_size):
    L = _lib.HuffLayout()
    _lib.check(_lib.core().b200_huffman_layout(n, block_size, C.byref(L)))
    return L


def huffman_alloc(ctx, n, block_size):
    L = huffman_layout(n, block_size)
    cap = int(_lib.core().b200_huffman_max_words(n, block_size))
    words = torch.empty(cap, dtype=torch.int32, device=ctx.device)
    side = torch.empty(L.bytes, dtype=torch.uint8, device=ctx.device)
    return HuffmanStream(words=words, side=side, layout=L, n=n, block_size=block_size)


def huffman_encode(ctx, data, block_size=0, stream=None, sync=True):
    """build_huffman_tree + gather_codes + _huffman_compress per block
    (/root/reference/algorithms/huffman/huffman.c:179-285), device resident."""
    _check_u8(data)
    n = data.numel()
    st = stream if stream is not None else huffman_alloc(ctx, n, block_size)
    tw = C.c_uint64(0)
    ws = C.c_uint32(0)
    _lib.check(_lib.core().b200_huffman_encode_dev(
        ctx.handle, _ptr(data), n, block_size, _ptr(st.words), st.words.numel(), _ptr(st.side), st.side.numel(),
        C.byref(tw) if sync else None, C.byref(ws) if sync else None))
    if sync:
        st.total_words = tw.value
        st.worst_status = ws.value
    return st


def huffman_tables(ctx, data, block_size=0):
    _check_u8(data)
    n = data.numel()
    st = huffman_alloc(ctx, n, block_size)
    _lib.check(_lib.core().b200_huffman_tables_dev(ctx.handle, _ptr(data), n, block_size, _ptr(st.side), st.side.numel()))
    return st


def huffman_histogram(ctx, data):
    """Shard histogram as int64[256] on the device (huffman.c:184-187), ready for an all-reduce."""
    _check_u8(data)
    freq = torch.empty(256, dtype=torch.int64, device=ctx.device)
    _lib.check(_lib.core().b200_huffman_histogram_dev(ctx.handle, _ptr(data), data.numel(), _ptr(freq)))
    return freq


def huffman_encode_with_freq(ctx, data, freq64):
    """Table from the given (global) int64[256] histogram, this shard packed from bit 0 of its own words.
    -> (HuffmanStream with block_size 0, bits of the shard's stream)."""
    _check_u8(data)
    n = data.numel()
    st = huffman_alloc(ctx, n, 0)
    # the table comes from the GLOBAL histogram: a shard whose own distribution differs (binary data on one rank,
    # text on the others) can need up to the reference's 32 bits per symbol, not the 8 of a self-built table
    st.words = torch.empty(n + 8, dtype=torch.int32, device=ctx.device)
    tw, tb, ws = C.c_uint64(0), C.c_uint64(0), C.c_uint32(0)
    _lib.check(_lib.core().b200_huffman_encode_with_freq_dev(
        ctx.handle, _ptr(data), n, _ptr(freq64), _ptr(st.words), st.words.numel(), _ptr(st.side), st.side.numel(),
        C.byref(tw), C.byref(tb), C.byref(ws)))
    st.total_words, st.worst_status = tw.value, ws.value
    return st, tb.value


def huffman_splice(ctx, dst_words, dst_bit, src_words, src_bits):
    """OR `src_bits` bits of src_words (MSB first) into the zeroed dst_words at bit `dst_bit`."""
    _lib.check(_lib.core().b200_huffman_splice_dev(ctx.handle, _ptr(dst_words), dst_words.numel(), dst_bit, _ptr(src_words), src_bits))


def huffman_decode(ctx, st, out=None):
    """Table-lookup decoder (replaces huffman_decompress, huffman.c:330-364)."""
    if out is None:
        out = torch.empty(st.n, dtype=torch.uint8, device=ctx.device)
    _lib.check(_lib.core().b200_huffman_decode_dev(
        ctx.handle, _ptr(st.words), max(st.total_words, 0) if st.total_words >= 0 else st.words.numel(),
        _ptr(st.side), st.side.numel(), st.n, st.block_size, _ptr(out)))
    return out


def huffman_decode_serial(ctx, words, buffer_size, codes, lens, out_capacity):
    """Index-free decode of one foreign stream with the reference's termination rule."""
    out = torch.zeros(out_capacity, dtype=torch.uint8, device=ctx.device)
    cnt = C.c_uint64(0)
    _lib.check(_lib.core().b200_huffman_decode_serial_dev(
        ctx.handle, _ptr(words), words.numel(), buffer_size, _ptr(codes), _ptr(lens), _ptr(out), out_capacity, C.byref(cnt)))
    return out, cnt.value


# ------------------------------------------------------------------ LZ77
@dataclass
class Lz77Stream:
    variant: int
    out: torch.Tensor          # uint8 compacted token stream
    block_sizes: torch.Tensor  # int64[nblocks]: bit_index (variant 0) or bytes (variant 1)
    block_off: torch.Tensor    # int64[nblocks+1]: byte offsets into out
    n: int
    block_size: int
    total_bytes: int = -1


def lz77_alloc(ctx, n, block_size, variant):
    bs = n if (block_size == 0 or block_size > n) else block_size
    nblocks = max(1, (n + bs - 1) // max(bs, 1))
    # variant 0 worst case 9 bits/byte + 1 byte/block; variant 1 worst case 2 bytes/byte
    cap = (2 * n + 2 * nblocks if variant == LZ_DEFLATE else n + n // 8 + 8 * nblocks) + 64
    return Lz77Stream(variant=variant,
                      out=torch.empty(cap, dtype=torch.uint8, device=ctx.device),
                      block_sizes=torch.empty(nblocks, dtype=torch.int64, device=ctx.device),
                      block_off=torch.empty(nblocks + 1, dtype=torch.int64, device=ctx.device),
                      n=n, block_size=block_size)


def lz77_encode(ctx, data, variant=LZ_DEFLATE, block_size=DEFAULT_BLOCK, stream=None, sync=True):
    """lz77_compress per block with a fresh table
    (algorithms/lz77/lz77.c:264-345 or algorithms/deflate/lz77.c:199-280) + the
    block concatenation of deflate.c:47-63."""
    _check_u8(data)
    n = data.numel()
    st = stream if stream is not None else lz77_alloc(ctx, n, block_size, variant)
    tb = C.c_uint64(0)
    _lib.check(_lib.core().b200_lz77_encode_dev(
        ctx.handle, variant, _ptr(data), n, block_size, _ptr(st.out), st.out.numel(), _ptr(st.block_sizes), _ptr(st.block_off),
        C.byref(tb) if sync else None))
    if sync:
        st.total_bytes = tb.value
    return st


def lz77_encode_debug(ctx, data, variant, block_size):
    """Test hook: (stream, tok) where tok[b, p] is the match finder's candidate for position p
    of block b (0 = literal, else offset | len << 16). Blocks must be <= 65536 bytes."""
    _check_u8(data)
    n = data.numel()
    st = lz77_alloc(ctx, n, block_size, variant)
    nblocks = st.block_sizes.numel()
    tok = torch.zeros(nblocks * 65536 + nblocks * 136, dtype=torch.int32, device=ctx.device)
    tb = C.c_uint64(0)
    _lib.check(_lib.core().b200_lz77_encode_debug_dev(
        ctx.handle, variant, _ptr(data), n, block_size, _ptr(st.out), st.out.numel(), _ptr(st.block_sizes), _ptr(st.block_off),
        C.byref(tb), _ptr(tok)))
    st.total_bytes = tb.value
    st.debug_stats = tok[nblocks * 65536:].view(nblocks, 136)   # 8 phase stamps + 32 warps x (cycles, entries, rounds, coop)
    return st, tok[: nblocks * 65536].view(nblocks, 65536)


def lz77_decode(ctx, st, out=None):
    if out is None:
        out = torch.empty(st.n, dtype=torch.uint8, device=ctx.device)
    _lib.check(_lib.core().b200_lz77_decode_dev(
        ctx.handle, st.variant, _ptr(st.out), _ptr(st.block_off), _ptr(st.block_sizes), st.n, st.block_size, _ptr(out)))
    return out


# ------------------------------------------------------------------ deflate token entropy stage
@dataclass
class DeflateStream:
    """deflate = LZ77 byte tokens (lz) + their Huffman-coded form (words, side)."""
    lz: Lz77Stream             # the byte tokens (scratch on the decode side)
    words: torch.Tensor        # int32 storage of the MSB-first u32 word stream
    side: torch.Tensor         # uint8 side buffer: frequencies[286], codes, trees, token offsets, decode index
    layout: _lib.DflLayout
    n: int
    block_size: int
    total_words: int = -1
    worst_status: int = 0

    def _view(self, off, count, dtype):
        nbytes = count * torch.empty((), dtype=dtype).element_size()
        return self.side[off: off + nbytes].view(dtype)

    def freq(self):
        return self._view(self.layout.off_freq, self.layout.nblocks * 288, torch.int32).view(-1, 288)[:, :286]

    def codes(self):
        return self._view(self.layout.off_codes, self.layout.nblocks * 288, torch.int32).view(-1, 288)[:, :286]

    def lens(self):
        return self.side[self.layout.off_lens: self.layout.off_lens + self.layout.nblocks * 288].view(-1, 288)[:, :286]

    def meta(self):
        return self._view(self.layout.off_meta, self.layout.nblocks * 4, torch.int32).view(-1, 4)

    def block_bits(self):
        return self._view(self.layout.off_block_bits, self.layout.nblocks, torch.int64)

    def block_word(self):
        return self._view(self.layout.off_block_word, self.layout.nblocks + 1, torch.int64)


def dfl_layout(n, block_size):
    L = _lib.DflLayout()
    _lib.check(_lib.core().b200_dfl_layout_for(n, block_size, C.byref(L)))
    return L


def deflate_alloc(ctx, n, block_size, lz=None):
    L = dfl_layout(n, block_size)
    cap = int(_lib.core().b200_dfl_max_words(n, block_size))
    return DeflateStream(lz=lz if lz is not None else lz77_alloc(ctx, n, block_size, LZ_DEFLATE),
                         words=torch.empty(cap, dtype=torch.int32, device=ctx.device),
                         side=torch.empty(L.bytes, dtype=torch.uint8, device=ctx.device),
                         layout=L, n=n, block_size=block_size)


def dfl_encode(ctx, lz, stream=None, sync=True):
    """Entropy stage only: the byte tokens of lz77_encode(variant LZ_DEFLATE) -> frequencies[286],
    code tables and the packed words (the TODO of algorithms/deflate/lz77.c:279)."""
    st = stream if stream is not None else deflate_alloc(ctx, lz.n, lz.block_size, lz)
    tw, ws = C.c_uint64(0), C.c_uint32(0)
    _lib.check(_lib.core().b200_dfl_encode_dev(
        ctx.handle, _ptr(lz.out), lz.out.numel(), _ptr(lz.block_off), _ptr(lz.block_sizes), lz.n, lz.block_size,
        _ptr(st.words), st.words.numel(), _ptr(st.side), st.side.numel(),
        C.byref(tw) if sync else None, C.byref(ws) if sync else None))
    if sync:
        st.total_words, st.worst_status = tw.value, ws.value
    return st


def dfl_decode(ctx, st, tokens_out=None):
    """packed words -> byte tokens (uint8 tensor as large as the encoder's token buffer)"""
    if tokens_out is None:
        tokens_out = torch.zeros_like(st.lz.out)
    _lib.check(_lib.core().b200_dfl_decode_dev(
        ctx.handle, _ptr(st.words), max(st.total_words, 0) if st.total_words >= 0 else st.words.numel(),
        _ptr(st.side), st.side.numel(), st.n, st.block_size, _ptr(tokens_out)))
    return tokens_out


def deflate_compress(ctx, data, block_size=DEFAULT_BLOCK, stream=None, sync=True):
    """lz77_compress per block (algorithms/deflate/lz77.c:199-277) + the entropy stage."""
    _check_u8(data)
    n = data.numel()
    st = stream if stream is not None else deflate_alloc(ctx, n, block_size)
    tw, ws = C.c_uint64(0), C.c_uint32(0)
    _lib.check(_lib.core().b200_deflate_compress_dev(
        ctx.handle, _ptr(data), n, block_size, _ptr(st.lz.out), st.lz.out.numel(), _ptr(st.lz.block_sizes), _ptr(st.lz.block_off),
        _ptr(st.words), st.words.numel(), _ptr(st.side), st.side.numel(),
        C.byref(tw) if sync else None, C.byref(ws) if sync else None))
    if sync:
        st.total_words, st.worst_status = tw.value, ws.value
    return st


def deflate_decompress(ctx, st, out=None, tokens=None):
    """The decoder the reference never wrote (deflate/deflate.c:78-79): words -> tokens -> bytes."""
    if out is None:
        out = torch.empty(st.n, dtype=torch.uint8, device=ctx.device)
    if tokens is None:
        tokens = st.lz.out
    _lib.check(_lib.core().b200_deflate_decompress_dev(
        ctx.handle, _ptr(st.words), max(st.total_words, 0) if st.total_words >= 0 else st.words.numel(),
        _ptr(st.side), st.side.numel(), st.n, st.block_size, _ptr(tokens), _ptr(out)))
    return out


# ------------------------------------------------------------------ FSE
DEFAULT_FSE_SEG = 1024


@dataclass
class FseStream:
    words: torch.Tensor        # int64 storage of the u64 word stream
    side: torch.Tensor         # uint8: histogram, normalised counts, TT, segment index
    layout: _lib.FseLayout
    n: int
    block_size: int
    seg_size: int
    total_words: int = -1

    def _view(self, off, count, dtype):
        nbytes = count * torch.empty((), dtype=dtype).element_size()
        return self.side[off: off + nbytes].view(dtype)

    def freq(self):
        return self._view(self.layout.off_freq, self.layout.nblocks * 256, torch.int32).view(-1, 256)

    def norm(self):
        return self._view(self.layout.off_norm, self.layout.nblocks * 256, torch.int16).view(-1, 256)

    def tt(self):
        return self._view(self.layout.off_tt, self.layout.nblocks * 256, torch.int32).view(-1, 256)

    def seg_bits(self):
        return self._view(self.layout.off_seg_bits, self.layout.nsegs, torch.int32)

    def seg_word(self):
        return self._view(self.layout.off_seg_word, self.layout.nsegs + 1, torch.int64)


def fse_layout(n, block_size, seg_size):
    L = _lib.FseLayout()
    _lib.check(_lib.core().b200_fse_layout_for(n, block_size, seg_size, C.byref(L)))
    return L


def fse_alloc(ctx, n, block_size, seg_size):
    L = fse_layout(n, block_size, seg_size)
    cap = int(_lib.core().b200_fse_max_words(n, seg_size))
    return FseStream(words=torch.empty(cap, dtype=torch.int64, device=ctx.device),
                     side=torch.empty(L.bytes, dtype=torch.uint8, device=ctx.device),
                     layout=L, n=n, block_size=block_size, seg_size=seg_size)


def fse_encode(ctx, data, block_size=DEFAULT_BLOCK, seg_size=DEFAULT_FSE_SEG, stream=None, sync=True):
    """C mirror of compress() (/root/reference/algorithms/fse/src/main.zig:50-68) per segment,
    one table per block."""
    _check_u8(data)
    n = data.numel()
    st = stream if stream is not None else fse_alloc(ctx, n, block_size, seg_size)
    tw = C.c_uint64(0)
    _lib.check(_lib.core().b200_fse_encode_dev(
        ctx.handle, _ptr(data), n, block_size, seg_size, _ptr(st.words), st.words.numel(), _ptr(st.side), st.side.numel(),
        C.byref(tw) if sync else None))
    if sync:
        st.total_words = tw.value
    return st


def fse_normalize(ctx, data, block_size=DEFAULT_BLOCK):
    """buildFrequencyTable + normalizeFrequencyTable (main.zig:88-149) per block."""
    _check_u8(data)
    n = data.numel()
    st = fse_alloc(ctx, n, block_size, DEFAULT_FSE_SEG)
    _lib.check(_lib.core().b200_fse_normalize_dev(ctx.handle, _ptr(data), n, block_size, _ptr(st.side), st.side.numel()))
    return st


def fse_decode(ctx, st, out=None, sync=True):
    if out is None:
        out = torch.empty(st.n, dtype=torch.uint8, device=ctx.device)
    bad = C.c_uint32(0)
    _lib.check(_lib.core().b200_fse_decode_dev(
        ctx.handle, _ptr(st.words), _ptr(st.side), st.side.numel(), st.n, st.block_size, st.seg_size, _ptr(out),
        C.byref(bad) if sync else None))
    if sync and bad.value:
        raise RuntimeError("fse_decode: %d corrupt segment(s)" % bad.value)
    return out
