"""GPU: out-of-bounds WRITE detection without compute-sanitizer (the tool is closed on this GPU pool, see
profiles/r02_sanitizer_refused.txt): every device buffer a codec call writes is carved out of a larger allocation
with 64 KiB guard zones filled with a pattern on both sides; after the call the guards must be untouched. Inputs are
the edge cases of the parity tests (ragged tails, 1-byte blocks, zero bytes behind the block end, skewed long chains,
large blocks in slices). Reads are covered by the bit-exact comparisons of the parity tests (a wrong read changes a
stream), shared-memory indexing by the same plus the static smem budget check at compile time."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GUARD = 65536
PAT = 0xA5


class Guarded:
    """a device tensor of `n` elements between two guard zones"""

    def __init__(self, ctx, n, dtype):
        import torch
        self.es = torch.empty((), dtype=dtype).element_size()
        nbytes = n * self.es
        pad = (-nbytes) % 256
        self.raw = torch.full((GUARD + nbytes + pad + GUARD,), PAT, dtype=torch.uint8, device=ctx.device)
        self.t = self.raw[GUARD: GUARD + nbytes].view(dtype)
        self.nbytes = nbytes

    def check(self, what):
        import torch
        lo = self.raw[:GUARD]
        hi = self.raw[GUARD + self.nbytes + ((-self.nbytes) % 256):]
        assert bool((lo == PAT).all()), "%s: write BELOW the buffer" % what
        assert bool((hi == PAT).all()), "%s: write ABOVE the buffer" % what


def _inputs():
    from compression_algorithms_b200 import corpus
    out = []
    for kind, n in ((0, 3 * 65536 + 123), (0, 1), (0, 65537), (2, 200_000), (3, 65536), (1, 70_001)):
        out.append(("kind%d_n%d" % (kind, n), corpus.generate(n, kind, 5).copy()))
    z = corpus.generate(65536 * 2, 0, 9).copy(); z[-40:] = 0          # zero bytes at the block end: matches run into the padding (U1)
    out.append(("zeros_at_end", z))
    return out


@pytest.mark.parametrize("variant", [1, 0])
@pytest.mark.parametrize("block", [65536, 4096, 0])
def test_lz77_writes_stay_inside(ctx, variant, block):
    import torch
    from compression_algorithms_b200 import _lib, device as dv
    lib = _lib.core()
    for name, data in _inputs():
        n = data.size
        bs = n if (block == 0 or block > n) else block
        nb = (n + bs - 1) // bs
        cap = int(lib.b200_lz77_max_bytes(variant, n, block))
        d_in = Guarded(ctx, n, torch.uint8); d_in.t.copy_(torch.from_numpy(data))
        out, sizes, off, dec = Guarded(ctx, cap, torch.uint8), Guarded(ctx, nb, torch.int64), Guarded(ctx, nb + 1, torch.int64), Guarded(ctx, n + 32, torch.uint8)
        tot = C.c_uint64(0)
        _lib.check(lib.b200_lz77_encode_dev(ctx.handle, variant, d_in.t.data_ptr(), n, block, out.t.data_ptr(), cap, sizes.t.data_ptr(), off.t.data_ptr(), C.byref(tot)))
        _lib.check(lib.b200_lz77_decode_dev(ctx.handle, variant, out.t.data_ptr(), off.t.data_ptr(), sizes.t.data_ptr(), n, block, dec.t.data_ptr()))
        torch.cuda.synchronize()
        for g, w in ((d_in, "input"), (out, "token stream"), (sizes, "block sizes"), (off, "block offsets"), (dec, "decoded bytes")):
            g.check("lz77 variant %d block %d %s: %s" % (variant, block, name, w))
        assert torch.equal(dec.t[:n], d_in.t)


@pytest.mark.parametrize("block", [0, 65536, 4096])
def test_huffman_fse_deflate_writes_stay_inside(ctx, block):
    import torch
    from compression_algorithms_b200 import _lib, device as dv
    lib = _lib.core()
    for name, data in _inputs():
        n = data.size
        if len(np.unique(data[: (n if block == 0 else block)])) < 2 or (block and n % block and len(np.unique(data[n - n % block:])) < 2):
            continue    # the reference cannot encode a one-symbol table scope
        d_in = Guarded(ctx, n, torch.uint8); d_in.t.copy_(torch.from_numpy(data))
        # Huffman
        L = dv.huffman_layout(n, block)
        capw = int(lib.b200_huffman_max_words(n, block))
        words, side, dec = Guarded(ctx, capw, torch.int32), Guarded(ctx, L.bytes, torch.uint8), Guarded(ctx, n, torch.uint8)
        tw, ws = C.c_uint64(0), C.c_uint32(0)
        rc = lib.b200_huffman_encode_dev(ctx.handle, d_in.t.data_ptr(), n, block, words.t.data_ptr(), capw, side.t.data_ptr(), L.bytes, C.byref(tw), C.byref(ws))
        if rc == 0 and ws.value == 0:
            _lib.check(lib.b200_huffman_decode_dev(ctx.handle, words.t.data_ptr(), tw.value, side.t.data_ptr(), L.bytes, n, block, dec.t.data_ptr()))
            torch.cuda.synchronize()
            assert torch.equal(dec.t, d_in.t)
        torch.cuda.synchronize()
        for g, w in ((words, "words"), (side, "side buffer"), (dec, "decoded bytes"), (d_in, "input")):
            g.check("huffman block %d %s: %s" % (block, name, w))
        # FSE (table scope = block, 1 KiB segments)
        fb = block if block else 0
        FL = _lib.FseLayout()
        _lib.check(lib.b200_fse_layout_for(n, fb, 1024, C.byref(FL)))
        capf = int(lib.b200_fse_max_words(n, 1024))
        fw, fside, fdec = Guarded(ctx, capf, torch.int64), Guarded(ctx, FL.bytes, torch.uint8), Guarded(ctx, n, torch.uint8)
        bad = C.c_uint32(0)
        _lib.check(lib.b200_fse_encode_dev(ctx.handle, d_in.t.data_ptr(), n, fb, 1024, fw.t.data_ptr(), capf, fside.t.data_ptr(), FL.bytes, C.byref(tw)))
        _lib.check(lib.b200_fse_decode_dev(ctx.handle, fw.t.data_ptr(), fside.t.data_ptr(), FL.bytes, n, fb, 1024, fdec.t.data_ptr(), C.byref(bad)))
        torch.cuda.synchronize()
        assert bad.value == 0 and torch.equal(fdec.t, d_in.t)
        for g, w in ((fw, "words"), (fside, "side buffer"), (fdec, "decoded bytes")):
            g.check("fse block %d %s: %s" % (block, name, w))
        # deflate with the entropy stage
        DL = dv.dfl_layout(n, block)
        tok_cap = int(lib.b200_lz77_max_bytes(1, n, block))
        nb = int(DL.nblocks)
        capd = int(lib.b200_dfl_max_words(n, block))
        tok, tsz, toff = Guarded(ctx, tok_cap, torch.uint8), Guarded(ctx, nb, torch.int64), Guarded(ctx, nb + 1, torch.int64)
        dw, dside, ddec, tok2 = Guarded(ctx, capd, torch.int32), Guarded(ctx, DL.bytes, torch.uint8), Guarded(ctx, n + 32, torch.uint8), Guarded(ctx, tok_cap, torch.uint8)
        _lib.check(lib.b200_deflate_compress_dev(ctx.handle, d_in.t.data_ptr(), n, block, tok.t.data_ptr(), tok_cap, tsz.t.data_ptr(), toff.t.data_ptr(),
                                                 dw.t.data_ptr(), capd, dside.t.data_ptr(), DL.bytes, C.byref(tw), C.byref(ws)))
        if ws.value == 0:
            _lib.check(lib.b200_deflate_decompress_dev(ctx.handle, dw.t.data_ptr(), tw.value, dside.t.data_ptr(), DL.bytes, n, block, tok2.t.data_ptr(), ddec.t.data_ptr()))
            torch.cuda.synchronize()
            assert torch.equal(ddec.t[:n], d_in.t)
        torch.cuda.synchronize()
        for g, w in ((tok, "tokens"), (tsz, "token sizes"), (toff, "token offsets"), (dw, "words"), (dside, "side buffer"), (ddec, "decoded bytes"), (tok2, "decoded tokens")):
            g.check("deflate block %d %s: %s" % (block, name, w))
