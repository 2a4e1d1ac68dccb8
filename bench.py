#!/usr/bin/env python
"""Headline benchmark (BASELINE.json): deflate compress + decompress of an
enwik9-shaped 1 GB buffer cut into 64 KiB blocks, sharded by block over the GPUs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A step = one pass of the hot path over one batch: the repo's deflate LZ77 match
finder + greedy parse + token emission + block compaction of the whole buffer
(algorithms/deflate `compress`), followed by the decode of that stream.
  value     : device-resident throughput, uncompressed GB / (compress + decompress) s,
              whole job over all N GPUs (strong scaling, BASELINE.json configs[3] as written: ONE 1 GB
              buffer, rank r owns the blocks of sharding.byte_range(); the ranks all-gather their
              shard sizes). "weak_scaling" (1 GB per rank) is an extra key at N > 1.
  e2e       : the same through the host-buffer C-ABI calls a reference driver would
              make (b200_lz77_compress_host / _decompress_host), pinned host buffers,
              H2D/D2H copies inside the timed region.
  roofline  : the dominant kernel (LZ77 parse) against measured HBM copy bandwidth.
  cpu_baseline : the reference's own C (oracle/_ref, built from /root/reference) timed on
              this box's host cores on a bounded sample, blocks spread over all cores (and one core,
              the reference as written); its tokens are compared with the GPU stream of the same blocks.
  detail    : configs[0..2] (Huffman, FSE, LZ77, entropy-coded deflate at 100 MB) with their own
              roofline, cpu_baseline and e2e objects.
Rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "compress/decompress GB/s (deflate), enwik9-shaped 1 GB"
UNIT = "GB/s"
BLOCK = 65536
N_BYTES = 1_000_000_000
CPU_SAMPLE = 256 * 1024 * 1024


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons while the timed region runs."""

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        """Started before the warm-up steps (nvidia-smi takes a few hundred ms to print its first row); only rows
        that arrive between mark_begin() and stop() are used."""
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        self.t_begin = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def wait_first(self, timeout=5.0):
        t0 = time.perf_counter()
        while self.proc is not None and not self.rows and time.perf_counter() - t0 < timeout:
            time.sleep(0.02)

    def mark_begin(self):
        self.t_begin = time.perf_counter()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        t_end = time.perf_counter()
        time.sleep(0.12)
        self.proc.terminate()
        t0 = self.t_begin if self.t_begin is not None else 0.0
        rows = [r for t, r in self.rows if t0 <= t <= t_end + 0.06]
        window = "timed region"
        if not rows:   # a timed region shorter than the sampling period: fall back to every row taken under load
            rows, window = [r for _, r in self.rows], "warm-up + timed region"
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 4), ("hw_thermal_slowdown", 5), ("sw_thermal_slowdown", 6), ("sw_power_cap", 7)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "window": window, "reasons": sorted(reasons)}


def _cores():
    try:
        return len(os.sched_getaffinity(0)) or 1
    except AttributeError:
        return os.cpu_count() or 1


def cpu_baseline(data_np, sample_bytes, threads=0, gpu_stream=None, one_core_bytes=0):
    """The reference's deflate lz77_compress per 64 KiB block on a fresh table (oracle/_ref when
    present, else the oracle port), blocks spread over all host cores; decode = the oracle
    port's byte-token decoder (the reference ships none, deflate.c:78-79).
    gpu_stream = (token bytes, block offsets, block sizes) of the GPU for the same buffer: the reference's
    tokens of the sampled blocks are compared with it byte for byte (parity at the headline scale).
    one_core_bytes > 0 additionally times the reference as written (one thread) on a shorter prefix."""
    from oracle import bindings as ob
    n = min(sample_bytes, data_np.size)
    sample = np.ascontiguousarray(data_np[:n])
    cores = _cores()
    if threads <= 0:
        threads = cores     # explicit: torchrun exports OMP_NUM_THREADS=1, which would leave the harness on one thread
    kind = "reference" if ob.have_ref() else "port"

    def compress(buf, thr):
        if kind == "reference":
            return ob.ref_deflate_lz77_compress_blocks(buf, BLOCK, persistent=False, threads=thr)
        out, sizes = ob.port_lz77_compress_blocks(buf, BLOCK, 1, thr)
        return [out[b, : int(sizes[b])] for b in range(len(sizes))], sizes

    t0 = time.perf_counter()
    blocks, sizes = compress(sample, threads)
    t1 = time.perf_counter()
    stream = np.concatenate(blocks)
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
    t2 = time.perf_counter()
    dec, bad = ob.port_lz77_decompress_blocks(stream, off, BLOCK, n, 1, threads)
    t3 = time.perf_counter()
    ok = bad == 0 and np.array_equal(dec, sample)
    tc, td = t1 - t0, t3 - t2
    res = {"value": n / 1e9 / (tc + td), "unit": UNIT, "cores": threads, "kind": kind,
           "sample": "first %d MiB of the workload, %d blocks of 64 KiB, one %s lz77_compress call per block on a fresh table "
                     "spread over %d threads; decode by the oracle port (the reference has no deflate decoder)"
                     % (n >> 20, len(sizes), "reference" if kind == "reference" else "oracle-port", threads),
           "compress_gbps": n / 1e9 / tc, "decompress_gbps": n / 1e9 / td, "roundtrip_ok": bool(ok),
           "token_bytes": int(stream.size)}
    if gpu_stream is not None:
        g_tok, g_off, g_sizes = gpu_stream
        nb = len(sizes)
        same_sizes = bool(np.array_equal(np.asarray(g_sizes[:nb], dtype=np.uint64), np.asarray(sizes, dtype=np.uint64)))
        same_off = bool(np.array_equal(np.asarray(g_off[: nb + 1], dtype=np.uint64), off))
        same_bytes = bool(same_off and np.array_equal(g_tok[: stream.size], stream))
        res["parity_blocks_checked"] = int(nb)
        res["parity_ok"] = bool(same_sizes and same_off and same_bytes)
        res["parity"] = "GPU token stream, block sizes and offsets of the sampled blocks == %s output, byte for byte" % (
            "compiled reference (oracle/_ref)" if kind == "reference" else "oracle port")
    if one_core_bytes > 0:
        m = min(one_core_bytes, n)
        s1 = np.ascontiguousarray(sample[:m])
        t0 = time.perf_counter()
        b1, z1 = compress(s1, 1)
        t1 = time.perf_counter()
        st1 = np.concatenate(b1)
        o1 = np.concatenate([[0], np.cumsum(z1)]).astype(np.uint64)
        t2 = time.perf_counter()
        ob.port_lz77_decompress_blocks(st1, o1, BLOCK, m, 1, 1)
        t3 = time.perf_counter()
        res["one_core"] = {"value": m / 1e9 / ((t1 - t0) + (t3 - t2)), "unit": UNIT, "cores": 1,
                           "compress_gbps": m / 1e9 / (t1 - t0), "decompress_gbps": m / 1e9 / (t3 - t2),
                           "sample": "first %d MiB, the reference as written (single thread)" % (m >> 20)}
    return res


def run_reference(args):
    """--impl reference: the reference's CPU implementation on this box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from compression_algorithms_b200 import corpus
    data = corpus.generate(CPU_SAMPLE, corpus.ENWIK, corpus.DEFAULT_SEED)
    for _ in range(args.warmup):
        cpu_baseline(data, CPU_SAMPLE // 8)
    t0 = time.perf_counter()
    last = None
    vals = []
    for _ in range(args.steps):
        last = cpu_baseline(data, CPU_SAMPLE)
        vals.append(last["value"])
    dt = time.perf_counter() - t0
    value = float(np.mean(vals))
    cb = dict(last); cb["value"] = value
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "deflate (algorithms/deflate lz77_compress per 64 KiB block, fresh table) compress+decompress; "
                                   "each step = a %d MiB sample of the enwik9-shaped 1 GB buffer on the host CPU" % (CPU_SAMPLE >> 20),
                       "block_size": BLOCK, "bytes_per_step": CPU_SAMPLE},
            "cpu_baseline": cb,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def _timed(torch, fn, reps=3):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        r = fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps / 1e3, r


def _wall(torch, fn, reps=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def _roof(n, c, te, td, peak):
    """Encode / decode against the HBM roofline: algorithmic bytes N + C each way."""
    enc, dec = (n + c) / 1e9 / te, (n + c) / 1e9 / td
    return {"bound": "hbm", "unit": "GB/s", "peak": peak, "algorithmic_bytes": int(n + c),
            "encode": {"achieved": enc, "frac": enc / peak, "ms": te * 1e3}, "decode": {"achieved": dec, "frac": dec / peak, "ms": td * 1e3}}


def detail_codecs(ctx, dv, torch, data100, h_data100):
    """BASELINE.json configs[0..2] at 100 MB: device-resident GB/s, roofline of the codec's kernels, the reference on
    this box's CPU (one core = as written, and all cores where the reference call can be split), and the same
    codec through the host-buffer C-ABI (pinned buffers, copies inside the timed region)."""
    import ctypes as C
    from compression_algorithms_b200 import _lib
    from oracle import bindings as ob
    lib = _lib.core()
    out = {}
    n = data100.numel()
    peak, _ = hbm_peak()
    cores = _cores()
    have_ref = ob.have_ref()
    host = h_data100.numpy()
    cpu_n = min(n, 32 << 20)           # bounded CPU samples
    cpu_s = np.ascontiguousarray(host[:cpu_n])

    # ---- configs[0]: Huffman (whole buffer = the reference call; 64 KiB table scopes = the block-parallel mode)
    for name, block in (("huffman_whole_buffer", 0), ("huffman_64k_blocks", BLOCK)):
        st = dv.huffman_alloc(ctx, n, block)
        te, st = _timed(torch, lambda: dv.huffman_encode(ctx, data100, block, stream=st, sync=False))
        st = dv.huffman_encode(ctx, data100, block, stream=st)
        dec = torch.empty_like(data100)
        td, _ = _timed(torch, lambda: dv.huffman_decode(ctx, st, out=dec))
        cbytes = st.total_words * 4
        r = {"compress_gbps": n / 1e9 / te, "decompress_gbps": n / 1e9 / td, "ratio": n / float(cbytes),
             "roundtrip_ok": bool(torch.equal(dec, data100)), "roofline": _roof(n, cbytes, te, td, peak)}
        # e2e through b200_huffman_compress_host / _decompress_host
        L = dv.huffman_layout(n, block)
        cap = int(lib.b200_huffman_max_words(n, block))
        h_words = torch.empty(cap, dtype=torch.int32).pin_memory()
        h_side = torch.empty(L.bytes, dtype=torch.uint8).pin_memory()
        h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
        tw, ws = C.c_uint64(0), C.c_uint32(0)

        def hc():
            _lib.check(lib.b200_huffman_compress_host(ctx.handle, h_data100.data_ptr(), n, block, h_words.data_ptr(), cap, h_side.data_ptr(),
                                                      L.bytes, C.byref(tw), C.byref(ws)))

        def hd():
            _lib.check(lib.b200_huffman_decompress_host(ctx.handle, h_words.data_ptr(), tw.value, h_side.data_ptr(), L.bytes, n, block, h_out.data_ptr()))

        tce, tde = _wall(torch, hc), _wall(torch, hd)
        r["e2e"] = {"value": n / 1e9 / (tce + tde), "unit": UNIT, "compress_gbps": n / 1e9 / tce, "decompress_gbps": n / 1e9 / tde,
                    "h2d_bytes_per_step": int(n + tw.value * 4 + L.bytes), "d2h_bytes_per_step": int(tw.value * 4 + L.bytes + n),
                    "roundtrip_ok": bool(np.array_equal(h_out.numpy(), host)), "api": "b200_huffman_compress_host + b200_huffman_decompress_host"}
        if have_ref:
            if block == 0:
                t = ob.ref_huffman_time(cpu_s)     # huffman_compress + huffman_decompress as main.c:50-80 calls them
                r["cpu_baseline"] = {"kind": "reference", "cores": 1, "unit": UNIT, "value": cpu_n / 1e9 / (t["t_comp"] + t["t_decomp"]),
                                     "compress_gbps": cpu_n / 1e9 / t["t_comp"], "decompress_gbps": cpu_n / 1e9 / t["t_decomp"],
                                     "sample": "first %d MiB, one huffman_compress + huffman_decompress call (the reference as written; one tree over the "
                                               "whole buffer cannot be split over cores)" % (cpu_n >> 20)}
            else:
                t0 = time.perf_counter(); ob.ref_huffman_compress_blocks(cpu_s, block, threads=1); t1 = time.perf_counter()
                ob.ref_huffman_compress_blocks(cpu_s, block, threads=cores); t2 = time.perf_counter()
                r["cpu_baseline"] = {"kind": "reference", "unit": UNIT, "cores": cores, "compress_gbps": cpu_n / 1e9 / (t2 - t1),
                                     "one_core": {"cores": 1, "compress_gbps": cpu_n / 1e9 / (t1 - t0)},
                                     "sample": "first %d MiB, one huffman_compress call per 64 KiB block (compress only: the harness keeps no per-block trees "
                                               "for the reference decoder)" % (cpu_n >> 20)}
        out[name] = r

    # ---- configs[1]: FSE (parity unpinned beyond histogram + normalisation: the CPU side is the oracle port)
    st = dv.fse_alloc(ctx, n, BLOCK, dv.DEFAULT_FSE_SEG)
    te, st = _timed(torch, lambda: dv.fse_encode(ctx, data100, BLOCK, dv.DEFAULT_FSE_SEG, stream=st, sync=False))
    st = dv.fse_encode(ctx, data100, BLOCK, dv.DEFAULT_FSE_SEG, stream=st)
    dec = torch.empty_like(data100)
    td, _ = _timed(torch, lambda: dv.fse_decode(ctx, st, out=dec, sync=False))
    cbytes = st.total_words * 8
    r = {"compress_gbps": n / 1e9 / te, "decompress_gbps": n / 1e9 / td, "ratio": n / float(cbytes), "roundtrip_ok": bool(torch.equal(dec, data100)),
         "roofline": _roof(n, cbytes, te, td, peak)}
    capw = int(lib.b200_fse_container_max_words(n, BLOCK, dv.DEFAULT_FSE_SEG))
    h_cont = torch.empty(capw, dtype=torch.int64).pin_memory()
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    tw, hn = C.c_uint64(0), C.c_uint64(0)

    def fc():
        _lib.check(lib.b200_fse_compress_host(ctx.handle, h_data100.data_ptr(), n, BLOCK, dv.DEFAULT_FSE_SEG, h_cont.data_ptr(), capw, C.byref(tw)))

    def fd():
        _lib.check(lib.b200_fse_decompress_host(ctx.handle, h_cont.data_ptr(), tw.value, h_out.data_ptr(), n, C.byref(hn)))

    tce, tde = _wall(torch, fc), _wall(torch, fd)
    r["e2e"] = {"value": n / 1e9 / (tce + tde), "unit": UNIT, "compress_gbps": n / 1e9 / tce, "decompress_gbps": n / 1e9 / tde,
                "h2d_bytes_per_step": int(n + tw.value * 8), "d2h_bytes_per_step": int(tw.value * 8 + n),
                "roundtrip_ok": bool(np.array_equal(h_out.numpy(), host)), "api": "b200_fse_compress_host + b200_fse_decompress_host (self-describing container)"}
    fs = np.ascontiguousarray(host[: 4 << 20])
    t0 = time.perf_counter(); w, norm, tb, _sz = ob.port_fse_compress(fs); t1 = time.perf_counter()
    ob.port_fse_decompress(w, tb, fs.size, norm); t2 = time.perf_counter()
    r["cpu_baseline"] = {"kind": "port", "cores": 1, "unit": UNIT, "value": fs.size / 1e9 / (t2 - t0), "compress_gbps": fs.size / 1e9 / (t1 - t0),
                         "decompress_gbps": fs.size / 1e9 / (t2 - t1),
                         "sample": "first 4 MiB as ONE single-state stream through the oracle port (the reference's FSE does not compile: main.zig:47)"}
    out["fse_64k_blocks_1k_segments"] = r

    # ---- configs[2] (and the L rows): standalone LZ77, 64 KiB blocks
    st = dv.lz77_alloc(ctx, n, BLOCK, dv.LZ_STANDALONE)
    te, st = _timed(torch, lambda: dv.lz77_encode(ctx, data100, dv.LZ_STANDALONE, BLOCK, stream=st, sync=False), reps=2)
    st = dv.lz77_encode(ctx, data100, dv.LZ_STANDALONE, BLOCK, stream=st)
    dec = torch.empty_like(data100)
    td, _ = _timed(torch, lambda: dv.lz77_decode(ctx, st, out=dec), reps=2)
    r = {"compress_gbps": n / 1e9 / te, "decompress_gbps": n / 1e9 / td, "ratio": n / float(st.total_bytes), "roundtrip_ok": bool(torch.equal(dec, data100)),
         "roofline": _roof(n, st.total_bytes, te, td, peak)}
    if have_ref:
        ls = np.ascontiguousarray(host[: 16 << 20])
        t0 = time.perf_counter(); ob.ref_lz77_compress_blocks(ls, BLOCK, threads=cores); t1 = time.perf_counter()
        l1 = np.ascontiguousarray(host[: 2 << 20])
        t2 = time.perf_counter(); ob.ref_lz77_compress_blocks(l1, BLOCK, threads=1); t3 = time.perf_counter()
        r["cpu_baseline"] = {"kind": "reference", "unit": UNIT, "cores": cores, "compress_gbps": ls.size / 1e9 / (t1 - t0),
                             "one_core": {"cores": 1, "compress_gbps": l1.size / 1e9 / (t3 - t2)},
                             "sample": "first 16 MiB (all cores) / 2 MiB (one core), one algorithms/lz77 lz77_compress call per 64 KiB block, "
                                       "each allocating and clearing its own 24 MiB table as the reference does"}
    out["lz77_standalone_64k_blocks"] = r

    # ---- configs[2]: deflate with the entropy stage (the reference's TODO, deflate/lz77.c:279): LZ77 tokens -> Huffman-coded words
    ds = dv.deflate_alloc(ctx, n, BLOCK)
    te, ds = _timed(torch, lambda: dv.deflate_compress(ctx, data100, BLOCK, stream=ds, sync=False), reps=2)
    ds = dv.deflate_compress(ctx, data100, BLOCK, stream=ds)
    tes, _ = _timed(torch, lambda: dv.dfl_encode(ctx, ds.lz, stream=ds, sync=False))
    dec = torch.empty_like(data100)
    td, _ = _timed(torch, lambda: dv.deflate_decompress(ctx, ds, out=dec), reps=2)
    tok = torch.empty_like(ds.lz.out)
    tds, _ = _timed(torch, lambda: dv.dfl_decode(ctx, ds, tok))
    cbytes = ds.total_words * 4
    r = {"compress_gbps": n / 1e9 / te, "decompress_gbps": n / 1e9 / td, "ratio": n / float(cbytes),
         "entropy_stage_encode_ms": tes * 1e3, "entropy_stage_decode_ms": tds * 1e3,
         "token_bytes": int(ds.lz.block_off[-1].item()), "stream_bytes": int(cbytes),
         "roundtrip_ok": bool(torch.equal(dec, data100)), "roofline": _roof(n, cbytes, te, td, peak)}
    L = ds.layout
    if L is not None:
        capw = int(lib.b200_dfl_max_words(n, BLOCK))
        h_words = torch.empty(capw, dtype=torch.int32).pin_memory()
        h_side = torch.empty(L.bytes, dtype=torch.uint8).pin_memory()
        h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
        tw, ws = C.c_uint64(0), C.c_uint32(0)

        def dc():
            _lib.check(lib.b200_deflate_compress_host(ctx.handle, h_data100.data_ptr(), n, BLOCK, h_words.data_ptr(), capw, h_side.data_ptr(), L.bytes,
                                                      C.byref(tw), C.byref(ws)))

        def dd():
            _lib.check(lib.b200_deflate_decompress_host(ctx.handle, h_words.data_ptr(), tw.value, h_side.data_ptr(), L.bytes, n, BLOCK, h_out.data_ptr()))

        tce, tde = _wall(torch, dc, reps=2), _wall(torch, dd, reps=2)
        r["e2e"] = {"value": n / 1e9 / (tce + tde), "unit": UNIT, "compress_gbps": n / 1e9 / tce, "decompress_gbps": n / 1e9 / tde,
                    "h2d_bytes_per_step": int(n + tw.value * 4 + L.bytes), "d2h_bytes_per_step": int(tw.value * 4 + L.bytes + n),
                    "roundtrip_ok": bool(np.array_equal(h_out.numpy(), host)), "api": "b200_deflate_compress_host + b200_deflate_decompress_host"}
    out["deflate_entropy_coded_64k_blocks"] = r
    return out


def block_sweep(ctx, dv, torch, corpus, nbytes=32 << 20):
    """BASELINE.json configs[4], bounded: block size 64 KiB .. 4 MiB x {low-entropy, enwik-shaped, near-random} x
    {deflate, Huffman, FSE} on a 32 MiB buffer: ratio and device-resident GB/s, round trip checked. A 4 MiB block is one
    table scope for the entropy coders and one sequential table walk for the match finder (8 blocks for 148 SMs)."""
    rows = []
    for kind, kname in ((corpus.ACGT, "low-entropy (acgt)"), (corpus.ENWIK, "enwik-shaped"), (corpus.RANDOM, "near-random")):
        d = torch.from_numpy(corpus.generate(nbytes, kind, 7)).to(ctx.device)
        dec = torch.empty_like(d)
        for bs in (1 << 16, 1 << 18, 1 << 20, 1 << 22):
            row = {"input": kname, "block": bs}
            st = dv.lz77_alloc(ctx, nbytes, bs, dv.LZ_DEFLATE)
            tc, _ = _timed(torch, lambda: dv.lz77_encode(ctx, d, dv.LZ_DEFLATE, bs, stream=st, sync=False), reps=1)
            st = dv.lz77_encode(ctx, d, dv.LZ_DEFLATE, bs, stream=st)
            td, _ = _timed(torch, lambda: dv.lz77_decode(ctx, st, out=dec), reps=1)
            row["deflate"] = {"ratio": nbytes / float(st.total_bytes), "compress_gbps": nbytes / 1e9 / tc, "decompress_gbps": nbytes / 1e9 / td,
                              "ok": bool(torch.equal(dec, d))}
            hs = dv.huffman_alloc(ctx, nbytes, bs)
            tc, _ = _timed(torch, lambda: dv.huffman_encode(ctx, d, bs, stream=hs, sync=False), reps=2)
            hs = dv.huffman_encode(ctx, d, bs, stream=hs)
            td, _ = _timed(torch, lambda: dv.huffman_decode(ctx, hs, out=dec), reps=2)
            row["huffman"] = {"ratio": nbytes / (hs.total_words * 4.0), "compress_gbps": nbytes / 1e9 / tc, "decompress_gbps": nbytes / 1e9 / td,
                              "ok": bool(torch.equal(dec, d))}
            fs = dv.fse_alloc(ctx, nbytes, bs, dv.DEFAULT_FSE_SEG)
            tc, _ = _timed(torch, lambda: dv.fse_encode(ctx, d, bs, dv.DEFAULT_FSE_SEG, stream=fs, sync=False), reps=2)
            fs = dv.fse_encode(ctx, d, bs, dv.DEFAULT_FSE_SEG, stream=fs)
            td, _ = _timed(torch, lambda: dv.fse_decode(ctx, fs, out=dec, sync=False), reps=2)
            row["fse"] = {"ratio": nbytes / (fs.total_words * 8.0), "compress_gbps": nbytes / 1e9 / tc, "decompress_gbps": nbytes / 1e9 / td,
                          "ok": bool(torch.equal(dec, d))}
            rows.append(row)
    return {"bytes": nbytes, "rows": rows,
            "note": "device resident, CUDA events after a warm-up; deflate = algorithms/deflate lz77_compress per block (blocks above 64 KiB are simulated "
                    "in 64 KiB slices by one CTA each, so few large blocks leave most SMs idle)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--bytes", type=int, default=N_BYTES, help="bytes of the WHOLE job (default: the 1 GB headline config), sharded by block over the ranks")
    ap.add_argument("--no-detail", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-weak", action="store_true", help="skip the extra weak-scaling figure (N > 1)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.warmup < 3:
        args.warmup = 3

    import ctypes as C
    import torch
    import torch.distributed as dist
    from compression_algorithms_b200 import _lib, corpus, device as dv, sharding

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    torch.cuda.set_device(local_rank)
    numa = None
    if world > 1 and os.environ.get("B200_BENCH_NO_AFFINITY") != "1":
        # one process per GPU: run (and first-touch the pinned host buffers) on the CPUs next to this rank's GPU
        try:
            import pynvml
            pynvml.nvmlInit()
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_rank))
            numa = sorted(os.sched_getaffinity(0))
            numa = "cpus %d-%d (%d)" % (numa[0], numa[-1], len(numa))
        except Exception as e:   # affinity is an optimisation, never a requirement
            numa = "unchanged (%s)" % type(e).__name__
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = dv.Context(local_rank)
    lib = _lib.core()
    n_global = args.bytes
    peak, peak_src = hbm_peak()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_config(start, n, steps, warmup, sampler=None, want_e2e=True):
        """One timed configuration on this rank's shard [start, start + n) of the global buffer."""
        nblocks = (n + BLOCK - 1) // BLOCK
        h_in = torch.empty(max(n, 1), dtype=torch.uint8).pin_memory()
        corpus.generate_range(start, n, corpus.ENWIK, corpus.DEFAULT_SEED, out=h_in.numpy())
        d_in = h_in[:n].to(ctx.device, non_blocking=False)
        st = dv.lz77_alloc(ctx, n, BLOCK, dv.LZ_DEFLATE)
        d_dec = torch.empty(n, dtype=torch.uint8, device=ctx.device)
        size1 = torch.zeros(1, dtype=torch.int64, device=ctx.device)

        def step():
            if n:
                dv.lz77_encode(ctx, d_in, dv.LZ_DEFLATE, BLOCK, stream=st, sync=False)
            if world > 1:
                # the one exchange of the path: all-gather the shard sizes -> global offset of every shard
                sharding.exchange_sizes(st.block_off[-1:] if n else size1, ctx.device)
            if n:
                dv.lz77_decode(ctx, st, out=d_dec)

        for _ in range(warmup):
            step()
        barrier()
        launches0 = ctx.launches
        ctx.set_timing(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        if sampler is not None:
            sampler.mark_begin()
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        barrier()
        clocks = sampler.stop() if sampler is not None else None
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=ctx.device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        res = {"ms_per_step": ms.item() / steps, "launches": int(ctx.launches - launches0), "timings": ctx.timings(), "clocks": clocks,
               "n": n, "nblocks": nblocks}
        ctx.set_timing(False)
        # correctness of what was timed + sizes
        if n:
            st2 = dv.lz77_encode(ctx, d_in, dv.LZ_DEFLATE, BLOCK, stream=st, sync=True)
            dv.lz77_decode(ctx, st2, out=d_dec)
            res["ok"] = bool(torch.equal(d_dec, d_in))
            res["T"] = int(st2.total_bytes)
        else:
            res["ok"], res["T"] = True, 0
        res["st"], res["h_in"], res["d_in"] = st, h_in, d_in
        if want_e2e:
            # ---- e2e: the host-buffer C-ABI path with the copies in the timed region
            cap = int(lib.b200_lz77_max_bytes(dv.LZ_DEFLATE, n, BLOCK))
            h_out = torch.empty(cap, dtype=torch.uint8).pin_memory()
            h_sizes = torch.empty(max(nblocks, 1), dtype=torch.int64).pin_memory()
            h_off = torch.empty(nblocks + 1, dtype=torch.int64).pin_memory()
            h_dec = torch.empty(max(n, 1), dtype=torch.uint8).pin_memory()
            tot = C.c_uint64(0)

            def e2e_step():
                _lib.check(lib.b200_lz77_compress_host(ctx.handle, dv.LZ_DEFLATE, h_in.data_ptr(), n, BLOCK, h_out.data_ptr(), cap,
                                                       h_sizes.data_ptr(), h_off.data_ptr(), C.byref(tot)))
                _lib.check(lib.b200_lz77_decompress_host(ctx.handle, dv.LZ_DEFLATE, h_out.data_ptr(), tot.value, h_off.data_ptr(),
                                                         h_sizes.data_ptr(), n, BLOCK, h_dec.data_ptr()))

            for _ in range(2):
                e2e_step()
            barrier()
            e2e_steps = max(10, steps)
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                e2e_step()
            torch.cuda.synchronize()
            t_e2e = torch.tensor([(time.perf_counter() - t0) / e2e_steps], dtype=torch.float64, device=ctx.device)
            if world > 1:
                dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
            idx_bytes = (2 * nblocks + 1) * 8
            res["e2e"] = {"seconds": t_e2e.item(), "steps": e2e_steps, "ok": bool(np.array_equal(h_dec.numpy()[:n], h_in.numpy()[:n])),
                          "h2d": int(n + tot.value + idx_bytes), "d2h": int(tot.value + idx_bytes + n)}
            # the same from PAGEABLE (malloc) buffers, what an unmodified reference driver passes
            p_in = np.array(h_in.numpy()[:n], copy=True)
            p_out = np.empty(cap, dtype=np.uint8); p_dec = np.empty(max(n, 1), dtype=np.uint8)
            p_sizes = np.empty(max(nblocks, 1), dtype=np.uint64); p_off = np.empty(nblocks + 1, dtype=np.uint64)

            def e2e_pageable():
                _lib.check(lib.b200_lz77_compress_host(ctx.handle, dv.LZ_DEFLATE, p_in.ctypes.data, n, BLOCK, p_out.ctypes.data, cap,
                                                       p_sizes.ctypes.data, p_off.ctypes.data, C.byref(tot)))
                _lib.check(lib.b200_lz77_decompress_host(ctx.handle, dv.LZ_DEFLATE, p_out.ctypes.data, tot.value, p_off.ctypes.data,
                                                         p_sizes.ctypes.data, n, BLOCK, p_dec.ctypes.data))

            e2e_pageable()
            barrier()
            t0 = time.perf_counter()
            for _ in range(3):
                e2e_pageable()
            torch.cuda.synchronize()
            t_pg = torch.tensor([(time.perf_counter() - t0) / 3], dtype=torch.float64, device=ctx.device)
            if world > 1:
                dist.all_reduce(t_pg, op=dist.ReduceOp.MAX)
            res["e2e"]["pageable_seconds"] = t_pg.item()
            res["e2e"]["pageable_ok"] = bool(np.array_equal(p_dec[:n], p_in))
        return res

    # ---- headline: configs[3] as written -- ONE buffer of n_global bytes, sharded by block over the ranks (strong scaling)
    start, end = sharding.byte_range(n_global, BLOCK, rank, world)
    sampler = ClockSampler(local_rank)
    sampler.start()
    sampler.wait_first()
    R = run_config(start, end - start, args.steps, args.warmup, sampler=sampler)
    ms_per_step = R["ms_per_step"]
    value = n_global / 1e9 / (ms_per_step / 1e3)
    tot_T = torch.tensor([R["T"]], dtype=torch.int64, device=ctx.device)
    all_ok = torch.tensor([1 if (R["ok"] and R["e2e"]["ok"] and R["e2e"]["pageable_ok"]) else 0], dtype=torch.int64, device=ctx.device)
    h2d = torch.tensor([R["e2e"]["h2d"], R["e2e"]["d2h"]], dtype=torch.int64, device=ctx.device)
    if world > 1:
        dist.all_reduce(tot_T); dist.all_reduce(all_ok, op=dist.ReduceOp.MIN); dist.all_reduce(h2d)
    n, T = R["n"], R["T"]
    parse_ms = [m for k, m in R["timings"] if k == 0]
    dec_ms = [m for k, m in R["timings"] if k == 1]
    parse_avg = float(np.mean(parse_ms)) if parse_ms else float("nan")
    achieved = (n + T) / 1e9 / (parse_avg / 1e3)
    traffic = None
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tp):
        try:
            with open(tp) as f:
                tj = json.load(f)
                per_byte = (tj.get("lz77_v4_kernel") or tj.get("lz77_v2_kernel<1>", {})).get("dram_bytes_per_input_byte")
            traffic = int(per_byte * n) if per_byte else None
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": "deflate-variant match finder of the encode call: lz77_v4_kernel for input of at least 1.5 bits per byte (a byte-entropy sample decides), lz77_v2_kernel for the blocks it hands back and for few-symbol input; + greedy parse + token emission; rank 0's shard",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_source": "static: dram__bytes per input byte of one ncu --set full capture of this kernel (profiles/roofline_traffic.json: lz77_v4_kernel over 296 blocks, lz77_v2_kernel over 100 MB), scaled to this launch's input; not re-measured per run",
                "peak_source": peak_src, "algorithmic_bytes_per_launch": n + T, "avg_launch_ms": parse_avg,
                "launches_timed": len(parse_ms), "share_of_step": parse_avg / ms_per_step,
                "decode_kernel_avg_ms": float(np.mean(dec_ms)) if dec_ms else None,
                "decode_kernel": {"name": "lz77_decode_units_kernel", "achieved": (n + T) / 1e9 / (float(np.mean(dec_ms)) / 1e3) if dec_ms else None,
                                  "frac": (n + T) / 1e9 / (float(np.mean(dec_ms)) / 1e3) / peak if dec_ms else None}}
    e2e = {"value": n_global / 1e9 / R["e2e"]["seconds"], "unit": UNIT,
           "h2d_bytes_per_step": int(h2d[0].item()), "d2h_bytes_per_step": int(h2d[1].item()),
           "steps": R["e2e"]["steps"], "roundtrip_ok": R["e2e"]["ok"],
           "api": "b200_lz77_compress_host + b200_lz77_decompress_host (pinned host buffers)",
           "pageable": {"value": n_global / 1e9 / R["e2e"]["pageable_seconds"], "unit": UNIT, "steps": 3, "roundtrip_ok": R["e2e"]["pageable_ok"],
                        "note": "same calls with malloc'd (pageable) buffers, what an unmodified reference driver passes"}}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": {"workload": "deflate (algorithms/deflate LZ77 match finder + greedy parse + byte-token emission + block "
                                   "compaction) compress + decompress of ONE enwik9-shaped buffer sharded by block over the GPUs, BASELINE.json configs[3]",
                       "bytes": n_global, "bytes_rank0": n, "block_size": BLOCK, "blocks": (n_global + BLOCK - 1) // BLOCK, "blocks_rank0": R["nblocks"],
                       "token_bytes": int(tot_T.item()),
                       "l2_policy": "every rank's input shard and token stream (>= 125 MB + 158 MB at 8 GPUs) exceed the 126 MB L2; no flush needed",
                       "sharding": "contiguous block ranges per rank (sharding.byte_range); the only collective is one all-gather of the shard sizes "
                                   "(world x int64) per step; the payload stays sharded (decode needs nothing else)",
                       "cpu_affinity_rank0": numa},
            "roundtrip_ok": bool(all_ok.item() == 1), "gpu_launches": R["launches"], "clocks": R["clocks"], "e2e": e2e, "roofline": roofline}

    if world > 1 and not args.no_weak:
        # extra: weak scaling (every rank owns a full n_global-byte shard of a world x larger buffer)
        del R["st"], R["d_in"], R["h_in"]
        Wk = run_config(rank * n_global, n_global, max(2, min(args.steps, 3)), 3, sampler=None, want_e2e=False)
        line["weak_scaling"] = {"value": world * n_global / 1e9 / (Wk["ms_per_step"] / 1e3), "unit": UNIT, "ms_per_step": Wk["ms_per_step"],
                                "bytes_per_gpu": n_global, "roundtrip_ok": Wk["ok"], "note": "1 GB per rank, device resident; not the headline"}

    if rank == 0 and world == 1 and not args.no_cpu:
        st = R["st"]
        nb = min(CPU_SAMPLE // BLOCK, R["nblocks"])
        g_off = st.block_off[: nb + 1].cpu().numpy()
        g_tok = st.out[: int(g_off[-1])].cpu().numpy()
        g_sizes = st.block_sizes[:nb].cpu().numpy()
        line["cpu_baseline"] = cpu_baseline(R["h_in"].numpy()[:n], CPU_SAMPLE, gpu_stream=(g_tok, g_off, g_sizes), one_core_bytes=16 << 20)
    v4_line = None
    if rank == 0 and world == 1 and not args.no_detail:
        # both match finders on the SAME buffer as the headline (lz77_v2_kernel: B200_LZ_V4=0, lz77_v4_kernel: =1): their streams
        # must equal the headline's byte for byte; encode times by CUDA events
        try:
            st0 = R["st"]
            ref_total = st0.total_bytes
            ref_out = st0.out[: ref_total].clone(); ref_sizes = st0.block_sizes.clone()
            def _enc_ms(reps=3):
                ts = []
                for _ in range(reps):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(); dv.lz77_encode(ctx, R["d_in"], dv.LZ_DEFLATE, BLOCK, stream=st0, sync=False); e1.record()
                    ctx.sync(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
                return min(ts)
            os.environ["B200_LZ_V4"] = "0"
            dv.lz77_encode(ctx, R["d_in"], dv.LZ_DEFLATE, BLOCK, stream=st0, sync=True)
            t_def = _enc_ms()
            st2 = dv.lz77_encode(ctx, R["d_in"], dv.LZ_DEFLATE, BLOCK, stream=st0, sync=True)
            same_v2 = st2.total_bytes == ref_total and bool(torch.equal(st2.block_sizes, ref_sizes)) and bool(torch.equal(st2.out[: ref_total], ref_out))
            os.environ["B200_LZ_V4"] = "1"
            dv.lz77_encode(ctx, R["d_in"], dv.LZ_DEFLATE, BLOCK, stream=st0, sync=True)
            t_v4 = _enc_ms()
            st4 = dv.lz77_encode(ctx, R["d_in"], dv.LZ_DEFLATE, BLOCK, stream=st0, sync=True)
            same = st4.total_bytes == ref_total and bool(torch.equal(st4.block_sizes, ref_sizes)) and bool(torch.equal(st4.out[: ref_total], ref_out))
            v4_line = {"bytes": int(n), "v2_encode_ms": t_def, "v4_encode_ms": t_v4, "streams_equal": bool(same and same_v2),
                       "note": "whole encode call (match finder + size scan + compaction) with B200_LZ_V4=0 / =1; the headline runs the default: a "
                               "byte-entropy sample of the input picks lz77_v4_kernel for text-like input, lz77_v2_kernel otherwise (DESIGN.md 5c)"}
            del ref_out, ref_sizes
        except Exception as e:
            v4_line = {"error": repr(e)}
        finally:
            os.environ.pop("B200_LZ_V4", None)
        h100 = R["h_in"][:100_000_000]
        d100 = R["d_in"][:100_000_000].contiguous()
        del R["st"], R["d_in"]
        try:
            line["detail"] = detail_codecs(ctx, dv, torch, d100, h100)
            line["detail"]["lz77_match_finders"] = v4_line
        except Exception as e:  # secondary figures must never lose the headline line
            import traceback
            line["detail"] = {"error": repr(e), "trace": traceback.format_exc()[-600:]}
        try:
            del d100
            line["detail"]["block_sweep"] = block_sweep(ctx, dv, torch, corpus)
        except Exception as e:
            line["detail"]["block_sweep"] = {"error": repr(e)}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    import faulthandler
    faulthandler.enable()
    # stdout carries the ONE JSON line and nothing else: libraries that print to fd 1 (NCCL's version banner under
    # NCCL_DEBUG=VERSION, for one) are sent to stderr, the line goes to the saved descriptor.
    sys.stdout.flush()
    _real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = _real_stdout
    main()
    _real_stdout.flush()
