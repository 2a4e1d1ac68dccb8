#!/usr/bin/env python
"""Headline benchmark (BASELINE.json): deflate compress + decompress of an
enwik9-shaped 1 GB buffer cut into 64 KiB blocks, per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A step = one pass of the hot path over one batch: the repo's deflate LZ77 match
finder + greedy parse + token emission + block compaction of the whole buffer
(algorithms/deflate `compress`), followed by the decode of that stream.
  value     : device-resident throughput, uncompressed GB / (compress + decompress) s,
              whole job over all N GPUs (weak scaling: every rank owns a 1 GB shard and
              the ranks all-gather their shard sizes to place the shards).
  e2e       : the same through the host-buffer C-ABI calls a reference driver would
              make (b200_lz77_compress_host / _decompress_host), pinned host buffers,
              H2D/D2H copies inside the timed region.
  roofline  : the dominant kernel (LZ77 parse) against measured HBM copy bandwidth.
  cpu_baseline : the reference's own C (oracle/_ref, built from /root/reference) timed on
              this box's host cores on a bounded sample, blocks spread over all cores.
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

METRIC = "compress/decompress GB/s (deflate), enwik9-shaped 1 GB per GPU"
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


def cpu_baseline(data_np, sample_bytes, threads=0):
    """The reference's deflate lz77_compress per 64 KiB block on a fresh table (oracle/_ref when
    present, else the oracle port), blocks spread over all host cores; decode = the oracle
    port's byte-token decoder (the reference ships none, deflate.c:78-79)."""
    from oracle import bindings as ob
    n = min(sample_bytes, data_np.size)
    sample = np.ascontiguousarray(data_np[:n])
    try:
        cores = len(os.sched_getaffinity(0)) or 1
    except AttributeError:
        cores = os.cpu_count() or 1
    if threads <= 0:
        threads = cores     # explicit: torchrun exports OMP_NUM_THREADS=1, which would leave the harness on one thread
    kind = "reference" if ob.have_ref() else "port"
    t0 = time.perf_counter()
    if kind == "reference":
        blocks, sizes = ob.ref_deflate_lz77_compress_blocks(sample, BLOCK, persistent=False, threads=threads)
    else:
        out, sizes = ob.port_lz77_compress_blocks(sample, BLOCK, 1, threads)
        blocks = [out[b, : int(sizes[b])] for b in range(len(sizes))]
    t1 = time.perf_counter()
    stream = np.concatenate(blocks)
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
    t2 = time.perf_counter()
    dec, bad = ob.port_lz77_decompress_blocks(stream, off, BLOCK, n, 1, threads)
    t3 = time.perf_counter()
    ok = bad == 0 and np.array_equal(dec, sample)
    tc, td = t1 - t0, t3 - t2
    return {"value": n / 1e9 / (tc + td), "unit": UNIT, "cores": threads, "kind": kind,
            "sample": "first %d MiB of the workload, %d blocks of 64 KiB, one %s lz77_compress call per block on a fresh table "
                      "spread over %d threads; decode by the oracle port (the reference has no deflate decoder)"
                      % (n >> 20, len(sizes), "reference" if kind == "reference" else "oracle-port", threads),
            "compress_gbps": n / 1e9 / tc, "decompress_gbps": n / 1e9 / td, "roundtrip_ok": bool(ok),
            "token_bytes": int(stream.size)}


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
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "deflate (algorithms/deflate lz77_compress per 64 KiB block, fresh table) compress+decompress; "
                                   "each step = a %d MiB sample of the enwik9-shaped 1 GB buffer on the host CPU" % (CPU_SAMPLE >> 20),
                       "block_size": BLOCK, "bytes_per_step": CPU_SAMPLE},
            "cpu_baseline": cb,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def detail_codecs(ctx, dv, torch, data100):
    """Secondary figures (BASELINE.json configs 0-2) at 100 MB, device resident."""
    out = {}
    n = data100.numel()

    def timed(fn, reps=3):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            r = fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps / 1e3, r

    for name, block in (("huffman_whole_buffer", 0), ("huffman_64k_blocks", BLOCK)):
        st = dv.huffman_alloc(ctx, n, block)
        te, st = timed(lambda: dv.huffman_encode(ctx, data100, block, stream=st, sync=False))
        st = dv.huffman_encode(ctx, data100, block, stream=st)
        dec = torch.empty_like(data100)
        td, _ = timed(lambda: dv.huffman_decode(ctx, st, out=dec))
        out[name] = {"compress_gbps": n / 1e9 / te, "decompress_gbps": n / 1e9 / td, "ratio": n / (st.total_words * 4.0),
                     "roundtrip_ok": bool(torch.equal(dec, data100))}
    st = dv.fse_alloc(ctx, n, BLOCK, dv.DEFAULT_FSE_SEG)
    te, st = timed(lambda: dv.fse_encode(ctx, data100, BLOCK, dv.DEFAULT_FSE_SEG, stream=st, sync=False))
    st = dv.fse_encode(ctx, data100, BLOCK, dv.DEFAULT_FSE_SEG, stream=st)
    dec = torch.empty_like(data100)
    td, _ = timed(lambda: dv.fse_decode(ctx, st, out=dec, sync=False))
    out["fse_64k_blocks_1k_segments"] = {"compress_gbps": n / 1e9 / te, "decompress_gbps": n / 1e9 / td,
                                         "ratio": n / (st.total_words * 8.0), "roundtrip_ok": bool(torch.equal(dec, data100))}
    st = dv.lz77_alloc(ctx, n, BLOCK, dv.LZ_STANDALONE)
    te, st = timed(lambda: dv.lz77_encode(ctx, data100, dv.LZ_STANDALONE, BLOCK, stream=st, sync=False), reps=2)
    st = dv.lz77_encode(ctx, data100, dv.LZ_STANDALONE, BLOCK, stream=st)
    dec = torch.empty_like(data100)
    td, _ = timed(lambda: dv.lz77_decode(ctx, st, out=dec), reps=2)
    out["lz77_standalone_64k_blocks"] = {"compress_gbps": n / 1e9 / te, "decompress_gbps": n / 1e9 / td,
                                         "ratio": n / float(st.total_bytes), "roundtrip_ok": bool(torch.equal(dec, data100))}
    # deflate with the entropy stage (the reference's TODO, deflate/lz77.c:279): LZ77 tokens -> Huffman-coded words
    ds = dv.deflate_alloc(ctx, n, BLOCK)
    te, ds = timed(lambda: dv.deflate_compress(ctx, data100, BLOCK, stream=ds, sync=False), reps=2)
    ds = dv.deflate_compress(ctx, data100, BLOCK, stream=ds)
    tes, _ = timed(lambda: dv.dfl_encode(ctx, ds.lz, stream=ds, sync=False))
    dec = torch.empty_like(data100)
    td, _ = timed(lambda: dv.deflate_decompress(ctx, ds, out=dec), reps=2)
    tok = torch.empty_like(ds.lz.out)
    tds, _ = timed(lambda: dv.dfl_decode(ctx, ds, tok))
    out["deflate_entropy_coded_64k_blocks"] = {
        "compress_gbps": n / 1e9 / te, "decompress_gbps": n / 1e9 / td, "ratio": n / (ds.total_words * 4.0),
        "entropy_stage_encode_ms": tes * 1e3, "entropy_stage_decode_ms": tds * 1e3,
        "token_bytes": int(ds.lz.block_off[-1].item()), "stream_bytes": int(ds.total_words * 4),
        "roundtrip_ok": bool(torch.equal(dec, data100))}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--bytes", type=int, default=N_BYTES, help="bytes per GPU (default: the 1 GB headline config)")
    ap.add_argument("--no-detail", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
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
        # one process per GPU: run (and first-touch the pinned host buffers) on the CPUs next to this rank's GPU, so that
        # the host-buffer leg does not cross the socket interconnect
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
    n = args.bytes
    nblocks = (n + BLOCK - 1) // BLOCK

    # ---- synthetic shard of this rank, pinned on the host, then resident in HBM
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    corpus.generate(n, corpus.ENWIK, corpus.DEFAULT_SEED + rank, out=h_in.numpy())
    d_in = h_in.to(ctx.device, non_blocking=False)
    st = dv.lz77_alloc(ctx, n, BLOCK, dv.LZ_DEFLATE)
    d_dec = torch.empty(n, dtype=torch.uint8, device=ctx.device)

    def step():
        dv.lz77_encode(ctx, d_in, dv.LZ_DEFLATE, BLOCK, stream=st, sync=False)
        if world > 1:
            # the one exchange of the path: all-gather the shard sizes -> global offset of every shard
            sharding.exchange_sizes(st.block_off[-1:], ctx.device)
        dv.lz77_decode(ctx, st, out=d_dec)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    sampler.wait_first()
    for _ in range(args.warmup):
        step()
    barrier()
    launches0 = ctx.launches
    ctx.set_timing(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.mark_begin()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=ctx.device)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step = ms.item() / args.steps
    launches = ctx.launches - launches0
    timings = ctx.timings()
    ctx.set_timing(False)
    value = world * n / 1e9 / (ms_per_step / 1e3)

    # correctness of what was timed + sizes
    st = dv.lz77_encode(ctx, d_in, dv.LZ_DEFLATE, BLOCK, stream=st, sync=True)
    dv.lz77_decode(ctx, st, out=d_dec)
    ok = bool(torch.equal(d_dec, d_in))
    T = st.total_bytes
    parse_ms = [m for k, m in timings if k == 0]
    dec_ms = [m for k, m in timings if k == 1]
    peak, peak_src = hbm_peak()
    parse_avg = float(np.mean(parse_ms)) if parse_ms else float("nan")
    achieved = (n + T) / 1e9 / (parse_avg / 1e3)
    traffic = None
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tp):
        try:
            with open(tp) as f:
                per_byte = json.load(f).get("lz77_v2_kernel<1>", {}).get("dram_bytes_per_input_byte")
            traffic = int(per_byte * n) if per_byte else None   # ncu capture at 100 MB, scaled to this launch's input
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": "lz77_v2_kernel<1> (deflate-variant match finder: shared-memory table simulation + greedy parse + token emission)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": n + T, "avg_launch_ms": parse_avg,
                "launches_timed": len(parse_ms), "share_of_step": parse_avg / ms_per_step,
                "decode_kernel_avg_ms": float(np.mean(dec_ms)) if dec_ms else None}

    # ---- e2e: the host-buffer C-ABI path with the copies in the timed region
    lib = _lib.core()
    cap = int(lib.b200_lz77_max_bytes(dv.LZ_DEFLATE, n, BLOCK))
    h_out = torch.empty(cap, dtype=torch.uint8).pin_memory()
    h_sizes = torch.empty(nblocks, dtype=torch.int64).pin_memory()
    h_off = torch.empty(nblocks + 1, dtype=torch.int64).pin_memory()
    h_dec = torch.empty(n, dtype=torch.uint8).pin_memory()
    tot = C.c_uint64(0)

    def e2e_step():
        _lib.check(lib.b200_lz77_compress_host(ctx.handle, dv.LZ_DEFLATE, h_in.data_ptr(), n, BLOCK, h_out.data_ptr(), cap,
                                               h_sizes.data_ptr(), h_off.data_ptr(), C.byref(tot)))
        _lib.check(lib.b200_lz77_decompress_host(ctx.handle, dv.LZ_DEFLATE, h_out.data_ptr(), tot.value, h_off.data_ptr(),
                                                 h_sizes.data_ptr(), n, BLOCK, h_dec.data_ptr()))

    e2e_step()
    barrier()
    e2e_steps = max(2, min(args.steps, 3))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    t_e2e = torch.tensor([(time.perf_counter() - t0) / e2e_steps], dtype=torch.float64, device=ctx.device)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_ok = bool(np.array_equal(h_dec.numpy(), h_in.numpy()))
    idx_bytes = (2 * nblocks + 1) * 8
    e2e = {"value": world * n / 1e9 / t_e2e.item(), "unit": UNIT,
           "h2d_bytes_per_step": int(n + tot.value + idx_bytes), "d2h_bytes_per_step": int(tot.value + idx_bytes + n),
           "steps": e2e_steps, "roundtrip_ok": e2e_ok,
           "api": "b200_lz77_compress_host + b200_lz77_decompress_host (pinned host buffers)"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": {"workload": "deflate (algorithms/deflate LZ77 match finder + greedy parse + byte-token emission + block "
                                   "compaction) compress + decompress of an enwik9-shaped buffer, BASELINE.json configs[3]",
                       "bytes_per_gpu": n, "block_size": BLOCK, "blocks_per_gpu": nblocks, "token_bytes_per_gpu": int(T),
                       "l2_policy": "input (1 GB) and token stream are far larger than the 126 MB L2; no flush needed",
                       "sharding": "contiguous block ranges per rank; all-gather of shard sizes only", "cpu_affinity_rank0": numa},
            "roundtrip_ok": ok, "gpu_launches": int(launches), "clocks": clocks, "e2e": e2e, "roofline": roofline}

    if rank == 0 and world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline(h_in.numpy(), CPU_SAMPLE)
    if rank == 0 and world == 1 and not args.no_detail:
        del d_dec
        try:
            line["detail"] = detail_codecs(ctx, dv, torch, d_in[:100_000_000].contiguous())
        except Exception as e:  # secondary figures must never lose the headline line
            line["detail"] = {"error": repr(e)}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    # stdout carries the ONE JSON line and nothing else: libraries that print to fd 1 (NCCL's version banner under
    # NCCL_DEBUG=VERSION, for one) are sent to stderr, the line goes to the saved descriptor.
    sys.stdout.flush()
    _real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = _real_stdout
    main()
    _real_stdout.flush()
