"""Whole-buffer Huffman (ONE tree for the data of all ranks) over real ranks:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
      tools/huff_whole_ranks.py [n_bytes]
Every rank holds a contiguous shard, the histogram is all-reduced over NCCL, the shard bit counts all-gathered,
the shard streams spliced; rank 0 compares the result with the single-GPU whole-buffer encode of the same bytes
(itself bit-exact with the reference, tests/test_gpu_huffman.py) and prints one JSON line with device timings."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from compression_algorithms_b200 import corpus, device as dv, sharding  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = dv.Context(local)
    data = corpus.generate(n, corpus.ENWIK, 20261018)
    per = (n + world - 1) // world
    per += (-per) % 16                       # not required, shards are cloned; keeps the cut off a word boundary of bits only
    a, b = min(rank * per + (7 if rank else 0), n), min((rank + 1) * per + 7, n) if rank + 1 < world else n
    shard = torch.from_numpy(data[a:b].copy()).to(ctx.device)
    hist_fn, encode_fn, splice_fn = sharding.huffman_fns(ctx)
    times = []
    for it in range(4):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        sh = sharding.huffman_whole_compress_sharded(shard, hist_fn, encode_fn)
        e1.record()
        whole, nbytes = sharding.huffman_whole_gather(sh, sh.stream.words, splice_fn)
        ctx.sync()
        e2.record()
        torch.cuda.synchronize()
        times.append((e0.elapsed_time(e1), e1.elapsed_time(e2)))
    dec = dv.huffman_decode(ctx, sh.stream)
    ok_dec = bool(torch.equal(dec, shard))
    ok = torch.tensor([1 if ok_dec else 0], device=ctx.device)
    if world > 1:
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        ref = dv.huffman_encode(ctx, torch.from_numpy(data.copy()).to(ctx.device), 0)
        same = ref.total_words == whole.numel() and bool(torch.equal(ref.words[: ref.total_words], whole))
        print(json.dumps({"tool": "huff_whole_ranks", "n": n, "world": world, "stream_equals_single_gpu_whole_buffer": same,
                          "buffer_size": nbytes, "shard_bits": sh.shard_bits.tolist(), "every_rank_decodes_its_shard": bool(ok.item()),
                          "ms_hist_allreduce_encode_sizes": round(min(t[0] for t in times[1:]), 3),
                          "ms_gather_splice": round(min(t[1] for t in times[1:]), 3)}))
        if not (same and ok.item()):
            sys.exit(1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
