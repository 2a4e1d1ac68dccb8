"""Aggregate host<->device copy bandwidth with N ranks copying at once (pinned buffers, 1 GB each way per rank):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29631 tools/pcie_ranks.py
Explains the host-buffer (`e2e`) leg of bench.py at N > 1: the device-resident path scales with the GPUs, the
host-buffer path with what the host side of the box can move."""
import json, os, time
import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1_000_000_000
h_a = torch.empty(n, dtype=torch.uint8).pin_memory(); h_a.fill_(rank + 1)
h_b = torch.empty(n, dtype=torch.uint8).pin_memory(); h_b.fill_(0)
d_a = torch.empty(n, dtype=torch.uint8, device="cuda"); d_b = torch.full((n,), 7, dtype=torch.uint8, device="cuda")
s2 = torch.cuda.Stream()

def timed(fn, reps=4):
    best = 1e9
    for _ in range(reps + 1):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        t = torch.tensor([dt], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        best = min(best, t.item())
    return best

def both():
    d_a.copy_(h_a, non_blocking=True)
    with torch.cuda.stream(s2):
        h_b.copy_(d_b, non_blocking=True)

r = {"world": world, "bytes_per_rank_each_way": n}
for name, fn, nb in (("h2d", lambda: d_a.copy_(h_a, non_blocking=True), n), ("d2h", lambda: h_b.copy_(d_b, non_blocking=True), n), ("h2d_and_d2h", both, 2 * n)):
    t = timed(fn)
    r[name + "_ms"] = round(t * 1e3, 2)
    r[name + "_aggregate_GBps"] = round(world * nb / t / 1e9, 1)
if rank == 0:
    print(json.dumps(r))
if world > 1:
    dist.destroy_process_group()
