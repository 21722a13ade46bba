"""Pinned-memory copy bandwidth of every rank AT THE SAME TIME (run under torchrun, one
rank per GPU): H2D, D2H and both directions together, 1 GiB buffers.  Explains the
ceiling of bench.py's `e2e` at N > 1: the host side of the box, not the codec.
Usage: torchrun --nproc-per-node N scripts/pcie_concurrent.py"""
import json, os, time
import torch
import torch.distributed as dist

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1 << 28  # fp32 elements = 1 GiB
h_in = torch.empty(n, dtype=torch.float32, pin_memory=True).fill_(1.0)
h_out = torch.empty(n, dtype=torch.float32, pin_memory=True)
d_a = torch.empty(n, dtype=torch.float32, device="cuda")
d_b = torch.ones(n, dtype=torch.float32, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(kind, reps=4):
    def once():
        if kind in ("h2d", "both"):
            with torch.cuda.stream(s1):
                d_a.copy_(h_in, non_blocking=True)
        if kind in ("d2h", "both"):
            with torch.cuda.stream(s2):
                h_out.copy_(d_b, non_blocking=True)
    once()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        once()
    torch.cuda.synchronize()
    dt = torch.tensor([(time.perf_counter() - t0) / reps], device="cuda")
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    moved = n * 4 * (2 if kind == "both" else 1)
    return moved / float(dt) / 1e9  # GB/s per rank at the slowest rank's pace


res = {k: run(k) for k in ("h2d", "d2h", "both")}
if rank == 0:
    print(json.dumps({"n_gpus": world, "per_rank_GBs": {k: round(v, 1) for k, v in res.items()},
                      "aggregate_GBs": {k: round(v * world, 1) for k, v in res.items()},
                      "note": "1 GiB pinned buffers, all ranks at once, max time over ranks"}))
if world > 1:
    dist.destroy_process_group()
