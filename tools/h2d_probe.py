"""Host<->device copy ceiling of the box with all ranks copying at once (torchrun): what bounds bench.py's e2e number at N > 1.
Every rank copies 1 GiB from pinned host memory to its GPU (and 128 MiB back); reports per-rank and aggregate GB/s, alone vs together."""
import json
import os
import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
h = torch.empty(1 << 30, dtype=torch.uint8).pin_memory(); h.fill_(1)
d = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
hb = torch.empty(1 << 27, dtype=torch.uint8).pin_memory()


def timed(fn, reps=5):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    t = torch.tensor([best], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


res = {"world": world}
ms = timed(lambda: d.copy_(h, non_blocking=True))
res["h2d_all_ranks_ms_per_GiB"] = ms; res["h2d_aggregate_GBps"] = world * (1 << 30) / ms / 1e6
ms = timed(lambda: hb.copy_(d[:1 << 27], non_blocking=True))
res["d2h_all_ranks_ms_per_128MiB"] = ms; res["d2h_aggregate_GBps"] = world * (1 << 27) / ms / 1e6
if world > 1:                       # rank 0 alone, the others idle
    if rank == 0:
        best = 1e9
        for _ in range(5):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); d.copy_(h, non_blocking=True); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        res["h2d_one_rank_alone_GBps"] = (1 << 30) / best / 1e6
    dist.barrier()
if rank == 0:
    print(json.dumps(res))
if world > 1:
    dist.destroy_process_group()
