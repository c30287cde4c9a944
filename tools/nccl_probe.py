"""Development probe: NCCL all_to_all / all_gather bandwidth between the ranks of one box (run under torchrun)."""
import os, time, torch, torch.distributed as dist
local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
G = dist.get_world_size()
for mib in (64, 256, 1024):
    n = mib << 20
    a = torch.empty(n, dtype=torch.uint8, device="cuda"); b = torch.empty_like(a)
    for _ in range(3): dist.all_to_all_single(b, a)
    torch.cuda.synchronize(); dist.barrier(); t = time.perf_counter()
    for _ in range(5): dist.all_to_all_single(b, a)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 5
    if dist.get_rank() == 0: print("all_to_all %4d MiB per rank: %.2f ms -> %.0f GB/s sent per rank" % (mib, dt * 1e3, n * (G - 1) / G / dt / 1e9))
    t = time.perf_counter()
    for _ in range(5): b.copy_(a)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 5
    if dist.get_rank() == 0: print("   local copy %4d MiB: %.2f ms" % (mib, dt * 1e3))
dist.destroy_process_group()
