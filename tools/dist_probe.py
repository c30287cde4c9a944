"""Development probe: per-step timing of the sharded commitment on N GPUs (run under torchrun)."""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import Checker, srand
import hobbit_b200
from hobbit_b200.dist import GpuBackend

local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
G, rank = dist.get_world_size(), dist.get_rank()
ctx = hobbit_b200.Context(local)
orc = Checker("orc"); srand(1); orc.expander_init_store(1024); ctx.expander_set(1024, orc.expander_graphs(1024))
kl, B, trs = 32, 1 << 21, 1024
K, Bp = kl * G, B // G
poly = torch.randint(0, 1 << 31, (kl * B, 2), dtype=torch.int64, device="cuda"); poly[:, 1] = 0
be = GpuBackend(ctx, torch.device("cuda", local))

def T(name, fn, reps=3):
    fn(); torch.cuda.synchronize(); dist.barrier()
    t = time.perf_counter()
    for _ in range(reps): r = fn()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / reps
    if rank == 0: print("%-34s %.2f ms" % (name, dt * 1e3))
    return r

inner = T("encode 32 chunks (1 call)", lambda: be.encode_chunks(poly.data_ptr(), kl, B, trs, 1, 0))
T("encode 32 chunks (4 calls of 8)", lambda: [be.encode_chunks(poly.data_ptr(), 8, B, trs, 1, g * 8) for g in range(4)][0])
send = T("permute+contiguous 1 GiB", lambda: inner.view(kl, G, Bp, 32).permute(1, 0, 2, 3).contiguous())
recv = torch.empty_like(send)
T("all_to_all_single 1 GiB", lambda: dist.all_to_all_single(recv, send))
inner_all = recv.view(K, Bp, 32)
leaves = T("chain %d chunks x %d leaves" % (K, Bp), lambda: be.chain(inner_all, be.zeros(Bp, 32)))
sub = T("subtree", lambda: be.tree(leaves))
allsub = [torch.empty_like(sub) for _ in range(G)]
T("all_gather subtree levels", lambda: dist.all_gather(allsub, sub))
T("torch.empty 1 GiB + zeros", lambda: (be.empty(kl, B, 32), be.zeros(Bp, 32)))
dist.destroy_process_group()
