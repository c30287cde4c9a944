"""Sumcheck sharded by hypercube prefix at a size where it pays (torchrun, one rank per GPU): _generate_3product_sumcheck_proof over
2^LOGN-entry tables, replicated on every rank, proved (a) by every rank alone and (b) sharded (hb_dist_shard: each rank folds its
contiguous 1/N of every table, the round sums are added across ranks inside the round kernel).  Prints one JSON line on rank 0."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hobbit_b200  # noqa: E402
from hobbit_b200 import DevF  # noqa: E402

logn = int(sys.argv[1]) if len(sys.argv) > 1 else 26
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = hobbit_b200.Context(local)
if world > 1:
    ctx.dist_init_torch(1 << 20)
g = torch.Generator(device="cuda"); g.manual_seed(7)
tabs = [torch.randint(0, (1 << 61) - 1, (1 << logn, 2), dtype=torch.int64, device="cuda", generator=g) for _ in range(3)]
dv = [DevF.from_torch(t) for t in tabs]
pr = np.array([[5, 7]], dtype=np.uint64)


def run(shard):
    ctx.dist_shard(shard)
    best, proof = 1e9, None
    for _ in range(4):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        proof, _ = ctx.sumcheck3(dv[0], dv[1], dv[2], pr)
        best = min(best, time.perf_counter() - t0)
    ctx.dist_shard(False)
    return best, proof


t_single, p_single = run(False)
res = {"workload": "_generate_3product_sumcheck_proof over 2^%d-entry tables" % logn, "n_gpus": world, "single_gpu_ms": 1e3 * t_single}
if world > 1:
    t_sh, p_sh = run(True)
    t = torch.tensor([t_sh], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res.update({"sharded_ms": 1e3 * float(t.item()), "speedup": t_single / float(t.item()), "proofs_identical": bool(np.array_equal(p_single, p_sh))})
if rank == 0:
    print(json.dumps(res))
if world > 1:
    ctx.dist_disconnect()
    dist.destroy_process_group()
