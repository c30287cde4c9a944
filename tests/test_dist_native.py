"""The multi-GPU layer of the C ABI (hb_dist_*): ranks are separate processes that exchange data through CUDA-IPC windows.
When the box has fewer GPUs than ranks the ranks share GPU 0 (two processes, two contexts, time-sliced): the peer-window protocol is the
same, so the driver's single-GPU box still exercises it.  Every result is compared with the single-GPU entry point on the same input."""
import os
import socket

import numpy as np
import pytest

from helpers import Checker, rand_field, srand

pytestmark = pytest.mark.gpu


def _setup(rank, world, port, trs, data_bytes):
    import torch
    import torch.distributed as dist
    import hobbit_b200
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dev = rank if torch.cuda.device_count() >= world else 0
    torch.cuda.set_device(dev)
    dist.init_process_group("gloo", rank=rank, world_size=world)          # bootstrap only: the data path is the library's own
    ctx = hobbit_b200.Context(dev)
    if trs:
        orc = Checker("orc")
        srand(1); orc.expander_init_store(trs)
        ctx.expander_set(trs, orc.expander_graphs(trs))
    ctx.dist_init_torch(data_bytes)
    return ctx, dist


def _finish(ctx, dist, ok, rank, ret):
    import torch
    t = torch.tensor([1 if ok else 0])
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        ret.put(int(t.item()))
    ctx.dist_disconnect()
    dist.destroy_process_group()
    ctx.close()


def _commit_worker(rank, world, port, ret):
    K, B, trs = 8, 1 << 12, 16
    ctx, dist = _setup(rank, world, port, trs, K * (B // world) * 32 + 2 * B * 32 + 4096 + 4 * 4 * (1 << 12) * 32 * 2)
    ok = True
    for lin in (1, 0):
        poly = rand_field(np.random.default_rng(6 + lin), K * B, full=False)
        kl = K // world
        for rep in range(2):                                               # twice: the window is reused, barriers must separate the calls
            got = ctx.dist_commit_standard(np.ascontiguousarray(poly[rank * kl * B:(rank + 1) * kl * B]), K, B, trs, lin)
            want, _ = ctx.commit_standard(poly, K, trs, lin)
            ok = ok and np.array_equal(got, want)
    # Elastic_PC: groups of 4 chunks
    ngroups, Be = 4, 1 << 12
    stream = rand_field(np.random.default_rng(16), ngroups * 4 * Be, full=True)
    stream[2 * Be:3 * Be] = 0
    gl = ngroups // world
    got = ctx.dist_elastic_commit(np.ascontiguousarray(stream[rank * gl * 4 * Be:(rank + 1) * gl * 4 * Be]), ngroups, Be, trs, 1)
    want = ctx.elastic_commit([stream[i * Be:(i + 1) * Be] for i in range(len(stream) // Be)], Be, trs, 1)
    ok = ok and np.array_equal(got, want)
    # field all-reduce
    v = rand_field(np.random.default_rng(100 + rank), 1000, full=True)
    s = ctx.dist_allreduce(v)
    P = (1 << 61) - 1
    exp = np.zeros_like(v)
    for r in range(world):
        exp = (exp + rand_field(np.random.default_rng(100 + r), 1000, full=True)) % np.uint64(P)
    ok = ok and np.array_equal(s, exp)
    _finish(ctx, dist, ok, rank, ret)


def _run(worker, world, timeout=600):
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    c = mp.get_context("spawn")
    ret = c.Queue()
    procs = [c.Process(target=worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout)
        assert p.exitcode == 0
    assert ret.get(timeout=10) == 1


@pytest.mark.parametrize("world", [2, 4])
def test_native_sharded_commits(world):
    _run(_commit_worker, world)
