"""N > 1 orchestration of the sharded commitment (hobbit_b200/dist.py) on CPU: world_size 2 and 4 over gloo, the compute backend
being the C oracle.  Checks the exchange layout / chunk order / subtree assembly against the single-process oracle commitment.
(The same orchestration with GpuBackend + NCCL is exercised by tests/test_dist_gpu.py and bench.py --gpus N.)"""
import ctypes
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import Checker, _p, rand_field, srand


class OracleBackend:
    def __init__(self):
        self.orc = Checker("orc")

    def zeros(self, *shape):
        return torch.zeros(shape, dtype=torch.uint8)

    def empty(self, *shape):
        return torch.empty(shape, dtype=torch.uint8)

    def sync(self):
        pass

    def encode_chunks(self, poly, nchunks, B, trs, lin, first=0, parts=1, total=0):
        inner = np.zeros((nchunks, B, 32), dtype=np.uint8)
        src = np.ascontiguousarray(poly[first * B:(first + nchunks) * B])
        self.orc.fn("commit_encode_chunks")(_p(src), ctypes.c_size_t(nchunks), ctypes.c_size_t(B), trs, int(lin), _p(inner))
        # exchange layout [parts][nchunks][B/parts][32] (the CUDA kernel writes it directly)
        return torch.from_numpy(np.ascontiguousarray(inner.reshape(nchunks, parts, B // parts, 32).transpose(1, 0, 2, 3)))

    def encode_groups(self, chunks, ngroups, B, trs, lin, first=0, parts=1, total=0):
        inner = np.zeros((ngroups, 4 * B, 32), dtype=np.uint8)
        src = np.ascontiguousarray(chunks[first * 4 * B:(first + ngroups) * 4 * B])
        self.orc.fn("elastic_encode_groups")(_p(src), ctypes.c_size_t(ngroups), ctypes.c_size_t(B), trs, int(lin), _p(inner))
        return torch.from_numpy(np.ascontiguousarray(inner.reshape(ngroups, parts, 4 * B // parts, 32).transpose(1, 0, 2, 3)))

    def chain(self, inner, leaves):
        i = np.ascontiguousarray(inner.numpy()); l = leaves.numpy()
        self.orc.fn("md_chain")(_p(i), ctypes.c_size_t(i.shape[0]), ctypes.c_size_t(i.shape[1]), _p(l))
        return leaves

    def tree(self, leaves):
        return torch.from_numpy(self.orc.create_tree(np.ascontiguousarray(leaves.numpy())))


def free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def worker(rank, world, port, lin, K, B, trs, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hobbit_b200.dist import commit_standard_sharded
    be = OracleBackend()
    if lin:
        srand(1); be.orc.expander_init_store(trs)
    poly = rand_field(np.random.default_rng(77), K * B, full=False)
    kl = K // world
    levels = commit_standard_sharded(be, np.ascontiguousarray(poly[rank * kl * B:(rank + 1) * kl * B]), K, B, trs, lin)
    want, _ = be.orc.commit_standard(poly, K, trs, lin)
    ok = np.array_equal(levels.numpy(), want)
    t = torch.tensor([1 if ok else 0])
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        ret.put(int(t.item()))
    dist.destroy_process_group()


def elastic_worker(rank, world, port, lin, ngroups, B, trs, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hobbit_b200.dist import elastic_commit_sharded
    be = OracleBackend()
    if lin:
        srand(1); be.orc.expander_init_store(trs)
    stream = rand_field(np.random.default_rng(78), ngroups * 4 * B, full=False)
    stream[5 * B:6 * B] = 0                                   # an all-zero chunk (the reference skips its encode)
    gl = ngroups // world
    levels = elastic_commit_sharded(be, np.ascontiguousarray(stream[rank * gl * 4 * B:(rank + 1) * gl * 4 * B]), ngroups, B, trs, lin)
    want = np.zeros((8 * B - 1, 32), dtype=np.uint8)
    be.orc.fn("elastic_commit_stream")(_p(stream), ctypes.c_size_t(len(stream)), ctypes.c_size_t(B), trs, int(lin), _p(want))
    ok = np.array_equal(levels.numpy(), want)
    t = torch.tensor([1 if ok else 0])
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        ret.put(int(t.item()))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,lin", [(2, 1), (4, 0)])
def test_sharded_elastic_commit_gloo(world, lin):
    """Elastic_PC commit sharded by groups of 4 chunks == the single-process restatement (every level)."""
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=elastic_worker, args=(r, world, port, lin, 4, 1 << 9, 16, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert ret.get(timeout=10) == 1


@pytest.mark.parametrize("world,lin", [(2, 1), (2, 0), (4, 1)])
def test_sharded_commit_gloo(world, lin):
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=worker, args=(r, world, port, lin, 8, 1 << 10, 16, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert ret.get(timeout=10) == 1


class OracleSumcheckBackend:
    def __init__(self):
        self.orc = Checker("orc")

    def sc3_round(self, cur, L, rand):
        outs = [np.zeros((L, 2), dtype=np.uint64) for _ in range(3)]
        co = np.zeros((4, 2), dtype=np.uint64)
        r = np.ascontiguousarray(rand, dtype=np.uint64)
        cur = [np.ascontiguousarray(c) for c in cur]
        self.orc.fn("sc3_round")(_p(cur[0]), _p(cur[1]), _p(cur[2]), _p(outs[0]), _p(outs[1]), _p(outs[2]), ctypes.c_size_t(L), _p(r), _p(co))
        return co, outs

    def heads(self, cur):
        return np.stack([c[0] for c in cur])

    def tables_from(self, small):
        return [np.ascontiguousarray(small[k]) for k in range(3)]

    def to_comm(self, a):
        return torch.from_numpy(np.ascontiguousarray(a).view(np.int64).copy())

    def from_comm(self, t):
        return t.numpy().view(np.uint64)


def sc_worker(rank, world, port, n, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hobbit_b200.dist import sumcheck3_sharded
    be = OracleSumcheckBackend()
    rng = np.random.default_rng(3)
    v = [rand_field(rng, n) for _ in range(3)]
    pr = rand_field(rng, 1)
    nl = n // world
    got = sumcheck3_sharded(be, [x[rank * nl:(rank + 1) * nl] for x in v], nl, pr, be.orc.mimc)
    want, _ = be.orc.sumcheck3(v[0], v[1], v[2], pr)
    t = torch.tensor([1 if np.array_equal(got, want) else 0])
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        ret.put(int(t.item()))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 256), (4, 64), (2, 2)])
def test_sharded_sumcheck_gloo(world, n):
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=sc_worker, args=(r, world, port, n, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert ret.get(timeout=10) == 1
