"""Sharded commitment on real GPUs: GpuBackend pieces on one GPU (== the monolithic hb_commit_standard), and — when the box has
at least 2 GPUs — 2 ranks over NCCL against the single-GPU commitment."""
import os
import socket

import numpy as np
import pytest

from helpers import Checker, rand_field, srand

pytestmark = pytest.mark.gpu


def setup_ctx(dev, trs):
    import hobbit_b200
    ctx = hobbit_b200.Context(dev)
    orc = Checker("orc")
    srand(1); orc.expander_init_store(trs)
    ctx.expander_set(trs, orc.expander_graphs(trs))
    return ctx


def test_backend_pieces_equal_monolithic_commit():
    import torch
    from hobbit_b200.dist import GpuBackend, commit_standard_sharded
    K, B, trs = 8, 1 << 12, 16
    ctx = setup_ctx(0, trs)
    poly = rand_field(np.random.default_rng(5), K * B, full=False)
    want, _ = ctx.commit_standard(poly, K, trs, 1)
    got = commit_standard_sharded(GpuBackend(ctx, torch.device("cuda", 0)), poly, K, B, trs, 1)
    assert np.array_equal(got.cpu().numpy(), want)
    ctx.close()


def _elastic_levels(ctx, stream, B, trs, lin):
    return ctx.elastic_commit([stream[i * B:(i + 1) * B] for i in range(len(stream) // B)], B, trs, lin)


@pytest.mark.parametrize("lin", [0, 1])
def test_elastic_group_pieces_equal_streaming_commit(lin):
    """hb_elastic_encode_groups + hb_md_chain + hb_merkle_tree (the sharding split) == hb_elastic_begin/push/finish, every level."""
    import torch
    from hobbit_b200.dist import GpuBackend, elastic_commit_sharded
    ngroups, B, trs = 3, 1 << 12, 16
    ctx = setup_ctx(0, trs)
    stream = rand_field(np.random.default_rng(15), ngroups * 4 * B, full=True)
    stream[2 * B:3 * B] = 0
    want = _elastic_levels(ctx, stream, B, trs, lin)
    got = elastic_commit_sharded(GpuBackend(ctx, torch.device("cuda", 0)), stream, ngroups, B, trs, lin)
    assert np.array_equal(got.cpu().numpy(), want)
    orc = Checker("orc")
    ref = np.zeros((8 * B - 1, 32), dtype=np.uint8)
    import ctypes
    orc.fn("elastic_commit_stream")(stream.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(len(stream)), ctypes.c_size_t(B), trs, lin, ref.ctypes.data_as(ctypes.c_void_p))
    assert np.array_equal(want, ref)
    ctx.close()


def test_elastic_groups_from_pinned_host_double_buffered(monkeypatch):
    """The stream in pinned HOST memory, one group per launch: the copy of launch g+1 overlaps the encode of launch g (two staging buffers)."""
    import torch
    from hobbit_b200.dist import GpuBackend, elastic_commit_sharded
    monkeypatch.setenv("HB_ELASTIC_GROUPS_PER_LAUNCH", "1")
    ngroups, B, trs = 5, 1 << 12, 16
    ctx = setup_ctx(0, trs)
    stream = rand_field(np.random.default_rng(17), ngroups * 4 * B, full=True)
    want = _elastic_levels(ctx, stream, B, trs, 1)
    pinned = torch.from_numpy(stream.view(np.int64)).pin_memory()
    got = elastic_commit_sharded(GpuBackend(ctx, torch.device("cuda", 0)), pinned.data_ptr(), ngroups, B, trs, 1)
    assert np.array_equal(got.cpu().numpy(), want)
    ctx.close()


def _elastic_worker(rank, world, port, ngroups, B, trs, ret):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from hobbit_b200.dist import GpuBackend, elastic_commit_sharded
    ctx = setup_ctx(rank, trs)
    stream = rand_field(np.random.default_rng(16), ngroups * 4 * B, full=True)
    gl = ngroups // world
    got = elastic_commit_sharded(GpuBackend(ctx, torch.device("cuda", rank)), np.ascontiguousarray(stream[rank * gl * 4 * B:(rank + 1) * gl * 4 * B]),
                                 ngroups, B, trs, 1)
    want = _elastic_levels(ctx, stream, B, trs, 1)
    ok = np.array_equal(got.cpu().numpy(), want)
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        ret.put(int(t.item()))
    dist.destroy_process_group()
    ctx.close()


def test_sharded_elastic_commit_nccl_2gpus():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2); the CPU/gloo twin is tests/test_dist_gloo.py")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    c = mp.get_context("spawn")
    ret = c.Queue()
    procs = [c.Process(target=_elastic_worker, args=(r, 2, port, 4, 1 << 12, 16, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert ret.get(timeout=10) == 1


def _worker(rank, world, port, K, B, trs, ret):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from hobbit_b200.dist import GpuBackend, commit_standard_sharded
    ctx = setup_ctx(rank, trs)
    poly = rand_field(np.random.default_rng(6), K * B, full=False)
    kl = K // world
    got = commit_standard_sharded(GpuBackend(ctx, torch.device("cuda", rank)), np.ascontiguousarray(poly[rank * kl * B:(rank + 1) * kl * B]), K, B, trs, 1)
    want, _ = ctx.commit_standard(poly, K, trs, 1)
    ok = np.array_equal(got.cpu().numpy(), want)
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        ret.put(int(t.item()))
    dist.destroy_process_group()
    ctx.close()


def test_sharded_commit_nccl_2gpus():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2); the CPU/gloo twin is tests/test_dist_gloo.py")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    c = mp.get_context("spawn")
    ret = c.Queue()
    procs = [c.Process(target=_worker, args=(r, 2, port, 8, 1 << 12, 16, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert ret.get(timeout=10) == 1


def _sc_worker(rank, world, port, n, ret):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import hobbit_b200
    from hobbit_b200.dist import GpuSumcheckBackend, sumcheck3_sharded
    ctx = hobbit_b200.Context(rank)
    rng = np.random.default_rng(3)
    v = [rand_field(rng, n) for _ in range(3)]
    pr = rand_field(rng, 1)
    nl = n // world
    dev = [torch.from_numpy(np.ascontiguousarray(x[rank * nl:(rank + 1) * nl]).view(np.int64)).cuda() for x in v]
    got = sumcheck3_sharded(GpuSumcheckBackend(ctx, torch.device("cuda", rank)), dev, nl, pr, ctx.mimc)
    want, _ = ctx.sumcheck3(v[0], v[1], v[2], pr)
    t = torch.tensor([1 if np.array_equal(got, want) else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        ret.put(int(t.item()))
    dist.destroy_process_group()
    ctx.close()


def test_sharded_sumcheck_nccl_2gpus():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2); the CPU/gloo twin is tests/test_dist_gloo.py")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    c = mp.get_context("spawn")
    ret = c.Queue()
    procs = [c.Process(target=_sc_worker, args=(r, 2, port, 1 << 16, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert ret.get(timeout=10) == 1
