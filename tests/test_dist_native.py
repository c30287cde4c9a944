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
    # shapes that do not split over the ranks are refused loudly (no silent fallback to one GPU)
    import hobbit_b200
    try:
        ctx.dist_commit_standard(np.ascontiguousarray(poly[:3 * B]), 3 * world + 1, B, trs, 1)
        ok = False
    except hobbit_b200.HobbitError as e:
        ok = ok and "split evenly" in str(e)
    ctx.dist_barrier()
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
    try:
        for p in procs:
            p.join(timeout)
            assert p.exitcode == 0, "rank process failed or timed out (exitcode %r)" % (p.exitcode,)
        assert ret.get(timeout=10) == 1
    finally:
        for p in procs:                                                    # never leave a rank behind (it would spin in a device barrier)
            if p.is_alive():
                p.kill()


@pytest.mark.parametrize("world", [2, 4])
def test_native_sharded_commits(world):
    _run(_commit_worker, world)


def _prover_worker(rank, world, port, ret):
    """Every sharded prover against its own single-GPU run on the same (replicated) tables: proofs must be bit-identical."""
    from helpers import RawABI, consistent_trace, F
    ctx, dist = _setup(rank, world, port, 0, 1 << 20)
    raw = RawABI.__new__(RawABI); raw.lib = ctx.lib; raw.ctx = ctx.h
    orc = Checker("orc")
    rng = np.random.default_rng(77)
    cases = []
    for n in (1 << 12, 16, 2 * world, world):                                       # down to tables smaller than the slicing threshold
        v = [rand_field(rng, n) for _ in range(3)]
        cases.append(("sumcheck3 n=%d" % n, lambda v=v: ctx.sumcheck3(v[0], v[1], v[2], F([5, 7]))))
    for sizes in ([1 << 10, 1 << 10], [1 << 11, 1 << 6, 4, 1], [8 * world]):
        t = [rand_field(rng, sum(sizes)) for _ in range(3)]
        a = rand_field(rng, len(sizes))
        cases.append(("batch_sumcheck3 %s" % sizes, lambda t=t, a=a, sizes=sizes: ctx.batch_sumcheck3(t[0], t[1], t[2], sizes, a)))
    for vectors, n in ((1, 1 << 10), (4, 1 << 9)):
        x = rand_field(rng, vectors * n)
        xr = rand_field(rng, max(1, int(np.log2(vectors))))
        cases.append(("mul_tree %dx%d" % (vectors, n), lambda x=x, vectors=vectors, xr=xr: ctx.mul_tree(x, vectors, F([3, 0]), xr)))
    for total, B, layer in ((1 << 14, 1 << 8, 0), (1 << 16, 1 << 10, 3)):
        xy = rand_field(rng, total)
        r = rand_field(rng, int(np.log2((total >> layer) // 2)))
        rnd = rand_field(rng, 4)
        cases.append(("stream_layer %d/%d/%d" % (total, B, layer), lambda xy=xy, B=B, layer=layer, r=r, rnd=rnd: ctx.stream_layer(xy, B, layer, r, F([5, 0]), rnd)))
    for total, vectors, B, distance in ((1 << 15, 8, 1 << 10, 5), (1 << 16, 8, 1 << 8, 3)):      # shallow, and layers > distance (batched passes)
        xy = rand_field(rng, total)
        layers = int(np.log2(total // (2 * B)))
        if layers % distance and layers > distance:
            layers = distance + layers - layers % distance
        nr = 4 * layers if layers <= distance else (layers - distance) + distance * (3 * (layers // distance) + 1)
        rnd, pr, xr = rand_field(rng, nr), rand_field(rng, 1), rand_field(rng, int(np.log2(vectors)))
        cases.append(("mul_tree_stream %d/%d/%d/%d" % (total, vectors, B, distance),
                      lambda xy=xy, vectors=vectors, B=B, distance=distance, pr=pr, xr=xr, rnd=rnd: ctx.mul_tree_stream(xy, vectors, B, distance, 0, pr, xr, rnd)))
    for cs, B in ((1 << 12, 1 << 9), (1 << 10, 1 << 10)):
        L, R, O, S = consistent_trace(np.random.default_rng(cs), orc, cs)
        r, rnd = rand_field(rng, int(np.log2(B))), rand_field(rng, 10)
        cases.append(("gate_stream %d/%d" % (cs, B), lambda L=L, R=R, O=O, S=S, B=B, r=r, rnd=rnd: ctx.gate_consistency_stream(L, R, O, S, B, r, rnd)))
    for cs, B in ((1 << 12, 1 << 8), (1 << 12, 1 << 12)):
        r2 = np.random.default_rng(cs + B)
        L, R = rand_field(r2, cs), rand_field(r2, cs)
        S = np.zeros((cs, 2), dtype=np.uint64); S[:, 0] = r2.integers(0, 3, size=cs)
        O = rand_field(r2, cs)
        add, mul = S[:, 0] == 0, S[:, 0] == 1
        O[add] = orc.binop(0, L, R)[add]
        O[mul] = orc.binop(2, L, R)[mul]
        r, lr, rnd = rand_field(r2, int(np.log2(B))), rand_field(r2, 2), rand_field(r2, 13)
        cases.append(("gate_lookups %d/%d" % (cs, B), lambda L=L, R=R, O=O, S=S, B=B, r=r, lr=lr, rnd=rnd: raw.gate_consistency_lookups(L, R, O, S, B, r, lr, rnd)))

    def same(a, b):
        if isinstance(a, (tuple, list)):
            return len(a) == len(b) and all(same(x, y) for x, y in zip(a, b))
        if isinstance(a, np.ndarray):
            return np.array_equal(a, b)
        return a == b

    ok = True
    for name, fn in cases:
        ctx.dist_shard(False)
        want = fn()
        ctx.dist_shard(True)
        before = ctx.dist_stats()["reductions"]
        got = fn()
        sharded = ctx.dist_stats()["reductions"] > before
        if not sharded and name != "sumcheck3 n=%d" % world:                          # a table of `world` entries cannot be sliced
            print("rank %d: %s did not run sharded" % (rank, name), flush=True)
            ok = False
        if not same(want, got):
            print("rank %d: MISMATCH in %s" % (rank, name), flush=True)
            ok = False
    ctx.dist_shard(False)
    _finish(ctx, dist, ok, rank, ret)


@pytest.mark.parametrize("world", [2, 4])
def test_native_sharded_provers(world):
    _run(_prover_worker, world)


def _fullsize_worker(rank, world, port, ret):
    """BASELINE config 2 at full size, sharded over `world` ranks, against the SHA-256 level digests of the unmodified reference."""
    import hashlib
    import json
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fullsize.json")))["commit_standard_2e26"]
    N, K, trs = g["N"], g["K"], g["trs"]
    B = N // K
    import torch
    import torch.distributed as dist
    import hobbit_b200
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dev = rank if torch.cuda.device_count() >= world else 0
    torch.cuda.set_device(dev)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ctx = hobbit_b200.Context(dev)
    orc = Checker("orc")
    srand(1)
    poly = orc.generate_randomness(N)
    orc.expander_init_store(trs)
    ctx.expander_set(trs, orc.expander_graphs(trs))
    ctx.dist_init_torch(K * (B // world) * 32 + 2 * B * 32 + 4096)
    kl = K // world
    lv = ctx.dist_commit_standard(np.ascontiguousarray(poly[rank * kl * B:(rank + 1) * kl * B]), K, B, trs, 1)
    out, off, n = [], 0, B
    while n >= 1:
        out.append(hashlib.sha256(np.ascontiguousarray(lv[off:off + n]).tobytes()).hexdigest())
        off += n
        n //= 2
    _finish(ctx, dist, out == g["levels_sha256"], rank, ret)


def test_native_sharded_commit_2e26_equals_reference_digests():
    _run(_fullsize_worker, 2, timeout=900)
