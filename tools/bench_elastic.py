"""BASELINE config 5: Elastic_PC streaming commit + open (test_Elastic_PC(N, 2): Orion columns, tensor_row_size = BUFFER_SPACE / 2^14) on
1..8 GPUs, one process per GPU (torchrun), the stream sharded by groups of 4 chunks.

  python tools/bench_elastic.py --logn 28 --logb 20                                    # 1 GPU
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_elastic.py --logn 30 --logb 20

commit : every rank encodes its groups (RS rows, Orion columns, inner digests), all_to_all of inner digests by leaf range (pipelined with
         the encode), Merkle–Damgård chain + subtree per leaf range, all_gather of the subtrees        (hobbit_b200/dist.py)
open   : every rank runs ONE pass over its chunks (aggregate partial + query replies), the G partial aggregates are all_gathered and
         summed in the field, rank 0 runs the recursion (shockwave / WHIR / Spielman_stream) through the host mirror.
The synthetic stream "test" (witness_stream.cpp:2405-2411: every chunk is the same recurrence) is generated once on the host and kept in
HBM; --pinned streams the chunks of each group from pinned host memory instead (double buffered by the library's copy stream).
With --check (small sizes) every Merkle level is compared with the single-process C oracle.
"""
import argparse
import ctypes
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
from hobbit_b200.dist import GpuBackend, elastic_commit_sharded  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--logn", type=int, default=26)
    ap.add_argument("--logb", type=int, default=20)
    ap.add_argument("--lin", type=int, default=1)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--pinned", action="store_true", help="commit: stream this rank's chunks from pinned host memory (double buffered) instead of HBM")
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    N, B = 1 << a.logn, 1 << a.logb
    K = N // B
    ngroups = K // 4
    assert ngroups % world == 0 and ngroups >= world, "groups of 4 chunks must split evenly across ranks"
    trs = max(B >> 14, 16) if a.lin else max(B >> 11, 16)
    host = ctypes.CDLL(os.path.join(ROOT, "hobbit_b200", "libhobbit_host.so"))
    host.hobbit_c_backend.restype = ctypes.c_void_p
    host.hobbit_c_expander_init_store.restype = ctypes.c_longlong
    ctx = hobbit_b200.Context.from_handle(host.hobbit_c_backend(local))
    libc = ctypes.CDLL(None)
    libc.srand(1)                                                                   # every rank draws the same graphs and queries
    host.hobbit_c_set_globals(ctypes.c_size_t(B), trs, a.lin)
    if a.lin:
        host.hobbit_c_expander_init_store(ctypes.c_longlong(trs))
    be = GpuBackend(ctx, dev)
    # the stream: one chunk of the recurrence, replicated for this rank's groups, resident in HBM
    chunk = np.zeros((B, 2), dtype=np.uint64)
    ctx._ck(ctx.lib.hb_stream_pc_test(ctx.h, chunk.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(B)))
    gl = ngroups // world
    local_stream = torch.from_numpy(chunk.view(np.int64)).to(dev).repeat(4 * gl, 1)      # (4*gl*B, 2)
    torch.cuda.synchronize()
    host_stream = None
    if a.pinned:
        host_stream = torch.from_numpy(chunk.view(np.int64)).repeat(4 * gl, 1).pin_memory()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    t_commit, t_open, levels = [], [], None
    for rep in range(a.reps + 1):
        barrier(); t0 = time.perf_counter()
        levels = elastic_commit_sharded(be, host_stream.data_ptr() if a.pinned else local_stream.data_ptr(), ngroups, B, trs, a.lin)
        barrier(); t1 = time.perf_counter()
        # ---- open (Elastic_PC.cpp:625-726) ----
        x = np.zeros((a.logn, 2), dtype=np.uint64)
        host.hobbit_c_generate_randomness(a.logn, x.ctypes.data_as(ctypes.c_void_p))
        beta = ctx.precompute_beta(x[:int(np.log2(K))])
        rv = np.zeros((1, 2), dtype=np.uint64)
        host.hobbit_c_generate_randomness(1, rv.ctypes.data_as(ctypes.c_void_p))
        queries = 5900 if a.lin else 700
        col, row = np.zeros(queries, dtype=np.uint32), np.zeros(queries, dtype=np.uint32)
        host.hobbit_c_elastic_draw_queries(queries, col.ctypes.data_as(ctypes.c_void_p), row.ctypes.data_as(ctypes.c_void_p))
        kl = 4 * gl
        agg = torch.zeros((B, 2), dtype=torch.int64, device=dev)
        reply = torch.zeros((queries * kl, 2), dtype=torch.int64, device=dev)
        ctx._ck(ctx.lib.hb_elastic_open_begin(ctx.h, ctypes.c_size_t(B), trs, a.lin, col.ctypes.data_as(ctypes.c_void_p), row.ctypes.data_as(ctypes.c_void_p),
                                              ctypes.c_size_t(queries), ctypes.c_size_t(kl)))
        for i in range(kl):
            b = np.ascontiguousarray(beta[rank * kl + i:rank * kl + i + 1])
            ctx._ck(ctx.lib.hb_elastic_open_push(ctx.h, ctypes.c_void_p(local_stream.data_ptr() + i * B * 16), b.ctypes.data_as(ctypes.c_void_p)))
        ctx._ck(ctx.lib.hb_elastic_open_finish(ctx.h, ctypes.c_void_p(agg.data_ptr()), ctypes.c_void_p(reply.data_ptr())))
        if world > 1:
            parts = [torch.empty_like(agg) for _ in range(world)]
            dist.all_gather(parts, agg)
            torch.cuda.synchronize()
            for h in range(world):
                if h != rank:                                                       # field addition of the partial aggregates
                    ctx._ck(ctx.lib.hb_field_binop(ctx.h, 0, ctypes.c_void_p(agg.data_ptr()), ctypes.c_void_p(parts[h].data_ptr()),
                                                   ctypes.c_void_p(agg.data_ptr()), ctypes.c_size_t(B)))
        ps = ctypes.c_double(0.0)
        if rank == 0:
            fd = os.dup(1); dn = os.open(os.devnull, os.O_WRONLY); os.dup2(dn, 1)   # the reference-style printf chatter of the recursion
            try:
                host.hobbit_c_elastic_open_tail(ctypes.c_void_p(agg.data_ptr()), ctypes.c_size_t(K), ctypes.c_size_t(4 * B), int(np.log2(4 * B)) + 1,
                                                ctypes.byref(ps))
            finally:
                os.dup2(fd, 1); os.close(dn); os.close(fd)
        barrier(); t2 = time.perf_counter()
        if rep:
            t_commit.append(t1 - t0); t_open.append(t2 - t1)
    ok = None
    if a.check and rank == 0:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from helpers import Checker
        orc = Checker("orc")
        if a.lin:
            libc.srand(1); orc.expander_init_store(trs)
        want = np.zeros((8 * B - 1, 32), dtype=np.uint8)
        orc.fn("elastic_commit")(ctypes.c_size_t(N), ctypes.c_size_t(B), trs, a.lin, want.ctypes.data_as(ctypes.c_void_p))
        ok = bool(np.array_equal(levels.cpu().numpy(), want))
    if rank == 0:
        c, o = min(t_commit), min(t_open)
        print(json.dumps({"workload": "test_Elastic_PC(2^%d, %s), BUFFER_SPACE 2^%d, tensor_row_size %d" % (a.logn, "Orion columns" if a.lin else "RS columns", a.logb, trs),
                          "n_gpus": world, "commit_s": round(c, 5), "open_s": round(o, 5), "commit_field_elems_per_s": round(N / c, 1),
                          "commit_open_field_elems_per_s": round(N / (c + o), 1), "open_ps_kb": ps.value, "stream": ("commit from pinned host memory, double buffered; open from HBM" if a.pinned else "resident in HBM (one chunk replicated)"),
                          "levels_equal_oracle": ok}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
