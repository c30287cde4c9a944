"""Multi-GPU sharding of the commitment, one process per GPU over torch.distributed.

commit_standard over K chunks is chunk-parallel except for the Merkle–Damgård chain of every leaf position over the chunks.
Sharding: rank g owns the contiguous chunk range [g*K/G, (g+1)*K/G) for the heavy, chunk-independent part (RS rows, expander
columns, inner digests) and the contiguous LEAF range [g*B/G, (g+1)*B/G) for the chain and its Merkle subtree:

    1. every rank encodes its chunks                       -> inner[c_local][p]        (32 B per coefficient)
    2. all_to_all by leaf range                            -> inner[c_global][p_local]  (the one real exchange step; NCCL/NVLink)
    3. every rank chains its leaf range over ALL chunks in order and builds its subtree
    4. all_gather of the subtree levels; the top log2(G) levels are rebuilt from the G subtree roots on every rank

The compute steps go through a small backend object so the same orchestration runs on CPU tensors with gloo in the tests
(tests/test_dist_gloo.py, backend = the C oracle) and on CUDA tensors with NCCL in production (GpuBackend -> the C ABI).
"""
import numpy as np
import torch
import torch.distributed as dist


class GpuBackend:
    """Device tensors in, device tensors out; all arithmetic in libhobbit_b200.so."""

    def __init__(self, ctx, device):
        self.ctx, self.device = ctx, device

    def empty(self, *shape):
        return torch.empty(shape, dtype=torch.uint8, device=self.device)

    def zeros(self, *shape):
        return torch.zeros(shape, dtype=torch.uint8, device=self.device)

    def sync(self):
        torch.cuda.synchronize()

    def encode_chunks(self, poly, nchunks, B, trs, lin, first=0, parts=1, total=0):
        """poly: int device pointer / numpy host array holding this rank's chunks; encodes chunks [first, first+nchunks).
        Returns the inner digests in the exchange layout [parts][nchunks][B/parts][32] (written so by the kernel itself)."""
        inner = self.empty(parts, nchunks, B // parts, 32)
        if isinstance(poly, int):
            src = poly + first * B * 16
        else:
            src = poly[first * B:(first + nchunks) * B]
        self.ctx.commit_encode_chunks(src, nchunks, B, trs, lin, inner_out=inner.data_ptr(), leaf_parts=parts, first_chunk=first, total_chunks=total)
        return inner

    def encode_groups(self, chunks, ngroups, B, trs, lin, first=0, parts=1, total=0):
        """Elastic commit: groups [first, first+ngroups) of 4 chunks each -> inner digests [parts][ngroups][4B/parts][32]."""
        inner = self.empty(parts, ngroups, 4 * B // parts, 32)
        src = chunks + first * 4 * B * 16 if isinstance(chunks, int) else chunks[first * 4 * B:(first + ngroups) * 4 * B]
        self.ctx.elastic_encode_groups(src, ngroups, B, trs, lin, inner.data_ptr(), parts, first, total)
        return inner

    # torch (and NCCL) work is ordered on torch's streams, the library's on its own stream: every hand-over is a full
    # device synchronisation (each C-ABI call also synchronises its stream before returning).
    def chain(self, inner, leaves):
        torch.cuda.synchronize()
        self.ctx.md_chain(inner.data_ptr(), inner.shape[0], inner.shape[1], leaves.data_ptr())
        return leaves

    def tree(self, leaves):
        n = leaves.shape[0]
        lv = self.empty(2 * n - 1, 32)
        lv[:n] = leaves
        torch.cuda.synchronize()
        self.ctx.merkle_tree_inplace(lv.data_ptr(), n)
        self.ctx.sync()                  # device-pointer calls are stream-ordered on the library's stream: hand back to torch's
        return lv


def elastic_commit_sharded(backend, chunks_local, ngroups, B, trs, lin, group=None, groups=4, timing=None):
    """Elastic_PC commit (Elastic_PC.cpp:174-285) sharded the same way: the unit is a GROUP of 4 consecutive chunks of B coefficients
    (leaf[p] <- H1( H1(c0[p+1] | c1[p+1] | c2[p] | c3[p]) | leaf[p] ) over the groups in order, 4B leaf positions).
    chunks_local: this rank's ngroups/G consecutive groups (4*B*ngroups/G coefficients).  Returns all (8B-1, 32) levels on every rank."""
    return commit_standard_sharded(backend, chunks_local, ngroups, 4 * B, trs, lin, group=group, groups=groups, timing=timing, elastic_B=B)


def commit_standard_sharded(backend, poly_local, K, B, trs, lin, group=None, groups=4, timing=None, elastic_B=None):
    """poly_local: this rank's K/G consecutive chunks (chunk range [rank*K/G, (rank+1)*K/G)); numpy host array, int device
    pointer, or anything backend.encode_chunks accepts together with an element offset.
    Returns every Merkle level of the commitment as one (2B-1, 32) uint8 tensor on every rank (== MT_hashes, leaves first).

    The local chunks are encoded in `groups` pieces; the all_to_all of piece i runs (async, NCCL stream) while piece i+1 is being
    encoded, so the exchange hides behind the encode except for the last piece."""
    import time
    import os
    G = dist.get_world_size(group) if dist.is_initialized() else 1
    groups = int(os.environ.get("HB_SHARD_PIECES", groups))      # experiment switch: pieces of the local encode that the exchange is pipelined behind
    assert K % G == 0 and B % G == 0, "chunks and leaves must split evenly across ranks"
    kl, Bp = K // G, B // G
    if elastic_B is None:
        encode = lambda first, n, parts, total: backend.encode_chunks(poly_local, n, B, trs, lin, first, parts=parts, total=total)
    else:                                                    # units are groups of 4 chunks, B is the number of leaf positions (4 * BUFFER_SPACE)
        encode = lambda first, n, parts, total: backend.encode_groups(poly_local, n, elastic_B, trs, lin, first, parts=parts, total=total)
    t0 = time.perf_counter()
    if G == 1:
        inner_all = encode(0, kl, 1, 0).view(kl, B, 32)
    else:
        groups = max(1, min(groups, kl))
        while kl % groups:
            groups -= 1
        kg = kl // groups
        recv = backend.empty(G, kl, Bp, 32)                      # [source rank][its local chunk][my leaf range] == global chunk order
        works = []
        for g in range(groups):
            send = encode(g * kg, kg, G, kl)                                          # [G, kg, Bp, 32]: slice h -> rank h
            if dist.get_backend(group) == "nccl":
                # receive straight into the final [source rank][chunk] slots: no staging copy
                outs = [recv[h, g * kg:(g + 1) * kg] for h in range(G)]
                works.append((dist.all_to_all(outs, list(send.unbind(0)), group=group, async_op=True), send, None, g))
            else:                                                                     # gloo (CPU tests): single-buffer form + copy
                rbuf = backend.empty(G, kg, Bp, 32)
                works.append((dist.all_to_all_single(rbuf, send, group=group, async_op=True), send, rbuf, g))
        for w, _send, rbuf, g in works:
            w.wait()
            if rbuf is not None:
                recv[:, g * kg:(g + 1) * kg] = rbuf
        inner_all = recv.view(K, Bp, 32)
    t1 = time.perf_counter()
    leaves = backend.chain(inner_all, backend.zeros(Bp, 32))                         # chain starts from all-zero digests
    sub = backend.tree(leaves)                                                       # [(2Bp-1), 32]
    t2 = time.perf_counter()
    if G == 1:
        return sub
    allsub = [torch.empty_like(sub) for _ in range(G)]
    dist.all_gather(allsub, sub, group=group)
    out, n, off = [], Bp, 0
    while n >= 1:                                                                    # levels below the subtree roots: concatenate by rank
        out.append(torch.cat([s[off:off + n] for s in allsub], dim=0))
        off += n
        n //= 2
    roots = out[-1]                                                                  # G subtree roots == level log2(Bp) of the global tree
    top = backend.tree(roots)                                                        # top log2(G) levels (every rank, G-1 compressions)
    out.append(top[G:])
    res = torch.cat(out, dim=0)
    if timing is not None:
        backend.sync()
        timing.append((t1 - t0, t2 - t1, time.perf_counter() - t2))
    return res


# ---------------------------------------------------------------------------------------------------------------------------------
# Sumcheck sharded by hypercube prefix: rank g holds the contiguous slice [g*n/G, (g+1)*n/G) of every table.  Adjacent-pair
# folding keeps pairs rank-local for the first log2(n/G) rounds; per round the only communication is the all-reduce of the 4
# round-polynomial coefficients (64 B; summed as uint64 limbs — each limb < 2^61 and G <= 8, so the sum cannot overflow — then reduced
# mod p), after which every rank derives the same MiMC challenge.  Once a slice is down to one entry the G values per table are
# all-gathered and the last log2 G rounds run redundantly on every rank.
P61 = (1 << 61) - 1


def _mod_p(a):
    a = (a & np.uint64(P61)) + (a >> np.uint64(61))
    return np.where(a >= np.uint64(P61), a - np.uint64(P61), a)


def _fadd(a, b):
    s = a + b
    return np.where(s >= np.uint64(P61), s - np.uint64(P61), s)


def sumcheck3_sharded(backend, tables, n_local, prev_r, mimc, group=None):
    """_generate_3product_sumcheck_proof (sumcheck.cpp:1974-2058) over tables of n = n_local * G entries.
    tables: backend-specific handles of this rank's 3 slices.  mimc(x, k) -> (1,2) uint64.
    Returns the flat proof (5*rounds + 4, 2) uint64, identical on every rank and identical to the single-device prover."""
    G = dist.get_world_size(group) if dist.is_initialized() else 1
    rounds_local = int(np.log2(n_local))
    rounds = rounds_local + int(np.log2(G))
    proof = np.zeros((5 * rounds + 4, 2), dtype=np.uint64)
    rand = np.ascontiguousarray(prev_r, dtype=np.uint64).reshape(1, 2)
    cur = tables
    for i in range(rounds_local):
        L = n_local >> (i + 1)
        co, cur = backend.sc3_round(cur, L, rand)                                    # (4,2) uint64 partial coefficients
        if G > 1:
            t = backend.to_comm(co)
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)                    # limb-wise uint64 sums (< 8 * 2^61)
            co = _mod_p(backend.from_comm(t))
        proof[4 * rounds + i] = rand
        for c in range(4):
            proof[4 * i + c] = co[c]
            rand = mimc(rand, co[c:c + 1])
    heads = backend.heads(cur)                                                       # (3,2): the single remaining entry of each slice
    if G > 1:
        t = backend.to_comm(heads)
        allh = [torch.empty_like(t) for _ in range(G)]
        dist.all_gather(allh, t, group=group)
        small = np.stack([backend.from_comm(x) for x in allh], axis=1)               # (3, G, 2): rank order == index order
        cur = backend.tables_from(small)                                             # G entries per table, on every rank
        for i in range(int(np.log2(G))):
            co, cur = backend.sc3_round(cur, G >> (i + 1), rand)
            r_i = rounds_local + i
            proof[4 * rounds + r_i] = rand
            for c in range(4):
                proof[4 * r_i + c] = co[c]
                rand = mimc(rand, co[c:c + 1])
        heads = backend.heads(cur)
    rand = mimc(rand, heads[0:1]); rand = mimc(rand, heads[1:2])
    proof[5 * rounds:5 * rounds + 3] = heads
    proof[5 * rounds + 3] = rand
    return proof


class GpuSumcheckBackend:
    """Round kernel = hb_sc3_round on device slices (ping-pong scratch); NCCL moves 64 bytes per round."""

    def __init__(self, ctx, device):
        self.ctx, self.device = ctx, device
        self.scratch = None

    def sc3_round(self, cur, L, rand):
        if self.scratch is None:
            n = cur[0].shape[0]
            self.scratch = [[torch.empty((max(n // 2, 1), 2), dtype=torch.int64, device=self.device) for _ in range(3)],
                            [torch.empty((max(n // 4, 1), 2), dtype=torch.int64, device=self.device) for _ in range(3)]]
            self.flip = 0
        out = self.scratch[self.flip]
        self.flip ^= 1
        torch.cuda.synchronize()
        co = self.ctx.sc3_round([t.data_ptr() for t in cur], [t.data_ptr() for t in out], L, rand)
        return co, [t[:L] for t in out]

    def heads(self, cur):
        torch.cuda.synchronize()
        return np.stack([t[0].cpu().numpy().view(np.uint64) for t in cur])

    def tables_from(self, small):
        self.scratch = None                                                          # new (tiny) ping-pong buffers
        return [torch.from_numpy(np.ascontiguousarray(small[k]).view(np.int64)).to(self.device) for k in range(3)]

    def to_comm(self, a):
        return torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).to(self.device)

    def from_comm(self, t):
        return t.cpu().numpy().view(np.uint64)
