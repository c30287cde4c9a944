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

    def encode_chunks(self, poly, nchunks, B, trs, lin):
        """poly: int device pointer / numpy host array of nchunks*B elements."""
        inner = self.empty(nchunks, B, 32)
        self.ctx.commit_encode_chunks(poly, nchunks, B, trs, lin, inner_out=inner.data_ptr())
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
        return lv


def commit_standard_sharded(backend, poly_local, K, B, trs, lin, group=None):
    """poly_local: this rank's K/G consecutive chunks (chunk range [rank*K/G, (rank+1)*K/G)).
    Returns every Merkle level of the commitment as one (2B-1, 32) uint8 tensor on every rank (== MT_hashes, leaves first)."""
    G = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    assert K % G == 0 and B % G == 0, "chunks and leaves must split evenly across ranks"
    kl, Bp = K // G, B // G
    inner = backend.encode_chunks(poly_local, kl, B, trs, lin)                       # [kl, B, 32]
    if G > 1:
        # to rank h: my chunks, h's leaf range.  [kl, G, Bp, 32] -> [G, kl, Bp, 32]
        send = inner.view(kl, G, Bp, 32).permute(1, 0, 2, 3).contiguous()
        recv = torch.empty_like(send)                                                # [G(src rank), kl, Bp, 32] == global chunk order
        dist.all_to_all_single(recv, send, group=group)
        inner_all = recv.view(K, Bp, 32)
    else:
        inner_all = inner
    leaves = backend.chain(inner_all, backend.zeros(Bp, 32))                         # chain starts from all-zero digests
    sub = backend.tree(leaves)                                                       # [(2Bp-1), 32]
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
    return torch.cat(out, dim=0)
