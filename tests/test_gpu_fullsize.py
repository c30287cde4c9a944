"""BASELINE config 2 at FULL size (commit_standard, N = 2^26, K = 32, tensor_row_size 1024 — the bench.py workload) through
size-independent properties, each anchored on the C oracle:
  * the leaves are the Merkle–Damgård chain over the per-chunk inner digests (sharding split == monolithic commit), sampled positions
    recomputed with the oracle's BLAKE3;
  * inner digests of one whole chunk-column are recomputed from scratch by the oracle (row NTTs, expander encode of the column, H1);
  * every tree level obeys parent = H1(left | left) on sampled nodes;
  * the commit is deterministic (two runs, identical levels)."""
import numpy as np
import pytest

from helpers import Checker, srand

pytestmark = pytest.mark.gpu


def test_commit_standard_2e26_properties():
    import torch
    import hobbit_b200
    N, K, trs = 1 << 26, 32, 1024
    B, cols = N // K, 2 * (N // K) // trs
    orc = Checker("orc")
    ctx = hobbit_b200.Context(0)
    srand(1)
    poly = orc.generate_randomness(N)
    cw = orc.expander_init_store(trs)
    assert ctx.expander_set(trs, orc.expander_graphs(trs)) == cw == 1761
    lv, _ = ctx.commit_standard(poly, K, trs, 1)
    lv2, _ = ctx.commit_standard(poly, K, trs, 1)
    assert np.array_equal(lv, lv2)                                           # deterministic
    # sharding split on the same data: inner digests per chunk (kept on the device: 2 GiB), chained, == the leaves
    inner = torch.empty((K, B, 32), dtype=torch.uint8, device="cuda")
    dpoly = torch.from_numpy(poly.view(np.int64)).cuda()
    ctx.commit_encode_chunks(dpoly.data_ptr(), K, B, trs, 1, inner_out=inner.data_ptr())
    rng = np.random.default_rng(26)
    pos = rng.integers(0, B, size=12)
    samp = inner[:, torch.from_numpy(pos).cuda()].cpu().numpy()              # (K, 12, 32)
    for q, p in enumerate(pos):
        leaf = np.zeros(32, dtype=np.uint8)
        for c in range(K):
            leaf = orc.blake3(np.concatenate([samp[c, q], leaf]))
        assert np.array_equal(leaf, lv[p]), p
    # one chunk-column recomputed from scratch by the oracle
    c, k = 17, int(rng.integers(0, cols))
    rows = np.zeros((trs, cols, 2), dtype=np.uint64)
    rows[:, :cols // 2] = poly[c * B:(c + 1) * B].reshape(trs, cols // 2, 2)
    col = np.stack([orc.fft(rows[r], int(np.log2(cols)))[k] for r in range(trs)])
    code, _ = orc.encode(col, trs)
    col_inner = inner[c].cpu().numpy().reshape(trs // 2, cols, 32)[:, k]     # leaf order j*cols + k
    for j in rng.integers(0, trs // 2, size=8):
        cells = code[4 * j:4 * j + 4]
        assert np.array_equal(orc.blake3(cells.view(np.uint8).reshape(64)), col_inner[j]), j
    # tree levels: parent = H1(left | left) (the reference's rule, merkle_tree.cpp:275-280)
    off, n = 0, B
    while n > 1:
        for i in rng.integers(0, n // 2, size=4):
            left = lv[off + 2 * i]
            assert np.array_equal(orc.blake3(np.concatenate([left, left])), lv[off + n + i])
        off += n
        n //= 2
    ctx.close()
