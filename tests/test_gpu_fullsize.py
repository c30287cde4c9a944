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


def test_elastic_commit_2e26_properties():
    """BASELINE config 5 shape (test_Elastic_PC option 2: Orion columns, BUFFER_SPACE 2^20, tensor_row_size 64) at 2^26 coefficients:
    streaming commit == sharding split (groups of 4 chunks), sampled leaves recomputed from inner digests with the oracle's BLAKE3, one
    chunk-column of one group recomputed from scratch by the oracle, tree rule on sampled nodes."""
    import torch
    import hobbit_b200
    from hobbit_b200.dist import GpuBackend, elastic_commit_sharded
    N, B, trs = 1 << 26, 1 << 20, 64
    K, cols = N // B, 2 * B // trs
    orc = Checker("orc")
    ctx = hobbit_b200.Context(0)
    srand(1)
    cw = orc.expander_init_store(trs)
    assert ctx.expander_set(trs, orc.expander_graphs(trs)) == cw
    rng = np.random.default_rng(5)
    stream = rng.integers(0, 1 << 61, size=(N, 2), dtype=np.uint64) % np.uint64((1 << 61) - 1)
    lv = ctx.elastic_commit([stream[i * B:(i + 1) * B] for i in range(K)], B, trs, 1)
    dstream = torch.from_numpy(stream.view(np.int64)).cuda()
    got = elastic_commit_sharded(GpuBackend(ctx, torch.device("cuda", 0)), dstream.data_ptr(), K // 4, B, trs, 1)
    assert np.array_equal(got.cpu().numpy(), lv)
    # inner digests of group 7, chained into the sampled leaves by the oracle
    ng = K // 4
    inner = torch.empty((ng, 4 * B, 32), dtype=torch.uint8, device="cuda")
    ctx.elastic_encode_groups(dstream.data_ptr(), ng, B, trs, 1, inner.data_ptr())
    pos = rng.integers(0, 4 * B - 1, size=8)
    samp = inner[:, torch.from_numpy(pos).cuda()].cpu().numpy()
    for q, p in enumerate(pos):
        leaf = np.zeros(32, dtype=np.uint8)
        for g in range(ng):
            leaf = orc.blake3(np.concatenate([samp[g, q], leaf]))
        assert np.array_equal(leaf, lv[p]), p
    # group 7, column k: the four chunks' tensors at column k from the oracle -> inner digests of the positions (row, k), (row, k-1)
    g, k = 7, int(rng.integers(1, cols - 1))
    colcode = []
    for c in range(4):
        chunk = stream[(4 * g + c) * B:(4 * g + c + 1) * B].reshape(trs, cols // 2, 2)
        rows = np.zeros((trs, cols, 2), dtype=np.uint64)
        rows[:, :cols // 2] = chunk
        ff = [orc.fft(rows[r], int(np.log2(cols))) for r in range(trs)]
        colcode.append([orc.encode(np.stack([f[kk] for f in ff]), trs)[0] for kk in (k, k + 1)])       # columns k and k+1 (2*trs entries each)
    gi = inner[g].cpu().numpy()
    for row in rng.integers(0, 2 * trs, size=6):
        p = int(row) * cols + k                                                # tuple hashed at p: (c0[p+1], c1[p+1], c2[p], c3[p])
        cells = np.stack([colcode[0][1][row], colcode[1][1][row], colcode[2][0][row], colcode[3][0][row]])
        assert np.array_equal(orc.blake3(cells.view(np.uint8).reshape(64)), gi[p]), (row, k)
    off, n = 0, 4 * B
    while n > 1:
        for i in rng.integers(0, n // 2, size=3):
            left = lv[off + 2 * i]
            assert np.array_equal(orc.blake3(np.concatenate([left, left])), lv[off + n + i])
        off += n
        n //= 2
    ctx.close()


def _level_sha256(levels, nleaves):
    import hashlib
    out, off, n = [], 0, nleaves
    while n >= 1:
        out.append(hashlib.sha256(np.ascontiguousarray(levels[off:off + n]).tobytes()).hexdigest())
        off += n
        n //= 2
    return out


def _golden():
    import json
    import os
    return json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fullsize.json")))


def test_commit_standard_2e26_equals_reference_digests():
    """BASELINE config 2 pinned to the UNMODIFIED reference at full size: tests/golden/fullsize.json holds the SHA-256 of every Merkle level
    the reference's commit_standard produced for this input (tests/golden/make_golden_fullsize.py, 233 s on one host core).  Fails if any
    of the 2^22 - 1 digests differs."""
    import hobbit_b200
    g = _golden()["commit_standard_2e26"]
    N, K, trs = g["N"], g["K"], g["trs"]
    orc = Checker("orc")
    ctx = hobbit_b200.Context(0)
    srand(1)
    poly = orc.generate_randomness(N)
    orc.expander_init_store(trs)
    ctx.expander_set(trs, orc.expander_graphs(trs))
    lv, _ = ctx.commit_standard(poly, K, trs, 1)
    assert lv[-1].tobytes().hex() == g["root"]
    assert _level_sha256(lv, N // K) == g["levels_sha256"]
    ctx.close()


def test_elastic_commit_2e26_equals_reference_digests():
    """BASELINE config 5 shape pinned to the unmodified reference: Elastic_PC commit of its synthetic test stream, N = 2^26, BUFFER_SPACE 2^20,
    Orion columns — every level's SHA-256 (the last leaf is excluded: the reference reads past its buffers there)."""
    import hashlib
    import hobbit_b200
    g = _golden()["elastic_commit_2e26"]
    N, B, trs = g["N"], g["B"], g["trs"]
    orc = Checker("orc")
    ctx = hobbit_b200.Context(0)
    srand(1)
    orc.expander_init_store(trs)
    ctx.expander_set(trs, orc.expander_graphs(trs))
    chunk = ctx.stream_pc_test(B)
    lv = ctx.elastic_commit([chunk] * (N // B), B, trs, 1)
    got = _level_sha256(lv, 4 * B)
    ctx.close()
    assert lv[-1].tobytes().hex() == g["root"]
    assert got[1:] == g["levels_sha256"][1:]                       # every level above the leaves
    # the leaves: all but the very last one (the reference computes it from reads past its buffers — undefined behaviour; it is a right
    # child and, with the left||left parent rule, influences nothing)
    assert hashlib.sha256(np.ascontiguousarray(lv[:4 * B - 1]).tobytes()).hexdigest() == g["leaves_sha256_without_last"]
