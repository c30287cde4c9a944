"""8f.1 building blocks (csrc/open.cu) against their plain CPU restatements (oracle/hb_emul.cpp), same calls through the C ABI.
Bit-exact: field elements are canonical, digests are bytes."""
import ctypes

import numpy as np
import pytest

from helpers import RawABI, rand_field

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def abis():
    return RawABI("gpu"), RawABI("emul")


def test_rs_encode_rows(abis):
    g, e = abis
    rng = np.random.default_rng(1)
    for in_len, rows, logn in [(8, 5, 4), (256, 32, 9), (100, 3, 7), (2048, 4, 12), (8192, 2, 14), (16, 7, 4)]:
        src = rand_field(rng, in_len * rows)
        assert np.array_equal(g.rs_encode_rows(src, in_len, rows, logn), e.rs_encode_rows(src, in_len, rows, logn)), (in_len, rows, logn)


def test_matvec(abis):
    g, e = abis
    rng = np.random.default_rng(2)
    for rows, cols in [(32, 512), (1, 1000), (2048, 64), (33, 4097), (16, 32), (512, 4096)]:
        M, w, s = rand_field(rng, rows * cols), rand_field(rng, rows), rand_field(rng, cols)
        assert np.array_equal(g.matvec_cols(M, rows, cols, w), e.matvec_cols(M, rows, cols, w)), ("cols", rows, cols)
        assert np.array_equal(g.matvec_rows(M, rows, cols, s), e.matvec_rows(M, rows, cols, s)), ("rows", rows, cols)


def test_axpy_scatter_gather(abis):
    g, e = abis
    rng = np.random.default_rng(3)
    y, x, a = rand_field(rng, 5000), rand_field(rng, 5000), rand_field(rng, 1)
    assert np.array_equal(g.axpy(y, x, a), e.axpy(y, x, a))
    idx = rng.permutation(1 << 14)[:3000]
    val = rand_field(rng, 3000)
    assert np.array_equal(g.scatter(1 << 14, idx, val), e.scatter(1 << 14, idx, val))
    assert np.array_equal(g.scatter(64, np.zeros(0, dtype=np.uint64), rand_field(rng, 0)), np.zeros((64, 2), dtype=np.uint64))
    M = rand_field(rng, 32 * 1024)
    col = rng.integers(0, 1024, size=240)
    assert np.array_equal(g.gather_cols(M, 32, 1024, col), e.gather_cols(M, 32, 1024, col))


def test_phi_g_init(abis):
    g, e = abis
    rng = np.random.default_rng(4)
    for n in [1, 2, 5, 11, 12, 13, 16]:
        r = rand_field(rng, n)
        a, b = g.phi_g_init(r), e.phi_g_init(r)
        assert np.array_equal(a, b), n
        assert not a[len(a) // 2:].any()


def test_shockwave_leaves(abis):
    g, e = abis
    rng = np.random.default_rng(5)
    for k, cols in [(32, 300), (4, 64), (16, 1000)]:
        enc = rand_field(rng, k * cols)
        assert np.array_equal(g.shockwave_leaves(enc, k, cols), e.shockwave_leaves(enc, k, cols)), (k, cols)


def test_whir_blocks(abis):
    g, e = abis
    rng = np.random.default_rng(6)
    for logn in [1, 4, 10, 15]:
        p = rand_field(rng, 1 << logn)
        assert np.array_equal(g.change_form(p), e.change_form(p)), logn
    for logn in [4, 11]:
        p = rand_field(rng, 1 << logn)
        assert np.array_equal(g.regroup(p, 4), e.regroup(p, 4))
    for L in [1, 8, 1 << 12, 1 << 16]:
        p, b, a = rand_field(rng, 2 * L), rand_field(rng, 2 * L), rand_field(rng, 1)
        assert np.array_equal(g.whir_poly(p, b, L), e.whir_poly(p, b, L)), L
        gp, gb = g.whir_fold(p, b, L, a)
        ep, eb = e.whir_fold(p, b, L, a)
        assert np.array_equal(gp, ep) and np.array_equal(gb, eb), L
    for v, repeats in [(1, 3), (6, 100), (9, 100), (13, 33)]:
        p, b = rand_field(rng, 1 << v), rand_field(rng, 1 << v)
        z, pw = rand_field(rng, repeats * v), rand_field(rng, repeats)
        gb, gy = g.whir_zeta(p, b, z, pw)
        eb, ey = e.whir_zeta(p, b, z, pw)
        assert np.array_equal(gy, ey), (v, repeats)
        assert np.array_equal(gb, eb), (v, repeats)


def test_select_transpose_nonzero(abis):
    g, e = abis
    rng = np.random.default_rng(7)
    M = rand_field(rng, 16 * 300)
    col = rng.integers(0, 300, size=77)
    assert np.array_equal(g.select_cols(M, 16, 300, col), e.select_cols(M, 16, 300, col))
    for rows, cols in [(32, 77), (1, 5), (100, 33), (64, 64)]:
        M = rand_field(rng, rows * cols)
        assert np.array_equal(g.transpose(M, rows, cols), e.transpose(M, rows, cols)), (rows, cols)
    z = np.zeros((5000, 2), dtype=np.uint64)
    assert g.any_nonzero(z) == 0 and e.any_nonzero(z) == 0
    z[4999, 1] = 1
    assert g.any_nonzero(z) == 1 and e.any_nonzero(z) == 1
    import torch
    t = torch.zeros((1 << 16, 2), dtype=torch.int64, device="cuda")
    flag = __import__("ctypes").c_int(7)
    g.call("hb_any_nonzero", __import__("ctypes").c_void_p(t.data_ptr()), 1 << 16, __import__("ctypes").byref(flag))
    assert flag.value == 0
    t[12345, 0] = 3
    g.call("hb_any_nonzero", __import__("ctypes").c_void_p(t.data_ptr()), 1 << 16, __import__("ctypes").byref(flag))
    assert flag.value == 1


def test_trace_streams(abis):
    """W1/W2: witness / transcript / wiring streams derived from a trace on the GPU vs the sequential restatement."""
    from helpers import synthetic_trace
    g, e = abis
    rng = np.random.default_rng(8)
    for n, cs, lookups in [(3000, 2048, False), (50000, 65536, True), (10, 16, False)]:
        tr = synthetic_trace(rng, n, lookups)
        assert (tr["type"][:n] == 0).sum() <= cs and (tr["type"][:n] != 0).sum() <= cs
        cg, dg = g.trace_load(tr)
        ce, de = e.trace_load(tr)
        assert cg == ce and dg == de == 1 and cg[0] == n
        a_w, b_w = rand_field(rng, 1), rand_field(rng, 1)
        for x, y in zip(g.trace_streams(cs, a_w, b_w, int(lookups)), e.trace_streams(cs, a_w, b_w, int(lookups))):
            assert np.array_equal(x, y)
        if lookups:
            lr = rand_field(rng, 4)
            for x, y in zip(g.trace_lookup_streams(cs, lr), e.trace_lookup_streams(cs, lr)):
                assert np.array_equal(x, y)
            assert (e.trace_lookup_streams(cs, lr)[1][1::2, 0] > 0).any()         # some repeated table entries -> non-zero access counters


@pytest.mark.parametrize("cs,B", [(1 << 10, 1 << 8), (1 << 12, 1 << 12), (1 << 14, 1 << 10)])
def test_gate_consistency_lookups(abis, cs, B):
    """S8 on a consistent random transcript (add / mul / lookup rows): every output of the GPU prover == the restatement."""
    g, e = abis
    rng = np.random.default_rng(cs + B)
    L, R = rand_field(rng, cs), rand_field(rng, cs)
    S = np.zeros((cs, 2), dtype=np.uint64); S[:, 0] = rng.integers(0, 3, size=cs)
    O = rand_field(rng, cs)
    add, mul = S[:, 0] == 0, S[:, 0] == 1
    O[add] = e.binop(0, L, R)[add]
    O[mul] = e.binop(2, L, R)[mul]
    r, lr, rnd = rand_field(rng, int(np.log2(B))), rand_field(rng, 2), rand_field(rng, 13)
    og, pg = g.gate_consistency_lookups(L, R, O, S, B, r, lr, rnd)
    oe, pe = e.gate_consistency_lookups(L, R, O, S, B, r, lr, rnd)
    assert pg == pe
    assert np.array_equal(og, oe)


@pytest.mark.parametrize("total,vectors,B,distance", [(1 << 16, 8, 1 << 8, 3), (1 << 18, 2, 1 << 9, 4), (1 << 20, 8, 1 << 9, 5)])
def test_mul_tree_stream_deep(abis, total, vectors, B, distance):
    """8f.3: layers > distance — batched streaming sumchecks on the GPU vs the restatement (products, ps; the provers' own checks
    'Error in sumcheck 1/2' fail the call on either side)."""
    g, e = abis
    rng = np.random.default_rng(total + vectors)
    xy = rand_field(rng, total)
    layers = int(np.log2(total // (2 * B)))
    if layers % distance and layers > distance:
        layers = distance + layers - layers % distance
    assert layers > distance
    batches = layers // distance
    rnd = rand_field(rng, (layers - distance) + distance * (3 * batches + 1))
    pr, xr = rand_field(rng, 1), rand_field(rng, int(np.log2(vectors)))
    og, pg, lg = g.mul_tree_stream(xy, vectors, B, distance, 0, pr, xr, rnd)
    oe, pe, le = e.mul_tree_stream(xy, vectors, B, distance, 0, pr, xr, rnd)
    assert lg == le == layers and pg == pe
    assert np.array_equal(og, oe)


@pytest.mark.parametrize("shape", [(64, 32, 16), (300, 7, 1, 5), (1, 4, 1), (1024, 256, 16), (513, 2)])
def test_mlp_evaluator(abis, shape):
    """8f.4: the MLP circuit evaluated on the GPU == the gate-by-gate restatement of MLP_inference (every derived stream, hence every
    label, access counter and value of the trace)."""
    g, e = abis
    cg, ce = g.trace_generate_mlp(shape), e.trace_generate_mlp(shape)
    assert cg == ce
    cs = 1
    while cs < ce[2]:
        cs *= 2
    rng = np.random.default_rng(len(shape))
    a_w, b_w = rand_field(rng, 1), rand_field(rng, 1)
    for x, y in zip(g.trace_streams(cs, a_w, b_w, 0), e.trace_streams(cs, a_w, b_w, 0)):
        assert np.array_equal(x, y)


@pytest.mark.parametrize("blocks", [1, 3, 64, 300])
def test_aes_evaluator(abis, blocks):
    """8f.4: the AES circuit evaluated on the GPU (closed-form records, one thread per record) == the gate-by-gate restatement of
    encrypt / AES (Seval.cpp:957-1084): every derived stream including both lookup streams, hence every label, access counter and value."""
    g, e = abis
    cg, ce = g.trace_generate_aes(blocks), e.trace_generate_aes(blocks)
    assert cg == ce == (1824 * blocks + 16 * blocks + 161, 912 * blocks, 912 * blocks + 16 * blocks + 161)
    cs = 1
    while cs < ce[2]:
        cs *= 2
    rng = np.random.default_rng(blocks)
    a_w, b_w, lr = rand_field(rng, 1), rand_field(rng, 1), rand_field(rng, 4)
    for x, y in zip(g.trace_streams(cs, a_w, b_w, 1), e.trace_streams(cs, a_w, b_w, 1)):
        assert np.array_equal(x, y)
    for x, y in zip(g.trace_lookup_streams(cs, lr), e.trace_lookup_streams(cs, lr)):
        assert np.array_equal(x, y)


@pytest.mark.parametrize("nbytes", [1 << 20, (1 << 20) + 16, (8 << 20) - 32, 8 << 20, (8 << 20) + 4096, (24 << 20) + 48, 100_000_016])
def test_pageable_copies_roundtrip(abis, nbytes):
    """Pageable host buffers >= 1 MiB cross PCIe through the context's pinned double buffer (8 MiB pieces, threaded host memcpy):
    upload + download must be the identity at and around the piece boundaries, also back to back (the halves are reused)."""
    g, _ = abis
    rng = np.random.default_rng(nbytes % 1000)
    src = rng.integers(0, 256, nbytes, dtype=np.uint8)
    dev = ctypes.c_void_p()
    g.call("hb_malloc_device", ctypes.byref(dev), nbytes)
    try:
        for _ in range(2):
            back = np.zeros(nbytes, dtype=np.uint8)
            g.call("hb_memcpy", dev, src.ctypes.data_as(ctypes.c_void_p), nbytes)
            g.call("hb_memcpy", back.ctypes.data_as(ctypes.c_void_p), dev, nbytes)
            assert np.array_equal(back, src)
            src = src[::-1].copy()
    finally:
        g.call("hb_free_device", dev)


@pytest.mark.parametrize("shape", [(64, 16, 4, 0.3), (1000, 40, 8, 0.05), (16384, 1024, 128, 0.002), (300, 7, 3, 2.5)])
def test_pruned_mlp_evaluator(abis, shape):
    """8f.4: the pruned MLP (`inference`, Seval.cpp:1170-1236) evaluated on the GPU == the gate-by-gate restatement: random sparsity patterns
    with repeated columns (access counters = stable ranks), neurons without inputs (copies of `zero`) and hidden values never read."""
    g, e = abis
    n_in, n_h, n_o, density = shape
    rng = np.random.default_rng(n_in)
    rows0 = [[] for _ in range(n_h)]
    for _ in range(int(density * n_in * n_h)):
        rows0[int(rng.integers(n_h))].append(int(rng.integers(n_in)))
    rows1 = [[] for _ in range(n_o)]
    for _ in range(int(density * 40 * n_o) + 3):
        rows1[int(rng.integers(n_o))].append(int(rng.integers(min(n_h, 256))))
    rows0[0] = []                                   # a hidden neuron without inputs ...
    rows1[0] = [0, 0, 1] + rows1[0]                 # ... that is read twice
    if n_o > 1:
        rows1[1] = []
    cg, ce = g.trace_generate_pruned(n_in, rows0, rows1), e.trace_generate_pruned(n_in, rows0, rows1)
    assert cg == ce
    cs = 1
    while cs < max(ce[1], ce[2]):
        cs *= 2
    a_w, b_w = rand_field(rng, 1), rand_field(rng, 1)
    for x, y in zip(g.trace_streams(cs, a_w, b_w, 0), e.trace_streams(cs, a_w, b_w, 0)):
        assert np.array_equal(x, y)


@pytest.mark.parametrize("rows", [1, 2, 21, 512, 5000])
def test_sql_evaluator(abis, rows):
    """8f.4: the SQL range-query circuit evaluated on the GPU (closed-form records) == the gate-by-gate restatement of range_query
    (Seval.cpp:1085-1166): every derived stream including both lookup streams."""
    g, e = abis
    cg, ce = g.trace_generate_sql(rows), e.trace_generate_sql(rows)
    assert cg == ce
    cs = 1
    while cs < max(ce[1], ce[2]):
        cs *= 2
    rng = np.random.default_rng(rows)
    a_w, b_w, lr = rand_field(rng, 1), rand_field(rng, 1), rand_field(rng, 4)
    for x, y in zip(g.trace_streams(cs, a_w, b_w, 1), e.trace_streams(cs, a_w, b_w, 1)):
        assert np.array_equal(x, y)
    for x, y in zip(g.trace_lookup_streams(cs, lr), e.trace_lookup_streams(cs, lr)):
        assert np.array_equal(x, y)
