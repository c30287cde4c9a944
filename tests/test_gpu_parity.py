"""GPU parity: every C-ABI entry point of libhobbit_b200.so against the checker on the same seeded inputs.
Bit-exact (integer/byte work): np.array_equal everywhere.  The checker is the C oracle (oracle/hobbit_oracle.c) and,
where the prebuilt reference binary travelled with the repo (oracle/_ref), the unmodified reference itself."""
import numpy as np
import pytest

import ctypes

from helpers import Checker, F, P61, consistent_trace, rand_field, ref_available, srand, synthetic_chunks, synthetic_stream

pytestmark = pytest.mark.gpu

CHECKERS = ["orc"] + (["ref"] if ref_available() else [])


@pytest.fixture(scope="module")
def ctx():
    import hobbit_b200
    c = hobbit_b200.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module", params=CHECKERS)
def chk(request):
    return Checker(request.param)


def install_expander(ctx, chk, n, seed=1):
    srand(seed)
    cw = chk.expander_init_store(n)
    got = ctx.expander_set(n, chk.expander_graphs(n))
    assert got == cw
    return cw


def test_field_ops(ctx, chk):
    rng = np.random.default_rng(1)
    a, b = rand_field(rng, 1 << 16), rand_field(rng, 1 << 16)
    a[:4] = [[0, 0], [1, 0], [P61 - 1, P61 - 1], [P61 - 1, 0]]
    b[:4] = [[P61 - 1, P61 - 1], [P61 - 1, 0], [P61 - 1, P61 - 1], [0, P61 - 1]]
    for op in range(4):
        assert np.array_equal(ctx.binop(op, a, b), chk.binop(op, a, b)), op
    assert np.array_equal(ctx.binop(4, a[4:260]), chk.binop(4, a[4:260], a[4:260]))
    for n in (1, 12, 15, 20):
        assert np.array_equal(ctx.root_of_unity(n), chk.root_of_unity(n))
    assert np.array_equal(ctx.mimc(a[7], b[7]), chk.mimc(a[7], b[7]))


def test_field_empty(ctx):
    assert ctx.binop(2, np.zeros((0, 2), dtype=np.uint64), np.zeros((0, 2), dtype=np.uint64)).shape == (0, 2)


@pytest.mark.parametrize("logn", [1, 2, 5, 8, 12, 13, 15])
def test_ntt(ctx, chk, logn):
    batch = 3 if logn < 15 else 1
    x = rand_field(np.random.default_rng(logn), batch << logn)
    got = ctx.fft(x, logn, batch)
    for r in range(batch):
        want = chk.fft(x[r << logn:(r + 1) << logn], logn)
        assert np.array_equal(got[r << logn:(r + 1) << logn], want)


def test_ntt_linearity_large(ctx):
    # size-independent property at a size the oracle is too slow for: NTT(a + b) == NTT(a) + NTT(b)
    logn, batch = 12, 64
    rng = np.random.default_rng(5)
    a, b = rand_field(rng, batch << logn), rand_field(rng, batch << logn)
    s = ctx.binop(0, a, b)
    assert np.array_equal(ctx.fft(s, logn, batch), ctx.binop(0, ctx.fft(a, logn, batch), ctx.fft(b, logn, batch)))


@pytest.mark.parametrize("n", [8, 16, 64, 128, 1024])
def test_encode(ctx, chk, n):
    cw = install_expander(ctx, chk, n)
    ncols = 32
    x = rand_field(np.random.default_rng(n), n * ncols)
    got = ctx.encode(x, n, ncols).reshape(2 * n, ncols, 2)
    xm = x.reshape(n, ncols, 2)
    for c in (0, 1, 7, ncols - 1):
        want, cwl = chk.encode(np.ascontiguousarray(xm[:, c]), n)
        assert cwl == cw
        assert np.array_equal(got[:, c], want), (n, c)


def _encode_bigint(graphs, x, dep=0):
    """encode_monolithic restated on Python integers (linear_code_encode.h:62-119): enc_d(x) = x | enc_{d+1}(C_d x) | D_d enc_{d+1}(C_d x)."""
    if (0, dep) not in graphs:
        return list(x)

    def spmv(g, v):
        L, R, deg, nbr, w = g
        assert len(v) == L
        out = [[0, 0] for _ in range(R)]
        for i in range(L):
            for j in range(deg):
                t, wt = int(nbr[i * deg + j]), int(w[i * deg + j])
                out[t][0] = (out[t][0] + wt * v[i][0]) % P61
                out[t][1] = (out[t][1] + wt * v[i][1]) % P61
        return out
    z = _encode_bigint(graphs, spmv(graphs[(0, dep)], x), dep + 1)
    return list(x) + z + spmv(graphs[(1, dep)], z)


@pytest.mark.parametrize("wide", [False, True])
def test_encode_weight_paths(ctx, chk, wide):
    """Both accumulation paths of the encode kernel against a big-integer restatement: weights < 2^31 (four edges summed on the IMAD.WIDE
    addend) and a graph with weights >= 2^31 (one edge at a time), limbs at the edges of the field included."""
    n, ncols = 256, 8
    srand(7)
    chk.expander_init_store(n)
    graphs = chk.expander_graphs(n)
    rng = np.random.default_rng(11)
    if wide:
        graphs = {k: (L, R, deg, nbr, (w | (rng.integers(0, 2, len(w), dtype=np.uint64) << np.uint64(31)))) for k, (L, R, deg, nbr, w) in graphs.items()}
        assert any((g[4] >> np.uint64(31)).any() for g in graphs.values())
    else:
        graphs = {k: (L, R, deg, nbr, np.where(rng.integers(0, 4, len(w)) == 0, np.uint64((1 << 31) - 1), w)) for k, (L, R, deg, nbr, w) in graphs.items()}
    cw = ctx.expander_set(n, graphs)
    x = rand_field(rng, n * ncols).reshape(n, ncols, 2)
    x[:3, 0] = [[P61 - 1, P61 - 1], [0, P61 - 1], [P61 - 1, 0]]
    x[:, 1] = P61 - 1                                            # every limb maximal: the tightest case for the 64-bit partial sums
    got = ctx.encode(np.ascontiguousarray(x.reshape(-1, 2)), n, ncols).reshape(2 * n, ncols, 2)
    for c in (0, 1, 5):
        want = _encode_bigint(graphs, [[int(v[0]), int(v[1])] for v in x[:, c]])
        assert len(want) == cw
        assert np.array_equal(got[:cw, c], np.array(want, dtype=np.uint64)), (wide, c)
        assert not got[cw:, c].any()
    install_expander(ctx, chk, n)                                # leave a reference graph installed


def test_hashes(ctx, chk):
    rng = np.random.default_rng(3)
    s = rng.integers(0, 256, (1000, 64), dtype=np.uint8)
    got = ctx.blake3(s)
    for i in (0, 1, 500, 999):
        assert np.array_equal(got[i], chk.blake3(s[i]))
    # reference KATs
    assert ctx.blake3(np.arange(64, dtype=np.uint8))[0].tobytes().hex() == "4eed7141ea4a5cd4b788606bd23f46e212af9cacebacdc7d1f4c6dc7f2511b98"
    assert ctx.blake3(np.zeros(64, dtype=np.uint8))[0].tobytes().hex() == "4d006976636a8696d909a630a4081aad4d7c50f81afdee04020bf05086ab6a55"
    for nl in (1, 2, 64, 1024, 4096):
        lv = rng.integers(0, 256, (nl, 32), dtype=np.uint8)
        assert np.array_equal(ctx.create_tree(lv), chk.create_tree(lv)), nl
    lf = rand_field(rng, 4 * 2048)
    assert np.array_equal(ctx.mt_commit_blake(lf), chk.mt_commit_blake(lf))


@pytest.mark.parametrize("lin,trs,n", [(1, 16, 1 << 11), (1, 32, 1 << 12), (1, 128, 1 << 13), (0, 16, 1 << 11), (0, 128, 1 << 14), (0, 4, 1 << 10)])
def test_tensorcode(ctx, chk, lin, trs, n):
    if lin:
        install_expander(ctx, chk, trs)
    msg = rand_field(np.random.default_rng(n + trs), n)
    assert np.array_equal(ctx.tensorcode(msg, trs, lin), chk.tensorcode(msg, trs, lin))


@pytest.mark.parametrize("lin,trs,N,K", [(1, 16, 1 << 14, 4), (1, 64, 1 << 15, 2), (0, 16, 1 << 13, 4), (1, 16, 1 << 12, 1), (1, 1024, 1 << 18, 2)])
def test_commit_standard(ctx, chk, lin, trs, N, K):
    if lin:
        install_expander(ctx, chk, trs)
    srand(3)
    poly = chk.generate_randomness(N)
    lg, tg = ctx.commit_standard(poly, K, trs, lin, want_tensor=True)
    lw, tw = chk.commit_standard(poly, K, trs, lin, want_tensor=True)
    assert np.array_equal(tg, tw)
    assert np.array_equal(lg, lw)          # every level (the root alone depends on leaf 0 only)
    # open-side primitives on the resident tensor
    rng = np.random.default_rng(9)
    B = N // K; cols = 2 * B // trs
    col = rng.integers(0, cols, 50); row = rng.integers(0, 2 * trs, 50)
    rep = ctx.tensor_gather(col, row, K)
    T = tw.reshape(K, 2 * trs, cols, 2)
    for q in range(50):
        assert np.array_equal(rep[q], T[:, row[q], col[q]])
    beta = rand_field(rng, K)
    agg = ctx.aggregate(poly, K, beta)
    want = np.zeros((B, 2), dtype=np.uint64)
    for i in range(K):
        want = chk.binop(0, want, chk.binop(2, np.repeat(beta[i:i + 1], B, 0), poly[i * B:(i + 1) * B]))
    assert np.array_equal(agg, want)
    assert np.array_equal(ctx.aggregate(None, K, beta, N=N), want)     # resident device copy of poly


def test_commit_standard_cfg1_kat(ctx, chk):
    """BASELINE config 1 input (test_PC(2^20, 4, 32)): KAT captured from the reference."""
    N, K, trs = 1 << 20, 32, 16
    srand(1)
    poly = chk.generate_randomness(N)
    cw = chk.expander_init_store(trs)
    assert ctx.expander_set(trs, chk.expander_graphs(trs)) == cw == 27
    lv, _ = ctx.commit_standard(poly, K, trs, 1)
    assert lv[0].tobytes().hex() == "e2a5a3a00ba8238ac61e6ad0dd0193682ff0bbe15d8eac52014f6560dd0bb8e1"
    assert lv[1].tobytes().hex() == "dd3acd48eb1b58b27ec80dd1048437dd2f7ceef48fe0b695f016e312d32375e0"
    assert lv[32767].tobytes().hex() == "2a2faaac563ed907b27d64ce65089a9ddff897e220ff82d4042becfe377506e8"
    assert lv[32768 + 1].tobytes().hex() == "d29ac376c67155d04d93f0e7f2cf13871a070352185853dbb25fd2812f0f13fd"
    assert lv[-1].tobytes().hex() == "3c0096093a3cc2680dde0de2ec298ca1328d2b5e366949b0dc3e3ad4c5256650"


@pytest.mark.parametrize("lin", [0, 1])
def test_elastic_commit(ctx, chk, lin):
    N, B, trs = 1 << 14, 1 << 11, 16
    if lin:
        install_expander(ctx, chk, trs)
    chunk = ctx.stream_pc_test(B)
    assert np.array_equal(chunk, chk.read_stream_pc_test(B))
    got = ctx.elastic_commit([chunk] * (N // B), B, trs, lin)
    want = chk.elastic_commit(N, B, trs, lin)
    keep = np.ones(len(got), dtype=bool)
    if chk.kind == "ref":
        keep[4 * B - 1] = False      # the reference reads past the end for this one leaf (see oracle/hobbit_oracle.c)
    assert np.array_equal(got[keep], want[keep])


def test_elastic_zero_chunk_and_ragged_tail(ctx):
    # all-zero chunks skip the encode; K % 4 != 0 leaves the trailing chunks unhashed (Elastic_PC.cpp:206-243)
    orc = Checker("orc")
    B, trs = 1 << 11, 16
    z = np.zeros((B, 2), dtype=np.uint64)
    c = ctx.stream_pc_test(B)
    a = ctx.elastic_commit([c, z, c, c, c], B, trs, 0)
    b = ctx.elastic_commit([c, z, c, c], B, trs, 0)
    assert np.array_equal(a, b)
    assert np.array_equal(ctx.elastic_commit([c] * 4, B, trs, 0), orc.elastic_commit(4 * B, B, trs, 0))


def test_elastic_push_from_a_reused_pinned_buffer(ctx):
    """A streaming producer refills its (pinned) chunk buffer as soon as hb_elastic_push returns: the push must have READ the buffer by
    then (the copy runs on the copy stream and is waited for; round 1 returned while the DMA was still pending)."""
    B, trs = 1 << 12, 16
    rng = np.random.default_rng(31)
    stream = rand_field(rng, 12 * B, full=True)
    chunks = [stream[i * B:(i + 1) * B] for i in range(12)]
    orc = Checker("orc")
    srand(1); orc.expander_init_store(trs)
    ctx.expander_set(trs, orc.expander_graphs(trs))
    for lin in (0, 1):
        want = ctx.elastic_commit(chunks, B, trs, lin)
        got = ctx.elastic_commit(chunks, B, trs, lin, reuse_pinned=True)
        assert np.array_equal(want, got)


def test_elastic_levels_in_the_background(ctx):
    """hb_elastic_finish_levels_async + hb_levels_wait deliver exactly the levels of hb_elastic_finish (two commits queued back to back)."""
    B, trs = 1 << 14, 16
    rng = np.random.default_rng(33)
    orc = Checker("orc")
    srand(1); orc.expander_init_store(trs)
    ctx.expander_set(trs, orc.expander_graphs(trs))
    for lin in (0, 1):
        chunks = [rand_field(rng, B, full=True) for _ in range(8)]
        want = ctx.elastic_commit(chunks, B, trs, lin)
        got = ctx.elastic_commit_async_levels(chunks, B, trs, lin)
        assert np.array_equal(want, got)


def test_aggregate_never_trusts_a_host_address(ctx):
    """hb_aggregate must aggregate the data it is GIVEN: a host buffer at the same address with new contents after a commit (round 1 used
    the stale device copy whenever address and size matched)."""
    N, K, trs = 1 << 14, 4, 16
    orc = Checker("orc")
    srand(1); orc.expander_init_store(trs)
    ctx.expander_set(trs, orc.expander_graphs(trs))
    rng = np.random.default_rng(32)
    poly = rand_field(rng, N, full=True)
    ctx.commit_standard(poly, K, trs, 1)
    beta = rand_field(rng, K, full=True)
    a0 = ctx.aggregate(poly, K, beta)
    poly[:] = rand_field(rng, N, full=True)                          # same address, same size, new contents
    a1 = ctx.aggregate(poly, K, beta)
    want = np.zeros((N // K, 2), dtype=np.uint64)
    for i in range(K):
        want = orc.binop(0, want, orc.binop(2, np.repeat(beta[i:i + 1], N // K, axis=0), poly[i * (N // K):(i + 1) * (N // K)]))
    assert np.array_equal(a1, want) and not np.array_equal(a0, a1)


def test_beta_eval(ctx, chk):
    rng = np.random.default_rng(9)
    for nr in (0, 1, 5, 12, 13, 16):
        r = rand_field(rng, max(nr, 1))[:nr] if nr else np.zeros((0, 2), dtype=np.uint64)
        if nr == 0:
            assert np.array_equal(ctx.precompute_beta(r), F([1, 0]))
            continue
        assert np.array_equal(ctx.precompute_beta(r), chk.precompute_beta(r)), nr
    r = rand_field(rng, 14); v = rand_field(rng, 1 << 14)
    assert np.array_equal(ctx.evaluate_vector(v, r), chk.evaluate_vector(v, r))


@pytest.mark.parametrize("n", [1, 2, 8, 1024, 1 << 15])
def test_sumcheck_2_and_3(ctx, chk, n):
    rng = np.random.default_rng(n)
    v1, v2, v3, pr = rand_field(rng, n), rand_field(rng, n), rand_field(rng, n), rand_field(rng, 1)
    v2[: n // 4] = 0
    if n >= 2:
        a, psa = ctx.sumcheck2(v1, v2, pr); b, psb = chk.sumcheck2(v1, v2, pr)
        assert np.array_equal(a, b) and psa == psb
        a, psa = ctx.sumcheck3(v1, v2, v3, pr); b, psb = chk.sumcheck3(v1, v2, v3, pr)
        assert np.array_equal(a, b) and psa == psb
    else:
        a, _ = ctx.sumcheck3(v1, v2, v3, pr); b, _ = chk.sumcheck3(v1, v2, v3, pr)
        assert np.array_equal(a, b)


def test_sumcheck2_kat(ctx):
    """KAT captured from the reference: v1[i]=(i+1,i), v2[i]=(3i+2,7), prev_r=F(9)."""
    v1 = np.array([[i + 1, i] for i in range(8)], dtype=np.uint64)
    v2 = np.array([[3 * i + 2, 7] for i in range(8)], dtype=np.uint64)
    p, ps = ctx.sumcheck2(v1, v2, F([9, 0]))
    assert p[0].tolist() == [12, 12] and p[1].tolist() == [64, 108] and p[2].tolist() == [152, 304]
    assert p[9].tolist() == [2043575005956688095, 1046313741543418026]
    assert p[12].tolist() == [1935078710787187024, 2086613790047860569]
    assert p[14].tolist() == [92706978325602455, 344053465816199391]


@pytest.mark.parametrize("sizes", [[64, 16, 4, 1], [1024, 1024], [8], [4096, 64, 1]])
def test_batch_sumcheck3(ctx, chk, sizes):
    rng = np.random.default_rng(sum(sizes))
    tot = sum(sizes)
    t1, t2, t3, a = rand_field(rng, tot), rand_field(rng, tot), rand_field(rng, tot), rand_field(rng, len(sizes))
    pa, psa = ctx.batch_sumcheck3(t1, t2, t3, sizes, a)
    pb, psb = chk.batch_sumcheck3(t1, t2, t3, sizes, a)
    assert np.array_equal(pa, pb) and psa == psb


@pytest.mark.parametrize("vectors,n", [(2, 8), (8, 64), (1, 32), (8, 1 << 12)])
def test_mul_tree(ctx, chk, vectors, n):
    x = rand_field(np.random.default_rng(13), vectors * n)
    pr = F([32, 0])
    srand(1)
    xr = chk.generate_randomness(int(np.log2(vectors))) if vectors > 1 else None
    pa, nfa, psa = ctx.mul_tree(x, vectors, pr, xr)
    srand(1)
    pb, nfb, psb = chk.mul_tree(x, vectors, pr)
    assert nfa == nfb and psa == psb
    assert np.array_equal(pa, pb)


def libc_random():
    f = ctypes.CDLL(None).random
    f.restype = ctypes.c_long
    return int(f())


def draw_layer_randomness(orc):
    """(a, b0, b1, pad) in the reference's order: generate_randomness(1), generate_randomness(2), random()."""
    a = orc.generate_randomness(1)
    b = orc.generate_randomness(2)
    pad = F([libc_random(), 0])
    return np.concatenate([a, b, pad])


@pytest.mark.parametrize("total,B,layer,synthetic", [(1 << 13, 1 << 9, 0, True), (1 << 13, 1 << 9, 2, True), (1 << 15, 1 << 9, 1, True),
                                                     (1 << 14, 1 << 8, 0, False), (1 << 16, 1 << 10, 3, False)])
def test_stream_sumcheck_layer(ctx, chk, total, B, layer, synthetic):
    """S4.  The reference can only read its synthetic stream; the C oracle (and the GPU) take any resident stream."""
    if chk.kind == "ref" and not synthetic:
        pytest.skip("the reference only has its synthetic stream")
    xy = synthetic_stream(total) if synthetic else rand_field(np.random.default_rng(total), total)
    S = total >> layer
    r = rand_field(np.random.default_rng(total + layer), int(np.log2(S // 2)))
    oc = F([5, 0])
    orc = Checker("orc")
    srand(4); rnd = draw_layer_randomness(orc)
    got = ctx.stream_layer(xy, B, layer, r, oc, rnd)
    srand(4); want = chk.stream_layer(xy, B, layer, r, oc)
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]) and got[2] == want[2]


@pytest.mark.parametrize("total,vectors,B,synthetic", [(1 << 15, 8, 1 << 10, True), (1 << 12, 8, 1 << 11, True), (1 << 14, 2, 1 << 9, True),
                                                       (1 << 17, 8, 1 << 12, False)])
def test_mul_tree_stream(ctx, chk, total, vectors, B, synthetic):
    """S6 with the stream resident in HBM: products, ps, and (through the provers' own self-checks) every layer claim."""
    if chk.kind == "ref" and not synthetic:
        pytest.skip("the reference only has its synthetic stream")
    xy = synthetic_stream(total) if synthetic else rand_field(np.random.default_rng(total), total)
    orc = Checker("orc")
    layers = max(0, int(np.log2(total // (2 * B)))) if total > 2 * B else 0
    srand(2)
    xr = orc.generate_randomness(int(np.log2(vectors)))
    rnd = np.concatenate([draw_layer_randomness(orc) for _ in range(layers)]) if layers else np.zeros((4, 2), dtype=np.uint64)
    got = ctx.mul_tree_stream(xy, vectors, B, 5, 0, F([32, 0]), xr, rnd)
    srand(2)
    want = chk.mul_tree_stream(xy, vectors, B, 5, 0, F([32, 0]))
    assert got[2] == layers
    assert np.array_equal(got[0], want[0]) and got[1] == want[1]


@pytest.mark.parametrize("n", [2, 64, 1 << 14])
def test_gate_consistency_standard(ctx, chk, n):
    rng = np.random.default_rng(n)
    orc = Checker("orc")
    L, R, add = rand_field(rng, n), rand_field(rng, n), np.zeros((n, 2), dtype=np.uint64)
    add[:, 0] = rng.integers(0, 2, n)
    O = np.where(add[:, :1] == 1, orc.binop(0, L, R), orc.binop(2, L, R))
    r = rand_field(rng, int(np.log2(n)))
    got = ctx.gate_consistency(L, R, O, add, r)
    want = chk.gate_consistency(L, R, O, add, r)
    rounds = int(np.log2(n))
    if chk.kind == "ref":
        assert np.array_equal(got[6 * rounds:6 * rounds + 4], want)      # the reference exposes only the folded tables
    else:
        assert np.array_equal(got, want)


@pytest.mark.parametrize("cs,B", [(1 << 12, 1 << 9), (1 << 10, 1 << 10), (1 << 16, 1 << 12), (1 << 6, 2)])
def test_gate_consistency_stream(ctx, cs, B):
    """S7 against the C oracle (the reference's prove_gate_consistency returns nothing and reads its input from the circuit
    evaluator thread, so the oracle is pinned by its three internal consistency identities and the deterministic ps)."""
    orc = Checker("orc")
    rng = np.random.default_rng(cs + B)
    L, R, O, S = consistent_trace(rng, orc, cs)
    r = rand_field(rng, int(np.log2(B)))
    srand(3)
    want, wps = orc.gate_stream(L, R, O, S, B, r)
    srand(3)
    rnd = np.concatenate([orc.generate_randomness(4), orc.generate_randomness(6)])
    got, gps = ctx.gate_consistency_stream(L, R, O, S, B, r, rnd)
    assert gps == wps
    assert np.array_equal(got, want)


def test_gate_consistency_stream_rejects_bad_trace(ctx):
    import hobbit_b200
    orc = Checker("orc")
    rng = np.random.default_rng(1)
    L, R, O, S = consistent_trace(rng, orc, 1 << 10)
    O[700, 0] ^= 1                                   # one wrong gate output in the second chunk
    with pytest.raises(hobbit_b200.HobbitError, match="gate consistency"):
        ctx.gate_consistency_stream(L, R, O, S, 1 << 9, rand_field(rng, 9), rand_field(rng, 10))


@pytest.mark.parametrize("lin,trs,synthetic", [(0, 16, True), (0, 128, False), (1, 16, False)])
def test_elastic_open_front(ctx, chk, lin, trs, synthetic):
    """O2 front half: streaming aggregate + query replies (the reference only for RS columns on its synthetic stream)."""
    if chk.kind == "ref" and (lin or not synthetic):
        pytest.skip("reference: RS columns on its synthetic stream only (its Spielman reply path reads out of bounds)")
    N, B, Q = 1 << 14, 1 << 11, 300
    rng = np.random.default_rng(8 + trs)
    if lin:
        install_expander(ctx, chk, trs)
    col = rng.integers(0, 2 * B // trs, Q); row = rng.integers(0, 2 * trs, Q)
    beta = rand_field(rng, N // B)
    stream = synthetic_chunks(N // B, B) if synthetic else rand_field(rng, N)
    got = ctx.elastic_open_front([stream[i * B:(i + 1) * B] for i in range(N // B)], beta, B, trs, lin, col, row)
    want = chk.elastic_open_front(stream, B, trs, lin, beta, col, row)
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])


@pytest.mark.parametrize("n", [16, 64, 1024])
def test_encode_reseed_mirror(n):
    """E3: hobbit::encode (the graph re-drawn from fixed seeds per call) on the GPU == the reference's golden codeword, the libc state it
    leaves == the reference's, and the code installed by expander_init_store is still the resident one afterwards."""
    import ctypes
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    H = ctypes.CDLL(os.path.join(root, "hobbit_b200", "libhobbit_host.so"))
    H.hobbit_c_backend.restype = ctypes.c_void_p
    H.hobbit_c_expander_init_store.restype = ctypes.c_longlong
    import hobbit_b200
    hctx = hobbit_b200.Context.from_handle(H.hobbit_c_backend(0))
    g = np.load(os.path.join(root, "tests", "golden", "encode_reseed.npz"))
    libc = ctypes.CDLL(None); libc.rand.restype = ctypes.c_int
    srand(1)
    H.hobbit_c_expander_init_store(ctypes.c_longlong(64))                 # a resident "store" code of another size
    msg = rand_field(np.random.default_rng(3), 64 * 4)
    before = hctx.encode(msg, 64, 4)
    x = np.ascontiguousarray(g["in_%d" % n]); y = np.zeros((2 * n, 2), dtype=np.uint64)
    srand(5)
    cw = H.hobbit_c_encode_reseed(x.ctypes.data_as(ctypes.c_void_p), y.ctypes.data_as(ctypes.c_void_p), ctypes.c_longlong(n))
    assert cw == len(g["out_%d" % n]) and np.array_equal(y[:cw], g["out_%d" % n])
    assert libc.rand() == int(g["rand_after_%d" % n][0])
    assert np.array_equal(hctx.encode(msg, 64, 4), before)                # the store code was put back
