"""Golden vectors (tests/golden/hotpath.npz, generated from the unmodified reference by tests/golden/make_golden.py).
CPU: the C oracle reproduces every vector.  GPU: the CUDA library reproduces every vector through the C ABI."""
import os

import numpy as np
import pytest

from helpers import Checker, F, ROOT, srand

G = np.load(os.path.join(ROOT, "tests", "golden", "hotpath.npz"))


def graphs(prefix):
    out = {}
    for k in G.files:
        if k.startswith(prefix) and k.endswith("_dims"):
            _, which, dep, _ = k.split("_")
            L, R, deg = (int(x) for x in G[k])
            out[(int(which), int(dep))] = (L, R, deg, G["%s_%s_%s_nbr" % (prefix, which, dep)], G["%s_%s_%s_w" % (prefix, which, dep)])
    return out


class OracleImpl:
    """C oracle behind the common interface; expander graphs come from its own libc-RNG generator."""
    def __init__(self):
        self.c = Checker("orc")

    def install(self, n):
        srand(1)
        return self.c.expander_init_store(n)

    def mul_tree(self, x, vectors, pr, xr):
        srand(1)
        return self.c.mul_tree(x, vectors, pr)

    def elastic(self, N, B, trs, lin):
        lv = self.c.elastic_commit(N, B, trs, lin)
        lv[4 * B - 1] = 0
        return lv

    def __getattr__(self, k):
        return getattr(self.c, k)


class GpuImpl:
    """CUDA library behind the same interface; expander graphs are the GOLDEN ones (host-generated in the reference)."""
    def __init__(self):
        import hobbit_b200
        self.c = hobbit_b200.Context(0)

    def install(self, n):
        return self.c.expander_set(n, graphs("exp%d" % n))

    def encode(self, m, n):
        return self.c.encode(m, n, 1), None

    def fft(self, x, logn):
        return self.c.fft(x, logn, 1)

    def mul_tree(self, x, vectors, pr, xr):
        return self.c.mul_tree(x, vectors, pr, xr)

    def commit_standard(self, poly, K, trs, lin):
        return self.c.commit_standard(poly, K, trs, lin)

    def elastic(self, N, B, trs, lin):
        chunk = self.c.stream_pc_test(B)
        lv = self.c.elastic_commit([chunk] * (N // B), B, trs, lin)
        lv[4 * B - 1] = 0
        return lv

    def __getattr__(self, k):
        return getattr(self.c, k)


def run_all(impl, is_oracle):
    a, b = G["f_a"], G["f_b"]
    for op, nm in enumerate(["add", "sub", "mul", "neg", "inv"]):
        assert np.array_equal(impl.binop(op, a, b), G["f_" + nm]), nm
    assert np.array_equal(np.concatenate([impl.root_of_unity(n) for n in range(1, 21)]), G["rou"])
    assert np.array_equal(np.concatenate([impl.mimc(a[i], b[i]) for i in range(8)]), G["mimc"])
    assert np.array_equal(impl.fft(G["fft_in"], 10), G["fft_out"])
    if is_oracle:      # libc-RNG-driven generators (host side of the reference)
        srand(1)
        assert np.array_equal(impl.generate_randomness(300), G["rand300"])
        srand(1)
        assert impl.expander_init_store(64) == int(G["exp64_cw"][0])
        got, want = impl.expander_graphs(64), graphs("exp64")
        for k in want:
            assert got[k][:3] == want[k][:3] and np.array_equal(got[k][3], want[k][3]) and np.array_equal(got[k][4], want[k][4])
    assert impl.install(64) == int(G["exp64_cw"][0])
    assert np.array_equal(np.asarray(impl.encode(G["enc64_in"], 64)[0]).reshape(-1, 2), G["enc64_out"])
    b3 = impl.blake3(G["b3_in"]) if not is_oracle else np.stack([impl.blake3(r) for r in G["b3_in"]])
    assert np.array_equal(b3, G["b3_out"])
    assert np.array_equal(impl.mt_commit_blake(G["mt_in"]), G["mt_out"])
    impl.install(16)
    assert np.array_equal(impl.tensorcode(G["tc_msg"], 16, 1), G["tc_lin"])
    assert np.array_equal(impl.tensorcode(G["tc_msg"], 16, 0), G["tc_rs"])
    assert np.array_equal(impl.commit_standard(G["cs_poly"], 4, 16, 1)[0], G["cs_lin_levels"])
    assert np.array_equal(impl.commit_standard(G["cs_poly"], 4, 16, 0)[0], G["cs_rs_levels"])
    assert np.array_equal(impl.elastic(1 << 14, 1 << 11, 16, 1), G["el_lin_levels"])
    assert np.array_equal(impl.elastic(1 << 14, 1 << 11, 16, 0), G["el_rs_levels"])
    assert np.array_equal(impl.precompute_beta(G["beta_r"]), G["beta_out"])
    assert np.array_equal(impl.evaluate_vector(G["eval_v"], G["beta_r"]), G["eval_out"])
    v1, v2, v3, pr = G["sc_v1"], G["sc_v2"], G["sc_v3"], G["sc_pr"]
    ps = G["ps"]
    o, p = impl.sumcheck2(v1, v2, pr); assert np.array_equal(o, G["sc2_out"]) and p == ps[0]
    o, p = impl.sumcheck3(v1, v2, v3, pr); assert np.array_equal(o, G["sc3_out"]) and p == ps[1]
    o, p = impl.batch_sumcheck3(v1, v2, v3, [int(x) for x in G["bsc_sizes"]], G["bsc_a"]); assert np.array_equal(o, G["bsc_out"]) and p == ps[2]
    o, nf, p = impl.mul_tree(v1, 8, F([32, 0]), G["mt8_xr"]); assert np.array_equal(o, G["mt8_out"]) and p == ps[3] and nf == int(ps[5])
    o, nf, p = impl.mul_tree(v1[:32], 1, F([32, 0]), None); assert np.array_equal(o, G["mt1_out"]) and p == ps[4] and nf == int(ps[6])


def test_oracle_reproduces_golden():
    run_all(OracleImpl(), True)


@pytest.mark.gpu
def test_gpu_reproduces_golden():
    run_all(GpuImpl(), False)


def test_encode_reseed_golden():
    """E3, encode() (linear_code_encode.h:122-191): the C restatement against vectors from the unmodified reference, codeword and the libc
    state the call leaves behind."""
    import ctypes
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "encode_reseed.npz"))
    orc = Checker("orc")
    libc = ctypes.CDLL(None); libc.rand.restype = ctypes.c_int
    for n in (16, 64, 1024):
        srand(5)
        y, cw = orc.encode_reseed(g["in_%d" % n], n)
        assert cw == len(g["out_%d" % n]) and np.array_equal(y, g["out_%d" % n])
        assert libc.rand() == int(g["rand_after_%d" % n][0])
