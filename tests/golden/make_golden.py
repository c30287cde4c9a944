"""Generates tests/golden/*.npz from the UNMODIFIED reference compiled in place (oracle/_ref/libhobbit_ref.so).
Run here (needs /root/reference to have been built by oracle/Makefile):  python tests/golden/make_golden.py
The vectors are small, committed, and pin both the C oracle (CPU tests) and the CUDA path (GPU tests) on machines
where the reference binary is not available."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from helpers import Checker, F, rand_field, srand  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    ref = Checker("ref")
    rng = np.random.default_rng(2026)
    g = {}
    # field + mimc + root of unity
    a, b = rand_field(rng, 64), rand_field(rng, 64)
    g["f_a"], g["f_b"] = a, b
    for op, nm in enumerate(["add", "sub", "mul", "neg", "inv"]):
        g["f_" + nm] = ref.binop(op, a, b)
    g["rou"] = np.concatenate([ref.root_of_unity(n) for n in range(1, 21)])
    g["mimc"] = np.concatenate([ref.mimc(a[i], b[i]) for i in range(8)])
    # NTT
    x = rand_field(rng, 1 << 10)
    g["fft_in"], g["fft_out"] = x, ref.fft(x, 10)
    # RNG-driven generators from a fresh libc state (seed 1 == fresh process)
    srand(1)
    g["rand300"] = ref.generate_randomness(300)
    srand(1)
    g["exp64_cw"] = np.array([ref.expander_init_store(64)])
    gr = ref.expander_graphs(64)
    for (which, dep), (L, R, deg, nbr, w) in gr.items():
        g["exp64_%d_%d_dims" % (which, dep)] = np.array([L, R, deg])
        g["exp64_%d_%d_nbr" % (which, dep)] = nbr
        g["exp64_%d_%d_w" % (which, dep)] = w
    m = rand_field(rng, 64)
    g["enc64_in"] = m
    g["enc64_out"] = ref.encode(m, 64)[0]
    # hashes
    s = rng.integers(0, 256, (16, 64), dtype=np.uint8)
    g["b3_in"] = s
    g["b3_out"] = np.stack([ref.blake3(r) for r in s])
    lf = rand_field(rng, 256)
    g["mt_in"], g["mt_out"] = lf, ref.mt_commit_blake(lf)
    # tensor code + commits (expander for n=16 from a fresh RNG)
    srand(1)
    ref.expander_init_store(16)
    g16 = ref.expander_graphs(16)
    for (which, dep), (L, R, deg, nbr, w) in g16.items():
        g["exp16_%d_%d_dims" % (which, dep)] = np.array([L, R, deg])
        g["exp16_%d_%d_nbr" % (which, dep)] = nbr
        g["exp16_%d_%d_w" % (which, dep)] = w
    msg = rand_field(rng, 1 << 11)
    g["tc_msg"] = msg
    g["tc_lin"] = ref.tensorcode(msg, 16, 1)
    g["tc_rs"] = ref.tensorcode(msg, 16, 0)
    poly = rand_field(rng, 1 << 13, full=False)
    g["cs_poly"] = poly
    g["cs_lin_levels"] = ref.commit_standard(poly, 4, 16, 1)[0]
    g["cs_rs_levels"] = ref.commit_standard(poly, 4, 16, 0)[0]
    lv = ref.elastic_commit(1 << 14, 1 << 11, 16, 1)
    lv[4 * (1 << 11) - 1] = 0          # the one heap-dependent leaf of the reference (see oracle/hobbit_oracle.c)
    g["el_lin_levels"] = lv
    lv = ref.elastic_commit(1 << 14, 1 << 11, 16, 0)
    lv[4 * (1 << 11) - 1] = 0
    g["el_rs_levels"] = lv
    # eq / MLE / sumchecks
    r = rand_field(rng, 8); v = rand_field(rng, 256)
    g["beta_r"], g["beta_out"] = r, ref.precompute_beta(r)
    g["eval_v"], g["eval_out"] = v, ref.evaluate_vector(v, r)
    v1, v2, v3, pr = rand_field(rng, 256), rand_field(rng, 256), rand_field(rng, 256), rand_field(rng, 1)
    v2[:64] = 0
    g["sc_v1"], g["sc_v2"], g["sc_v3"], g["sc_pr"] = v1, v2, v3, pr
    g["sc2_out"], ps2 = ref.sumcheck2(v1, v2, pr)
    g["sc3_out"], ps3 = ref.sumcheck3(v1, v2, v3, pr)
    sizes = [128, 64, 32, 16, 8, 4, 2, 1, 1]          # 256 total
    av = rand_field(rng, len(sizes))
    g["bsc_sizes"], g["bsc_a"] = np.array(sizes), av
    g["bsc_out"], psb = ref.batch_sumcheck3(v1, v2, v3, sizes, av)
    srand(1)
    g["mt8_out"], nf, psm = ref.mul_tree(v1, 8, F([32, 0]))
    srand(1)
    g["mt8_xr"] = ref.generate_randomness(3)
    g["mt1_out"], nf1, psm1 = ref.mul_tree(v1[:32], 1, F([32, 0]))
    g["ps"] = np.array([ps2, ps3, psb, psm, psm1, nf, nf1], dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "hotpath.npz"), **g)
    print("wrote", os.path.join(OUT, "hotpath.npz"), "with", len(g), "arrays")


if __name__ == "__main__":
    main()
