#!/usr/bin/env python
"""Secondary measurements for the other BASELINE configs (not the driver's bench line): each case runs the same call
through the C ABI on the GPU and through the unmodified reference on the host CPU (bounded), checks the results agree
where the reference exposes them, and prints one JSON object.

    python tools/bench_extra.py [--skip-ref]
"""
import argparse
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import Checker, F, rand_field, ref_available, srand, synthetic_stream  # noqa: E402


def libc_random():
    f = ctypes.CDLL(None).random
    f.restype = ctypes.c_long
    return int(f())


def timed(fn, reps=3):
    fn()
    ts = []
    for _ in range(reps):
        t = time.perf_counter(); fn(); ts.append(time.perf_counter() - t)
    return min(ts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--skip-ref", action="store_true")
    args = ap.parse_args()
    import torch
    import hobbit_b200
    from hobbit_b200 import DevF
    ctx = hobbit_b200.Context(0)

    def dev(a):
        return DevF.from_torch(torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).cuda())

    orc = Checker("orc")
    ref = Checker("ref") if (ref_available() and not args.skip_ref) else None
    out = {}

    # ---- config 1: test_PC(2^20, 4, 32): commit_standard + the data-parallel front half of open_standard ----------------
    N, K, trs = 1 << 20, 32, 16
    srand(1)
    poly = orc.generate_randomness(N)
    cw = orc.expander_init_store(trs)
    ctx.expander_set(trs, orc.expander_graphs(trs))
    lv = [None]
    t_commit = timed(lambda: lv.__setitem__(0, ctx.commit_standard(poly, K, trs, 1)[0]))
    root = lv[0][-1].tobytes().hex()
    beta = rand_field(np.random.default_rng(1), K)
    rng = np.random.default_rng(2)
    col = rng.integers(0, 2 * (N // K) // trs, 5900); row = rng.integers(0, 2 * trs, 5900)

    def open_front():
        ctx.aggregate(None, K, beta, N=N)
        ctx.tensor_gather(col, row, K)
    t_open = timed(open_front)
    c1 = {"gpu_commit_s": t_commit, "gpu_open_front_s": t_open, "root": root,
          "root_matches_reference_kat": root == "3c0096093a3cc2680dde0de2ec298ca1328d2b5e366949b0dc3e3ad4c5256650",
          "field_elems_per_s": N / t_commit}
    if ref:
        srand(1); ref.generate_randomness(N); ref.expander_init_store(trs)
        t = time.perf_counter(); rl, _ = ref.commit_standard(poly, K, trs, 1); c1["ref_commit_s"] = time.perf_counter() - t
        c1["levels_identical"] = bool(np.array_equal(rl, lv[0]))
    out["cfg1_test_PC_2^20_K32"] = c1

    # ---- Elastic_PC streaming commit, N = 2^22, BUFFER_SPACE = 2^18 (test_ourPC_buff.sh scaled), both column codes ---------
    for lin, trs_e in ((0, 128), (1, 16)):
        Ne, B = 1 << 22, 1 << 18
        if lin:
            srand(1); orc.expander_init_store(trs_e); ctx.expander_set(trs_e, orc.expander_graphs(trs_e))
        chunk = ctx.stream_pc_test(B)
        res = [None]
        t_g = timed(lambda: res.__setitem__(0, ctx.elastic_commit([chunk] * (Ne // B), B, trs_e, lin)), reps=2)
        dchunk = dev(chunk)
        t_r = timed(lambda: ctx.elastic_commit([dchunk] * (Ne // B), B, trs_e, lin), reps=2)
        e = {"gpu_commit_s": t_g, "gpu_commit_resident_s": t_r, "field_elems_per_s": Ne / t_g, "field_elems_per_s_resident": Ne / t_r,
             "root": res[0][-1].tobytes().hex()}
        if ref:
            if lin:
                srand(1); ref.expander_init_store(trs_e)
            t = time.perf_counter(); rl = ref.elastic_commit(Ne, B, trs_e, lin); e["ref_commit_s"] = time.perf_counter() - t
            keep = np.ones(len(rl), dtype=bool); keep[4 * B - 1] = False
            e["levels_identical"] = bool(np.array_equal(rl[keep], res[0][keep]))
        out["elastic_commit_2^22_B2^18_%s" % ("spielman" if lin else "rs")] = e

    # ---- config 3 shape: streaming product tree, 8 vectors x 2^20 (MLP wiring stream size), BUFFER_SPACE = 2^18 ---------
    total, vectors, B = 1 << 23, 8, 1 << 18
    xy = synthetic_stream(total)
    layers = int(np.log2(total // (2 * B)))

    def draw():
        srand(2)
        xr = orc.generate_randomness(int(np.log2(vectors)))
        rnd = []
        for _ in range(layers):
            rnd += [orc.generate_randomness(1), orc.generate_randomness(2), F([libc_random(), 0])]
        return xr, np.concatenate(rnd)
    xr, rnd = draw()
    res = [None]
    t_g = timed(lambda: res.__setitem__(0, ctx.mul_tree_stream(xy, vectors, B, 5, 0, F([32, 0]), xr, rnd)), reps=2)
    dxy = dev(xy)
    t_r = timed(lambda: ctx.mul_tree_stream(dxy, vectors, B, 5, 0, F([32, 0]), xr, rnd), reps=2)
    s6 = {"gpu_s": t_g, "gpu_resident_s": t_r, "layers_streamed": res[0][2], "ps_kb": res[0][1], "stream_elems_per_s": total / t_g,
          "stream_elems_per_s_resident": total / t_r}
    if ref:
        srand(2)
        t = time.perf_counter(); ro, rps = ref.mul_tree_stream(xy, vectors, B, 5, 0, F([32, 0])); s6["ref_s"] = time.perf_counter() - t
        s6["identical"] = bool(np.array_equal(ro, res[0][0]) and rps == res[0][1])
    out["mul_tree_stream_8x2^20_B2^18"] = s6

    # ---- 3-product sumcheck over 2^22-entry tables (the HBM-side kernel class) ----------------------------------------------
    n = 1 << 22
    rng = np.random.default_rng(3)
    v1, v2, v3, pr = rand_field(rng, n), rand_field(rng, n), rand_field(rng, n), rand_field(rng, 1)
    res = [None]
    t_g = timed(lambda: res.__setitem__(0, ctx.sumcheck3(v1, v2, v3, pr)), reps=2)
    d1, d2, d3 = dev(v1), dev(v2), dev(v3)
    t_r = timed(lambda: ctx.sumcheck3(d1, d2, d3, pr), reps=3)
    ctx.profile(True)
    ctx.sumcheck3(d1, d2, d3, pr)
    prof = ctx.profile_report()
    ctx.profile(False)
    sc = {"gpu_s_incl_h2d": t_g, "gpu_resident_s": t_r, "n": n, "rounds": int(np.log2(n)), "kernels": prof}
    if ref:
        t = time.perf_counter(); ro, rps = ref.sumcheck3(v1, v2, v3, pr); sc["ref_s"] = time.perf_counter() - t
        sc["identical"] = bool(np.array_equal(ro, res[0][0]))
    out["sumcheck3_2^22"] = sc
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
