"""Generates tests/golden/open_blocks.npz from the UNMODIFIED reference (oracle/_ref/libhobbit_ref.so): known answers for the building
blocks of the opening recursion and for the batched streaming sumcheck (8f.3).  Run here: python tests/golden/make_golden_open.py
They pin the CPU restatements (oracle/hb_emul.cpp, oracle/hobbit_oracle.c) and the CUDA kernels where the reference binary is absent."""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from helpers import Checker, F, fzeros, rand_field, srand, synthetic_stream, _p  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    ref = Checker("ref")
    L = ref.lib
    rng = np.random.default_rng(2027)
    g = {}
    r = rand_field(rng, 9)
    out = fzeros(1 << 9)
    L.ref_phi_g_init(_p(r), 9, _p(out))
    g["phi_r"], g["phi_out"] = r, out
    p = rand_field(rng, 1 << 8)
    q = p.copy()
    L.ref_change_form(_p(q), 8)
    g["cf_in"], g["cf_out"] = p, q
    poly = rand_field(rng, 1 << 11)
    poly[3 * 64:4 * 64] = 0                                   # an all-zero row
    enc, lv = fzeros(2 << 11), np.zeros((2 * 128 - 1, 32), dtype=np.uint8)
    L.ref_shockwave_commit(_p(poly), ctypes.c_size_t(1 << 11), 32, _p(enc), _p(lv))
    g["sw_poly"], g["sw_enc"], g["sw_levels"] = poly, enc, lv
    wp = rand_field(rng, 1 << 9)
    wl = np.zeros((2 * 256 - 1, 32), dtype=np.uint8)
    L.ref_whir_commit(_p(wp), ctypes.c_size_t(1 << 9), _p(wl))
    g["whir_poly"], g["whir_levels"] = wp, wl
    m, rr, prev = rand_field(rng, 1 << 7), rand_field(rng, 8), rand_field(rng, 1)
    pf = fzeros(4 * 8 + 3)
    L.ref_prove_fft.restype = ctypes.c_double
    ps = L.ref_prove_fft(_p(m), ctypes.c_size_t(1 << 7), _p(rr), _p(prev), _p(pf))
    g["pf_m"], g["pf_r"], g["pf_prev"], g["pf_out"], g["pf_ps"] = m, rr, prev, pf, np.array([ps])
    # batched streaming sumcheck on the synthetic stream (total 2^15, BUFFER_SPACE 2^9, layer 1, distance 2, 2 batches)
    total, B, layer, dist, batches = 1 << 15, 1 << 9, 1, 2, 2
    nb = (total >> layer) // (2 * B)
    r0 = rand_field(rng, 9 + int(np.log2(nb)))
    rows = [np.concatenate([r0[:9 - j * dist], rand_field(rng, int(np.log2(nb)))]) for j in range(batches)]
    oc = rand_field(rng, batches)
    srand(4)
    nc, nr, ps = ref.stream_batch(synthetic_stream(total), B, layer, dist, batches, rows, oc)
    g["sb_r0"], g["sb_r1"], g["sb_oc"], g["sb_nc"], g["sb_nr0"], g["sb_nr1"], g["sb_ps"] = rows[0], rows[1], oc, nc, nr[0], nr[1], np.array([ps])
    np.savez_compressed(os.path.join(OUT, "open_blocks.npz"), **g)
    print("wrote open_blocks.npz with", len(g), "arrays")


if __name__ == "__main__":
    main()
