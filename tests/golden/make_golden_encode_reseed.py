"""Golden vectors for E3, encode() (linear_code_encode.h:122-191), from the UNMODIFIED reference: python tests/golden/make_golden_encode_reseed.py"""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from helpers import Checker, rand_field, srand  # noqa: E402

ref = Checker("ref")
libc = ctypes.CDLL(None); libc.rand.restype = ctypes.c_int
g = {}
for n in (16, 64, 1024):
    srand(1); ref.expander_init_store(n)
    x = rand_field(np.random.default_rng(1000 + n), n)
    srand(5)
    y, cw = ref.encode_reseed(x, n)
    g["in_%d" % n], g["out_%d" % n], g["rand_after_%d" % n] = x, y, np.array([libc.rand()])
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "encode_reseed.npz"), **g)
print({k: v.shape for k, v in g.items()})
