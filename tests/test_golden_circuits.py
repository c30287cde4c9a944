"""8f.4 circuit evaluators against golden digests captured from the UNMODIFIED reference (tests/golden/circuits.json, made by
tests/golden/make_golden_circuits.py): SHA-256 of the witness, wiring and transcript streams the reference derives from its own
evaluator.  CPU: the gate-by-gate restatements of the C-ABI emulation; GPU: the CUDA evaluators.  No reference binary needed."""
import ctypes
import hashlib
import json
import os

import numpy as np
import pytest

from helpers import RawABI

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "circuits.json")))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def pruned_pattern(rate, order):
    """The fun == 8 driver's sparsity pattern (Seval.cpp:1427-1437) replayed from libc rand() in its default state; `order` is the order
    in which the reference's compiler evaluated the two rand() calls of `indexes[l][rand() % rows].push_back(rand() % cols)`."""
    libc = ctypes.CDLL(None)
    libc.srand(1)
    rows0, rows1 = [[] for _ in range(1024)], [[] for _ in range(128)]
    for rows, count, nr, nc in ((rows0, int(rate * 1024 * 128 * 128), 1024, 128 * 128), (rows1, int(rate * 1024 * 128), 128, 256)):
        for _ in range(count):
            a, b = libc.rand(), libc.rand()
            r, c = (a % nr, b % nc) if order == "row_first" else (b % nr, a % nc)
            rows[r].append(c)
    return rows0, rows1


def check(abi, name):
    g = GOLD[name]
    if g["fun"] == 9:
        abi.trace_generate_mlp(g["extra"])
    elif g["fun"] == 5:
        abi.trace_generate_aes(1 << g["n"])
    elif g["fun"] == 6:
        abi.trace_generate_sql(1 << g["n"])
    else:
        rows0, rows1 = pruned_pattern(g["prune_rate"], g["pattern_order"])
        abi.trace_generate_pruned(128 * 128, rows0, rows1)
    cs = g["circuit_size"]
    a_w, b_w = np.array([g["a_w"]], dtype=np.uint64), np.array([g["b_w"]], dtype=np.uint64)
    w, L, R, O, S, xy = abi.trace_streams(cs, a_w, b_w, 1 if g["has_lookups"] else 0)
    got = {"witness": sha(w), "wiring_xy": sha(xy), "L": sha(L), "R": sha(R), "O": sha(O), "S": sha(S)}
    assert got == {k: g[k] for k in got}, name


@pytest.mark.parametrize("name", sorted(GOLD))
def test_evaluator_restatement_vs_reference_digests(name):
    check(RawABI("emul"), name)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(GOLD))
def test_gpu_evaluator_vs_reference_digests(name):
    check(RawABI("gpu"), name)
