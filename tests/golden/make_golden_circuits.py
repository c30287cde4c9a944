"""Generates tests/golden/circuits.json from the UNMODIFIED reference (oracle/_ref/libhobbit_ref.so): SHA-256 of the named streams the
reference derives from its own circuit evaluator (Seval.cpp) for a small MLP, AES and pruned-MLP circuit.  They pin the circuit
evaluators (the CUDA kernels and the gate-by-gate restatements in oracle/hb_emul.cpp) — every label, access counter and
value of the trace — where the reference binary is absent.  Run here: python tests/golden/make_golden_circuits.py
One circuit per process (the reference keeps its state in globals and its producer thread never exits)."""
import ctypes
import hashlib
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CASES = {
    "mlp_64_32_16": {"fun": 9, "b": 12, "n": 12, "d": 1, "extra": [64, 32, 16]},
    "aes_16_blocks": {"fun": 5, "b": 19, "n": 4, "d": 1, "extra": []},
    "sql_512_rows": {"fun": 6, "b": 11, "n": 9, "d": 1, "extra": []},
    # pattern_order: the order in which the reference's compiler (g++, gnu++14) evaluates the two rand() calls of
    # `indexes[l][rand() % rows].push_back(rand() % cols)` (Seval.cpp:1431-1436) — found by replaying both orders against these digests
    "pruned_mlp_rate_1pct": {"fun": 8, "b": 14, "n": 20, "d": 1, "extra": [], "prune_rate": 0.01, "pattern_order": "row_first"},
}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def child(name):
    c = CASES[name]
    L = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libhobbit_ref.so"))
    if "prune_rate" in c:
        ctypes.c_double.in_dll(L, "prune_rate").value = c["prune_rate"]
    L.ref_circuit_start.restype = ctypes.c_size_t
    extra = (ctypes.c_int * max(1, len(c["extra"])))(*c["extra"])
    cs = L.ref_circuit_start(c["fun"], c["b"], c["n"], c["d"], extra, len(c["extra"]))
    lookups = c["fun"] in (5, 6)
    ctypes.c_bool.in_dll(L, "has_lookups").value = lookups              # what prove_circuit sets before it reads the streams (main.cpp:888)
    L.ref_buffer_space.restype = ctypes.c_size_t
    B = L.ref_buffer_space()
    ab = np.zeros(4, dtype=np.uint64)
    L.ref_circuit_ab(ab.ctypes.data_as(ctypes.c_void_p))
    w = np.zeros((4 * cs, 2), dtype=np.uint64)
    L.ref_dump_stream(b"witness", ctypes.c_size_t(4 * cs), ctypes.c_size_t(B), w.ctypes.data_as(ctypes.c_void_p))
    xy = np.zeros((8 * cs, 2), dtype=np.uint64)
    L.ref_dump_stream(b"wiring_consistency_check_opt", ctypes.c_size_t(8 * cs), ctypes.c_size_t(2 * B), xy.ctypes.data_as(ctypes.c_void_p))
    tr = [np.zeros((cs, 2), dtype=np.uint64) for _ in range(4)]
    L.ref_dump_trace(ctypes.c_size_t(cs), ctypes.c_size_t(B), *[t.ctypes.data_as(ctypes.c_void_p) for t in tr])
    # the reference hands the wiring stream out in blocks [X half | Y half] of 2B: back to the logical [X | Y] form of the C ABI
    blocks = xy.reshape(-1, 2, B, 2)
    xy_logical = np.concatenate([blocks[:, 0].reshape(-1, 2), blocks[:, 1].reshape(-1, 2)])
    out = {"circuit_size": int(cs), "buffer_space": int(B), "has_lookups": lookups, "a_w": [int(ab[0]), int(ab[1])], "b_w": [int(ab[2]), int(ab[3])],
           "witness": sha(w), "wiring_xy": sha(xy_logical), "L": sha(tr[0]), "R": sha(tr[1]), "O": sha(tr[2]), "S": sha(tr[3])}
    print("GOLDEN " + json.dumps(out), flush=True)
    os._exit(0)


def main():
    if len(sys.argv) > 1:
        return child(sys.argv[1])
    res = {}
    for name in CASES:
        p = subprocess.run([sys.executable, __file__, name], capture_output=True, text=True, timeout=900)
        line = [l for l in p.stdout.splitlines() if l.startswith("GOLDEN ")]
        assert line, p.stdout[-2000:] + p.stderr[-2000:]
        res[name] = dict(CASES[name], **json.loads(line[-1][7:]))
    json.dump(res, open(os.path.join(HERE, "circuits.json"), "w"), indent=1)
    print("wrote circuits.json:", {k: v["circuit_size"] for k, v in res.items()})


if __name__ == "__main__":
    main()
