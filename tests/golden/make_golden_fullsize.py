"""Full-size golden digests from the UNMODIFIED reference (oracle/_ref/libhobbit_ref.so), run once here (about 5 minutes of CPU):
  * BASELINE config 2: commit_standard, N = 2^26, K = 32, tensor_row_size 1024, Orion columns — the bench.py workload —
    on poly = generate_randomness(N) after srand(1) (the libc-exact generator, reproducible on any box through the C oracle);
  * BASELINE config 5 shape: Elastic_PC commit of the reference's synthetic "test" stream, N = 2^26, BUFFER_SPACE 2^20 (Orion columns).
Stored: SHA-256 of every Merkle level (leaves first) + the root, tests/golden/fullsize.json.  tests/test_gpu_fullsize.py compares the GPU
commitments (single GPU and sharded) with them: a test that fails if any of the 2^22 - 1 digests differs from the reference's."""
import hashlib
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from helpers import Checker, srand  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fullsize.json")


def level_digests(levels, nleaves):
    out, off, n = [], 0, nleaves
    while n >= 1:
        out.append(hashlib.sha256(levels[off:off + n].tobytes()).hexdigest())
        off += n
        n //= 2
    return out


def main():
    ref = Checker("ref")
    res = json.load(open(OUT)) if os.path.exists(OUT) and "--elastic-only" in sys.argv else {}
    if "--elastic-only" not in sys.argv:
        commit_standard_part(ref, res)
    N, B, trs = 1 << 26, 1 << 20, 512
    srand(1)
    ref.expander_init_store(trs)
    t0 = time.time()
    lv = ref.elastic_commit(N, B, trs, 1)
    res["elastic_commit_2e26"] = {"N": N, "B": B, "trs": trs, "linear_time": 1, "input": "srand(1); expander_init_store(trs); the reference's synthetic test stream",
                                  "root": lv[-1].tobytes().hex(), "levels_sha256": level_digests(lv, 4 * B),
                                  # the reference computes the LAST leaf from reads past its buffers (Elastic_PC.cpp:234-236, undefined behaviour; a right
                                  # child, so it influences no parent): the leaf level is also stored without it
                                  "leaves_sha256_without_last": hashlib.sha256(lv[:4 * B - 1].tobytes()).hexdigest(),
                                  "reference_seconds": time.time() - t0}
    print("elastic commit done in %.1f s, root %s" % (time.time() - t0, lv[-1].tobytes().hex()), flush=True)
    json.dump(res, open(OUT, "w"), indent=1)


def commit_standard_part(ref, res):
    N, K, trs = 1 << 26, 32, 1024
    srand(1)
    poly = ref.generate_randomness(N)
    ref.expander_init_store(trs)
    t0 = time.time()
    lv, _ = ref.commit_standard(poly, K, trs, 1)
    res["commit_standard_2e26"] = {"N": N, "K": K, "trs": trs, "linear_time": 1, "input": "srand(1); generate_randomness(N); expander_init_store(trs)",
                                   "root": lv[-1].tobytes().hex(), "levels_sha256": level_digests(lv, N // K), "reference_seconds": time.time() - t0}
    print("commit_standard done in %.1f s, root %s" % (time.time() - t0, lv[-1].tobytes().hex()), flush=True)
    del poly


if __name__ == "__main__":
    main()
