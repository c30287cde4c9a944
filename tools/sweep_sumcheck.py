"""Development tool: time the sumcheck round kernel of several library builds (kernel variants compiled side by side by
tools/build_variants.sh) on the same box.  For every library: 3-product sumcheck over 2^24- and 2^23-entry tables, CUDA-event time of all
sc_round_kernel launches (min over repetitions); the difference of the two totals is the 2^23-pair round.  The proof digest must agree
across variants.  Usage: python tools/sweep_sumcheck.py lib1.so lib2.so ...   (one subprocess per library)"""
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

if len(sys.argv) > 1 and sys.argv[1] == "--one":
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch
    import hobbit_b200
    from hobbit_b200 import DevF
    ctx = hobbit_b200.Context(0)
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    tabs = [torch.randint(0, (1 << 61) - 1, (1 << 24, 2), dtype=torch.int64, device="cuda", generator=g) for _ in range(3)]
    pr = np.array([[5, 7]], dtype=np.uint64)
    out = {}
    for logn in (24, 23):
        dv = [DevF.from_torch(t[: 1 << logn]) for t in tabs]
        best = 1e9
        for _ in range(6):
            ctx.profile(True)
            proof, _ = ctx.sumcheck3(dv[0], dv[1], dv[2], pr)
            rep = ctx.profile_report()
            ctx.profile(False)
            best = min(best, rep["sc_round_kernel"]["total_ms"])
        out["rounds_ms_2^%d" % logn] = best
        out["digest_2^%d" % logn] = hashlib.sha256(np.ascontiguousarray(proof).tobytes()).hexdigest()[:16]
    out["round0_us"] = 1e3 * (out["rounds_ms_2^24"] - out["rounds_ms_2^23"])
    out["hbm_frac_round0"] = (3 * 48 * (1 << 23)) / (out["round0_us"] * 1e-6) / 6452.5e9
    print(json.dumps(out))
    sys.exit(0)

for lib in sys.argv[1:]:
    env = dict(os.environ, HOBBIT_B200_LIB=os.path.abspath(lib))
    p = subprocess.run([sys.executable, os.path.abspath(__file__), "--one"], env=env, capture_output=True, text=True)
    line = p.stdout.strip().splitlines()[-1] if p.stdout.strip() else "{\"error\": %s}" % json.dumps(p.stderr[-400:])
    print(json.dumps({"lib": os.path.basename(lib), **json.loads(line)}), flush=True)
