"""Profiling driver (run under ncu): one 3-product sumcheck over 2^24-entry device-resident tables + one streaming-fold layer."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import hobbit_b200
from hobbit_b200 import DevF
ctx = hobbit_b200.Context(0)
n = 1 << 24
g = torch.Generator(device="cuda"); g.manual_seed(1)
tabs = [torch.randint(0, (1 << 61) - 1, (n, 2), dtype=torch.int64, device="cuda", generator=g) for _ in range(3)]
pr = np.array([[5, 7]], dtype=np.uint64)
for _ in range(2):
    ctx.profile(True)
    out, ps = ctx.sumcheck3(*[DevF.from_torch(t) for t in tabs], pr)
    rep = ctx.profile_report()
    ctx.profile(False)
print(rep)
