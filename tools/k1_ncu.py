"""One K1 launch at a profiler-friendly size (plus one warm-up).  Measurement tooling.
usage: python tools/k1_ncu.py [d m kstar n2]"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from autofrk_b200 import mrts as G

d, m, kstar, n2 = (int(a) for a in (sys.argv[1:5] if len(sys.argv) >= 5 else (2, 1000, 497, 148 * 32 * 12)))
rng = np.random.default_rng(0)
Xu = rng.random((m, d))
UZ = np.zeros((m + d + 1, kstar + d + 1))
UZ[:m, :kstar] = rng.standard_normal((m, kstar))
plan = G.MrtsPlan(Xu, rng.standard_normal((d + 1, m)), UZ, np.ones(d), kstar)
x = torch.rand((d, n2), dtype=torch.float64, device="cuda").T
out = torch.empty((kstar, n2), dtype=torch.float64, device="cuda").T
for _ in range(2):
    plan.predict_dev(x, out)
torch.cuda.synchronize()
print("ok", plan.info)
