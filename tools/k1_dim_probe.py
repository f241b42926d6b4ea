"""C5-shaped K1 launch (BN = 512 configuration) for d = 1, 2, 3: how much does the kernel evaluation cost?  Tooling."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from autofrk_b200 import mrts as G

for d in (1, 2, 3):
    m, kstar, n2 = 1000, 497, 2_000_000
    rng = np.random.default_rng(0)
    Xu = rng.random((m, d))
    UZ = np.zeros((m + d + 1, kstar + d + 1))
    UZ[:m, :kstar] = rng.standard_normal((m, kstar))
    plan = G.MrtsPlan(Xu, rng.standard_normal((d + 1, m)), UZ, np.ones(d), kstar)
    x = torch.rand((d, n2), dtype=torch.float64, device="cuda").T
    out = torch.empty((kstar, n2), dtype=torch.float64, device="cuda").T
    plan.predict_dev(x, out)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); plan.predict_dev(x, out); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(json.dumps(dict(d=d, ms=best, tflops=2.0 * n2 * m * kstar / best / 1e9, cfg=plan.info)), flush=True)
