"""Quick K1 timing probe (synthetic W; values irrelevant to speed).  Measurement tooling."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from autofrk_b200 import mrts as G


def run(d, m, kstar, n2, reps=3):
    rng = np.random.default_rng(0)
    Xu = rng.random((m, d))
    UZ = np.zeros((m + d + 1, kstar + d + 1))
    UZ[:m, :kstar] = rng.standard_normal((m, kstar))
    BBBH = rng.standard_normal((d + 1, m))
    plan = G.MrtsPlan(Xu, BBBH, UZ, np.ones(d), kstar)
    x = torch.rand((d, n2), dtype=torch.float64, device="cuda").T
    out = torch.empty((kstar, n2), dtype=torch.float64, device="cuda").T
    plan.predict_dev(x, out)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        plan.predict_dev(x, out)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    flops = 2.0 * n2 * m * kstar
    res = dict(d=d, m=m, kstar=kstar, n2=n2, ms=best, tflops=flops / best / 1e9,
               basis_evals_per_s=n2 * (kstar + d + 1) / best * 1e3, info=plan.info)
    print(json.dumps(res), flush=True)
    return res


if __name__ == "__main__":
    out = []
    out.append(run(2, 1000, 497, 1_000_000))
    out.append(run(2, 1000, 497, 4_000_000))
    out.append(run(2, 500, 197, 1_000_000))
    out.append(run(3, 2000, 996, 1_000_000))
    out.append(run(2, 1000, 250, 1_000_000))
    out.append(run(2, 1000, 125, 1_000_000))
    out.append(run(1, 1000, 60, 1_000_000))
    json.dump(out, open("gpurun_out/k1_probe.json", "w"), indent=1)
