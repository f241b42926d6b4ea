"""K1 over the shapes autoFRK selects in practice (small k*) next to the bench shape.  Measurement tooling.
usage: python tools/k1_smallk_probe.py [out.json]      (one JSON list; also a text line per shape on stderr)"""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from autofrk_b200 import mrts as G

SHAPES = [(2, 1000, 497, 2_000_000), (2, 500, 197, 2_000_000), (3, 2000, 996, 1_000_000), (2, 1000, 250, 2_000_000),
          (2, 1000, 190, 2_000_000), (2, 1000, 125, 2_000_000), (2, 1000, 90, 2_000_000), (2, 1000, 60, 2_000_000),
          (2, 200, 60, 2_000_000), (1, 1000, 60, 2_000_000), (3, 1000, 60, 2_000_000), (2, 1000, 30, 2_000_000),
          (3, 1000, 20, 2_000_000), (3, 500, 197, 2_000_000)]
res = []
for d, m, kstar, n2 in SHAPES:
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
        e0.record()
        plan.predict_dev(x, out)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    r = dict(d=d, m=m, kstar=kstar, n2=n2, ms=best, tflops=2.0 * n2 * m * kstar / best / 1e9,
             pair_evals_per_s=n2 * m / best * 1e3, info=plan.info)
    res.append(r)
    print(d, m, kstar, "ms %.3f" % best, "TF %.2f" % r["tflops"], "pair-evals/s %.3e" % r["pair_evals_per_s"],
          plan.info["tile_rows"], plan.info["slice_cols"], file=sys.stderr, flush=True)
    del x, out, plan
txt = json.dumps(res, indent=1)
if len(sys.argv) > 1:
    open(sys.argv[1], "w").write(txt)
else:
    print(txt)
