import sys, json, numpy as np, torch
sys.path.insert(0, ".")
from autofrk_b200 import mrts as G
for d, m, kstar in [(2,1000,60),(2,1000,30),(1,1000,60),(3,1000,20),(2,500,197),(2,200,60)]:
    n2=2_000_000
    rng = np.random.default_rng(0)
    Xu = rng.random((m, d)); UZ = np.zeros((m + d + 1, kstar + d + 1)); UZ[:m, :kstar] = rng.standard_normal((m, kstar))
    plan = G.MrtsPlan(Xu, rng.standard_normal((d + 1, m)), UZ, np.ones(d), kstar)
    x = torch.rand((d, n2), dtype=torch.float64, device="cuda").T
    out = torch.empty((kstar, n2), dtype=torch.float64, device="cuda").T
    plan.predict_dev(x, out); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); plan.predict_dev(x, out); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(d, m, kstar, "ms %.3f" % best, "TF %.2f" % (2.0*n2*m*kstar/best/1e9), "pair-evals/s %.3e" % (n2*m/best*1e3), plan.info["slice_cols"], flush=True)
