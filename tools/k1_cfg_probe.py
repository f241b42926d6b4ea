"""Time the C5 shape under different K1 configurations (FRK_K1_CONFIG).  Measurement tooling."""
import json
import os
import subprocess
import sys

code = r'''
import sys, json, numpy as np, torch
sys.path.insert(0, ".")
from autofrk_b200 import mrts as G
import os
d, m, kstar, n2 = [int(v) for v in os.environ.get('FRK_PROBE_SHAPE', '2,1000,497,2000000').split(',')]
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
print(json.dumps(dict(cfg=plan.info, ms=best, tflops=2.0 * n2 * m * kstar / best / 1e9)))
'''
for cfg in sys.argv[1:]:
    env = dict(os.environ, FRK_K1_CONFIG=cfg)
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print(cfg, r.stdout.strip()[-300:], r.stderr.strip()[-300:])
