"""Phase timing of autoFRK() at BASELINE configs[0] (2,000 random 2-D locations, T=1).  Measurement tooling."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from autofrk_b200 import frk as GF, mrts as GM

rng = np.random.default_rng(0)
loc = rng.random((2000, 2))
z = (3.0 * np.sin(4 * loc[:, 0]) * np.cos(3 * loc[:, 1]) + rng.standard_normal(2000)).reshape(-1, 1)
GF.autoFRK(z, loc, maxK=20)
for rep in range(2):
    t0 = time.perf_counter(); Fk = GM.mrts(loc, 446, loc); t1 = time.perf_counter()
    G, C, zz = GF.gram_stats(Fk.X, z); t2 = time.perf_counter()
    K = GF.kgrid(2, 447)
    GF.cmle_sweep_from_stats(G, C, zz, 2000, 1, K)              # one Cholesky factor for all K
    t3 = time.perf_counter()
    for k in K:                                                  # the per-K eigen path it replaced
        GF.cmle_from_stats(G, C, zz, 2000, 1, K=int(k), onlylogLike=True)
    t3b = time.perf_counter()
    GF.cMLEimat(Fk.X[:, :18], z, 0.0, wSave=True); t4 = time.perf_counter()
    print(f"mrts fit+eval {t1-t0:.3f}s  gram {t2-t1:.3f}s  AIC sweep ({len(K)} K) {t3-t2:.4f}s (per-K eigen path {t3b-t3:.3f}s)  "
          f"final fit {t4-t3b:.3f}s  total {t4-t0-(t3b-t3):.3f}s")
t0 = time.perf_counter(); GF.autoFRK(z, loc); print("autoFRK", time.perf_counter() - t0)
