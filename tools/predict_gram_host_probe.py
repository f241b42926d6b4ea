"""Wall time of frk_plan_predict_gram on pageable host buffers, call by call (measurement tooling).
usage: python tools/predict_gram_host_probe.py [rows] [T]   (env FRK_HOST_COPY_THREADS to vary the gather threads)"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from autofrk_b200 import mrts as G
from autofrk_b200._lib import check, lib, ptr

n2 = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
T = int(sys.argv[2]) if len(sys.argv) > 2 else 100
d, m, K = 2, 1000, 500
rng = np.random.default_rng(7)
Xu = rng.random((m, d))
basis = G.mrts(Xu, K)
plan = basis.plan()
x = np.asfortranarray(rng.random((n2, d)))
Z = np.asfortranarray(rng.standard_normal((n2, T)))
Gm, Cm, zz = np.empty((K, K), order="F"), np.empty((K, T), order="F"), np.zeros(1)
walls = []
for _ in range(7):
    t0 = time.perf_counter()
    check(lib.frk_plan_predict_gram(plan._h, ptr(x), n2, ptr(Z), T, None, ptr(Gm), ptr(Cm), ptr(zz)))
    walls.append((time.perf_counter() - t0) * 1e3)
print(json.dumps({"rows": n2, "T": T, "threads": os.environ.get("FRK_HOST_COPY_THREADS", "default"),
                  "wall_ms": [round(w, 1) for w in walls]}))
