"""predict.FRK (yhat + se) through frk_plan_predict_frk: wall time per call on host buffers.  Measurement tooling."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from autofrk_b200 import _lib, frk as GF, mrts as GM  # noqa: E402

rng = np.random.default_rng(0)
out = []
for m, K, T, n, n2 in ((1000, 500, 1, 20000, 4_000_000), (1000, 500, 100, 20000, 2_000_000), (500, 200, 1, 20000, 4_000_000)):
    knots = rng.random((m, 2))
    basis = GM.mrts(knots, K)
    loc = rng.random((n, 2))
    F = GM.predict_mrts(basis, loc)
    Z = F[:, :8] @ (3.0 * rng.standard_normal((8, T))) + rng.standard_normal((n, T))
    fit = GF.cMLEimat(F, Z, 0.0, wSave=True)
    plan = basis.plan()
    x = np.asfortranarray(rng.random((n2, 2)))
    yhat = np.empty((n2, T), order="F")
    se = np.empty(n2)
    w, V = np.asfortranarray(fit["w"]), np.asfortranarray(fit["V"])
    for rep in range(3):
        t0 = time.perf_counter()
        _lib.check(_lib.lib.frk_plan_predict_frk(plan._h, _lib.ptr(x), n2, _lib.ptr(w), T, _lib.ptr(V), _lib.ptr(yhat), _lib.ptr(se)))
        dt = time.perf_counter() - t0
    # check a slice against the explicit products
    Fs = GM.predict_mrts(basis, x[:2000])
    ey = np.abs(Fs @ w - yhat[:2000]).max() / np.abs(yhat[:2000]).max()
    es = np.abs(np.sqrt(np.maximum(0, np.sum((Fs @ V) * Fs, axis=1))) - se[:2000]).max() / se[:2000].max()
    rec = {"knots": m, "K": K, "T": T, "n2": n2, "rank_V": int(np.linalg.matrix_rank(V)), "s_per_call": dt,
           "locations_per_s": n2 / dt, "yhat_relerr": float(ey), "se_relerr": float(es)}
    print(json.dumps(rec))
    out.append(rec)
json.dump(out, open("gpurun_out/predict_frk_probe.json", "w"), indent=1)
