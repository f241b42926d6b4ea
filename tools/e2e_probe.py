"""End-to-end timings through the host-buffer C ABI (what R's `.Call` would see): H2D + kernels + D2H inside the timed
region, on pinned and on pageable host buffers.  Run twice (FRK_HOST_STAGING=0 and unset) to see what page-locking the
caller's buffer for the duration of the call buys.  Measurement tooling.
usage: python tools/e2e_probe.py [rows ...]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from autofrk_b200 import _lib, mrts as G
from autofrk_b200._lib import FRK_OUT_BASIS, check, lib, ptr

d, m, K = 2, 1000, 500
rows_list = [int(a) for a in sys.argv[1:]] or [1_000_000, 4_000_000]
Xu = np.random.default_rng(7).random((m, d))
fit = G.mrtsrcpp(Xu, G.xobs_diag_of(Xu, m), K - d - 1)
basis = G.MrtsBasis(fit["X"], fit["UZ"], Xu, fit["nconst"], fit["BBBH"])
plan = basis.plan()
res = {"host_staging_env": os.environ.get("FRK_HOST_STAGING", "auto"), "cpus": os.cpu_count(), "runs": []}

# host memory bandwidth of this box (one thread, numpy copy of 2 GB): the ceiling of anything that lands in host memory
a = np.ones(1 << 28)
b = np.empty_like(a)
t0 = time.perf_counter(); b[:] = a; t1 = time.perf_counter()
b[:] = a
t2 = time.perf_counter()
res["host_memcpy_gbs_one_thread"] = 2 * a.nbytes / (t2 - t1) / 1e9
del a, b


def timeit(f, reps=3):
    f()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        f()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best


for n2 in rows_list:
    rng = np.random.default_rng(1)
    x_page = np.asfortranarray(rng.random((n2, d)))
    out_page = np.empty((n2, K), order="F")
    out_page[:] = 0.0                                                     # touch the pages
    xh = torch.empty((d, n2), dtype=torch.float64).pin_memory()
    xh.copy_(torch.from_numpy(x_page.T))
    oh = torch.empty((K, n2), dtype=torch.float64).pin_memory()
    x_pin, out_pin = xh.numpy().T, oh.numpy().T
    t_pin = timeit(lambda: plan.predict(x_pin, mode=FRK_OUT_BASIS, out=out_pin))
    t_page = timeit(lambda: plan.predict(x_page, mode=FRK_OUT_BASIS, out=out_page))
    assert np.array_equal(out_pin[:1000], out_page[:1000])
    run = {"rows": n2, "K": K, "predict_basis_pinned_s": t_pin, "predict_basis_pageable_s": t_page,
           "evals_per_s_pinned": n2 * K / t_pin, "evals_per_s_pageable": n2 * K / t_page,
           "d2h_gbs_pinned": n2 * K * 8 / t_pin / 1e9, "d2h_gbs_pageable": n2 * K * 8 / t_page / 1e9}
    del oh, out_pin, out_page
    # fused paths: the result is K x K + K x T + 1 numbers (fit) or n2 x (T + 1) (prediction), not n2 x K
    for T in (1, 100):
        Z = np.asfortranarray(rng.standard_normal((n2, T)))
        Gm, Cm, zz = np.empty((K, K), order="F"), np.empty((K, T), order="F"), np.zeros(1)
        t = timeit(lambda: check(lib.frk_plan_predict_gram(plan._h, ptr(x_page), n2, ptr(Z), T, None, ptr(Gm), ptr(Cm), ptr(zz))), reps=2)
        run[f"predict_gram_T{T}_pageable_s"] = t
        run[f"predict_gram_T{T}_evals_per_s"] = n2 * K / t
        run[f"predict_gram_T{T}_h2d_bytes"] = int(n2 * (d + T) * 8)
    w = np.asfortranarray(rng.standard_normal((K, 1)))
    R = rng.standard_normal((K, 3))
    V = np.asfortranarray(R @ R.T)
    yh, se = np.empty((n2, 1), order="F"), np.empty(n2)
    t = timeit(lambda: check(lib.frk_plan_predict_frk(plan._h, ptr(x_page), n2, ptr(w), 1, ptr(V), ptr(yh), ptr(se))), reps=2)
    run["predict_frk_T1_rank3_pageable_s"] = t
    run["predict_frk_evals_per_s"] = n2 * K / t
    res["runs"].append(run)
    print(json.dumps(run), flush=True)
tag = res["host_staging_env"]
json.dump(res, open(f"gpurun_out/e2e_probe_{tag}.json", "w"), indent=1)
print(json.dumps({k: v for k, v in res.items() if k != "runs"}))
