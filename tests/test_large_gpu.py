"""Maximum-size and 64-bit-indexing checks on the GPU through size-independent properties (no oracle needed)."""
import numpy as np
import pytest

from util import TOL, colrel, relmax, within

pytestmark = pytest.mark.gpu


def test_max_knots_fit_and_identity():
    """maxknot = 5000 knots (R/autoFRK.R:1050 default), k* = 1200 (three column slices): the fit's invariants and
    the known-answer identity predict(mrts(knot, k), knot) == mrts(knot, k)."""
    from autofrk_b200 import mrts as G

    rng = np.random.default_rng(0)
    m, d, kstar = 5000, 2, 1200
    Xu = rng.random((m, d))
    fit = G.mrtsrcpp(Xu, G.xobs_diag_of(Xu, m), kstar)
    g = fit["X"][:, d + 1:] / np.sqrt(m)
    assert np.max(np.abs(g.T @ g - np.eye(kstar))) < TOL
    B = np.column_stack([np.ones(m), Xu])
    assert np.max(np.abs(B.T @ g)) < 1e-8
    obj = G.MrtsBasis(fit["X"], fit["UZ"], Xu, fit["nconst"], fit["BBBH"])
    assert obj.plan().info["slices"] == 3
    at = G.predict_mrts(obj, Xu)
    within("test_max_knots_fit_and_identity:at", colrel(at, fit["X"]), 1e-9)          # 5000-term sums against eigenvalues down to ~1e-6 of the largest


def test_int64_indexing_and_row_equivariance():
    """5M x 497 output (2.5e9 elements > 2^31): rows far beyond the 32-bit element range equal the same rows
    evaluated on their own, and a row permutation of the input permutes the output (bitwise)."""
    import torch
    from autofrk_b200 import mrts as G

    rng = np.random.default_rng(1)
    d, m, kstar = 2, 1000, 497
    Xu = rng.random((m, d))
    UZ = np.zeros((m + d + 1, kstar + d + 1))
    UZ[:m, :kstar] = rng.standard_normal((m, kstar))
    plan = G.MrtsPlan(Xu, rng.standard_normal((d + 1, m)), UZ, np.ones(d), kstar)
    n2 = 5_000_000
    g = torch.Generator(device="cuda").manual_seed(9)
    x = torch.rand((d, n2), dtype=torch.float64, device="cuda", generator=g).T
    out = plan.predict_dev(x)
    torch.cuda.synchronize()
    tail = slice(n2 - 1000, n2)
    assert int(tail.start) * kstar > 2 ** 31
    sub = plan.predict_dev(x[tail].T.contiguous().T)
    assert torch.equal(out[tail], sub)
    perm = torch.randperm(200_000, device="cuda", generator=g)
    xs = x[:200_000]
    a = plan.predict_dev(xs[perm].T.contiguous().T)
    assert torch.equal(a, out[:200_000][perm])
    del out, sub, a
    torch.cuda.empty_cache()


def test_config3_full_size_3d():
    """BASELINE configs[2] at full size -- 2,000 3-D knots, K = 1,000 (k* = 996, two column slices), 10 M locations,
    80 GB of basis on one GPU -- through properties that need no oracle: knots planted among the locations reproduce
    the fitted basis (predict(mrts(knot, k), knot) == mrts(knot, k)), far rows equal the same rows evaluated alone
    (64-bit indexing: 1e10 elements), and the polynomial columns are exact."""
    import torch
    from autofrk_b200 import mrts as G
    from autofrk_b200._lib import FRK_OUT_BASIS

    free, _ = torch.cuda.mem_get_info()
    if free < 100 * 2 ** 30:
        pytest.skip("needs ~85 GB of free device memory")
    rng = np.random.default_rng(3)
    d, m, K = 3, 2000, 1000
    Xu = rng.random((m, d))
    basis = G.mrts(Xu, K)
    plan = basis.plan()
    n2 = 10_000_000
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.rand((d, n2), dtype=torch.float64, device="cuda", generator=g).T
    spots = torch.tensor([0, 4_999_999, n2 - m], device="cuda")          # plant the knots at three offsets
    Xu_t = torch.from_numpy(Xu).cuda()
    for s0 in spots.tolist():
        x[s0:s0 + m] = Xu_t
    out = torch.empty((K, n2), dtype=torch.float64, device="cuda").T
    plan.predict_dev(x, out, mode=FRK_OUT_BASIS)
    torch.cuda.synchronize()
    X = torch.from_numpy(np.ascontiguousarray(basis.X)).cuda()
    scale = X.abs().amax(dim=0)
    for s0 in spots.tolist():
        err = ((out[s0:s0 + m] - X).abs().amax(dim=0) / scale).max().item()
        within("test_config3_full_size_3d:err", err, 5e-11)   # observed 1.3e-12 on B200
    assert torch.equal(out[s0:s0 + m], out[0:m])                           # same rows, same values wherever they sit
    tail = slice(n2 - 3000, n2 - m)
    sub = torch.empty((K, tail.stop - tail.start), dtype=torch.float64, device="cuda").T
    plan.predict_dev(x[tail].T.contiguous().T, sub, mode=FRK_OUT_BASIS)
    assert torch.equal(out[tail], sub)
    assert torch.all(out[:, 0] == 1.0)
    shift, nconst = torch.from_numpy(Xu.mean(axis=0)).cuda(), torch.from_numpy(np.asarray(basis.nconst)).cuda()
    assert ((out[:100_000, 1:1 + d] - (x[:100_000] - shift) / nconst).abs().max().item()) < 1e-13
    del out, sub, x
    torch.cuda.empty_cache()


def test_gram_additivity_and_trace_at_scale():
    """2M x 500 basis: Gram of the whole equals the sum over three ragged row shards; trace(G) = ||F||_F^2."""
    import ctypes as C
    import torch
    from autofrk_b200 import _lib

    n, K, T = 2_000_001, 500, 7
    g = torch.Generator(device="cuda").manual_seed(4)
    F = torch.randn((K, n + 1), dtype=torch.float64, device="cuda", generator=g).T      # ld = n + 1 (even)
    Z = torch.randn((T, n + 1), dtype=torch.float64, device="cuda", generator=g).T
    plen = _lib.lib.frk_gram_packed_len(K, T)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def stats(r0, r1):
        p = torch.empty(plen, dtype=torch.float64, device="cuda")
        _lib.check(_lib.lib.frk_gram_stats_dev(C.c_void_p(F[r0:].data_ptr()), r1 - r0, K, n + 1, C.c_void_p(Z[r0:].data_ptr()),
                                               T, n + 1, None, C.c_void_p(p.data_ptr()), st))
        return p

    whole = stats(0, n)
    parts = stats(0, 700_000) + stats(700_000, 1_300_002) + stats(1_300_002, n)
    torch.cuda.synchronize()
    scale = whole[:K * K].abs().max()
    assert float((whole - parts).abs().max() / scale) < 1e-12
    Gm = whole[:K * K].reshape(K, K)
    assert torch.equal(Gm, Gm.T)
    fro = float((F[:n] ** 2).sum())
    assert abs(float(torch.trace(Gm)) - fro) / fro < 1e-12
    assert abs(float(whole[-1]) - float((Z[:n] ** 2).sum())) / float(whole[-1]) < 1e-12


def test_second_device_in_one_process():
    """frk_set_device(1) after work on device 0: plans, workspaces, cuSOLVER handles and shared-memory opt-ins are
    per device.  Needs two visible GPUs (skipped on a one-GPU box)."""
    from autofrk_b200 import _lib, frk as G, mrts as GM

    if _lib.device_count() < 2:
        pytest.skip("needs two visible GPUs")
    rng = np.random.default_rng(2)
    Xu = rng.random((300, 2))
    x = rng.random((5000, 2))
    F0 = Z = None
    results = []
    try:
        for dev in (0, 1):
            _lib.check(_lib.lib.frk_set_device(dev))
            basis = GM.mrts(Xu, 60)
            F = GM.predict_mrts(basis, x)
            if F0 is None:
                F0 = F
                Z = F[:, :4] @ rng.standard_normal((4, 3)) + rng.standard_normal((5000, 3))
            Gm, Cm, zz = G.gram_stats(np.asfortranarray(np.hstack([F, F, F])), Z)     # K = 180: the tensor-core job kernel
            fit = G.cMLEimat(F, Z, 0.0, wSave=True)
            results.append((F, Gm, Cm, zz, fit["M"], fit["w"]))
    finally:
        _lib.check(_lib.lib.frk_set_device(0))
    for a, b in zip(results[0], results[1]):
        assert np.array_equal(np.asarray(a), np.asarray(b))      # same code, same inputs: bitwise equal across devices


def test_bitwise_repeatability():
    """The dynamic work queue of the Gram kernel and the persistent tiles of K1 must not leak into the results:
    two runs on the same inputs are bitwise equal (partials are reduced in a fixed order, no atomics on data)."""
    import ctypes as C
    import torch
    from autofrk_b200 import _lib, mrts as G

    rng = np.random.default_rng(8)
    g = torch.Generator(device="cuda").manual_seed(21)
    n, K, T = 1_000_003, 300, 40
    F = torch.randn((K, n + 1), dtype=torch.float64, device="cuda", generator=g).T
    Z = torch.randn((T, n + 1), dtype=torch.float64, device="cuda", generator=g).T
    w = torch.rand(n + 1, dtype=torch.float64, device="cuda", generator=g)
    plen = _lib.lib.frk_gram_packed_len(K, T)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for wt in (None, w):
        runs = []
        for _ in range(3):
            p = torch.empty(plen, dtype=torch.float64, device="cuda")
            _lib.check(_lib.lib.frk_gram_stats_dev(C.c_void_p(F.data_ptr()), n, K, n + 1, C.c_void_p(Z.data_ptr()), T, n + 1,
                                                   None if wt is None else C.c_void_p(wt.data_ptr()), C.c_void_p(p.data_ptr()), st))
            runs.append(p)
        torch.cuda.synchronize()
        assert torch.equal(runs[0], runs[1]) and torch.equal(runs[0], runs[2])
    d, m, kstar = 2, 600, 330
    Xu = rng.random((m, d))
    UZ = np.zeros((m + d + 1, kstar + d + 1))
    UZ[:m, :kstar] = rng.standard_normal((m, kstar))
    plan = G.MrtsPlan(Xu, rng.standard_normal((d + 1, m)), UZ, np.ones(d), kstar)
    x = torch.rand((d, 300_001), dtype=torch.float64, device="cuda", generator=g).T
    a, b = plan.predict_dev(x), plan.predict_dev(x)
    torch.cuda.synchronize()
    assert torch.equal(a, b)


def test_two_host_threads_and_two_streams():
    """Concurrency contract of include/frk_b200.h: (i) two host threads drive the library at once (per-thread handles,
    workspaces, default stream) and get exactly the results of serial calls; (ii) one thread alternating two streams on
    the same workspace / plan is ordered by the library's events -- results equal the single-stream ones bit for bit."""
    import ctypes as C
    import threading

    import torch
    from autofrk_b200 import _lib, frk as GF, mrts as G

    rng = np.random.default_rng(21)
    jobs = []
    for seed, (m, K, n, T) in enumerate([(150, 40, 60_000, 3), (220, 90, 45_000, 1)]):
        knot = rng.random((m, 2))
        x = rng.random((n, 2))
        Z = rng.standard_normal((n, T))
        jobs.append((knot, K, x, Z))

    def work(job, out, idx, reps):
        knot, K, x, Z = job
        for _ in range(reps):
            obj = G.mrts(knot, K)                       # device fit (cuSOLVER, K x K stage on this thread's default stream)
            B = G.predict_mrts(obj, x)
            fit = GF.cMLEimat(B, Z, 0.0, wSave=True)
            out[idx] = (B, fit["M"], fit["w"], fit["v"])

    serial = [None, None]
    for i, job in enumerate(jobs):
        work(job, serial, i, 1)
    par = [None, None]
    th = [threading.Thread(target=work, args=(job, par, i, 3)) for i, job in enumerate(jobs)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for s, p in zip(serial, par):
        assert np.array_equal(s[0], p[0]) and np.array_equal(s[1], p[1]) and np.array_equal(s[2], p[2]) and s[3] == p[3]

    # (ii) two streams, one thread, one workspace
    n, K, T = 300_000, 64, 4
    g = torch.Generator(device="cuda").manual_seed(4)
    F = torch.randn((K, n), dtype=torch.float64, device="cuda", generator=g).T
    Zt = torch.randn((T, n), dtype=torch.float64, device="cuda", generator=g).T
    plen = _lib.lib.frk_gram_packed_len(K, T)
    ref = torch.empty(plen, dtype=torch.float64, device="cuda")
    st0 = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def gram(rows, out, st):
        _lib.check(_lib.lib.frk_gram_stats_dev(C.c_void_p(F.data_ptr()), rows, K, n, C.c_void_p(Zt.data_ptr()), T, n, None,
                                               C.c_void_p(out.data_ptr()), st))
    refs = []
    for rows in (n, n // 2, n // 3):
        r = torch.empty(plen, dtype=torch.float64, device="cuda")
        gram(rows, r, st0)
        torch.cuda.synchronize()
        refs.append(r.clone())
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    outs = [torch.empty(plen, dtype=torch.float64, device="cuda") for _ in range(6)]
    torch.cuda.synchronize()
    for k in range(6):                                   # back to back, alternating streams, no host synchronisation
        gram((n, n // 2, n // 3)[k % 3], outs[k], C.c_void_p((s1 if k % 2 == 0 else s2).cuda_stream))
    torch.cuda.synchronize()
    for k in range(6):
        assert torch.equal(outs[k], refs[k % 3]), k


def test_large_pageable_inputs_go_through_the_pinned_ring():
    """Host matrices above 64 MB reach the device through the two-slot pinned ring (host_staging.cu): whole-column
    pieces, strided row chunks (frk_gram_stats streams F in row chunks), and columns longer than one 128 MB slot."""
    from autofrk_b200 import frk as GF

    rng = np.random.default_rng(31)
    n, K, T = 1_200_000, 40, 12                      # F 384 MB (row-chunked, strided columns), Z 115 MB
    F = np.asfortranarray(rng.standard_normal((n, K)))
    Z = np.asfortranarray(rng.standard_normal((n, T)))
    Gm, Cm, zz = GF.gram_stats(F, Z)
    within("test_large_pageable_inputs:G", relmax(Gm, F.T @ F), 1e-12)
    within("test_large_pageable_inputs:C", relmax(Cm, F.T @ Z), 1e-12)
    assert abs(zz - np.sum(Z * Z)) < 1e-12 * zz
    del F, Z
    n = 17_000_000                                    # one column = 136 MB > a slot: split by rows
    f = np.asfortranarray(rng.standard_normal((n, 2)))
    z = np.asfortranarray(rng.standard_normal((n, 1)))
    Gm, Cm, zz = GF.gram_stats(f, z)
    within("test_large_pageable_inputs:G_long_columns", relmax(Gm, f.T @ f), 1e-12)
    within("test_large_pageable_inputs:C_long_columns", relmax(Cm, f.T @ z), 1e-11)


def test_predict_gram_host_path_uploads_Z_behind_the_basis_evaluation():
    """frk_plan_predict_gram with a large Z in (pageable and pinned) host memory: Z goes up chunk by chunk on a copy
    stream while K1 evaluates the same chunk (several row chunks here).  Same chunks, same kernels, same order of the
    partial sums as the device-resident entry point: bitwise identical statistics."""
    import ctypes as C

    import torch

    from autofrk_b200 import dist as D, mrts as GM
    from autofrk_b200._lib import check, lib, ptr

    rng = np.random.default_rng(41)
    d, m, K, T, n2 = 2, 300, 200, 12, 3_000_002                    # 2 GB basis chunks of 1.34 M rows: 3 chunks; Z 288 MB
    # (even row count: the host path pads its device copies to an even leading dimension, and the Gram kernel picks
    # its variant by alignment -- with equal layouts both entry points run the same kernels)
    Xu = rng.random((m, d))
    basis = GM.mrts(Xu, K)
    plan = basis.plan()
    x = np.asfortranarray(rng.random((n2, d)))
    Z = np.asfortranarray(rng.standard_normal((n2, T)))
    w = np.asfortranarray(rng.random(n2) + 0.5)
    for weights in (None, w):
        Gm, Cm, zz = np.empty((K, K), order="F"), np.empty((K, T), order="F"), np.zeros(1)
        check(lib.frk_plan_predict_gram(plan._h, ptr(x), n2, ptr(Z), T, None if weights is None else ptr(weights),
                                        ptr(Gm), ptr(Cm), ptr(zz)))
        xt = torch.from_numpy(x.T.copy()).cuda().T               # column-major device copies
        Zt = torch.from_numpy(Z.T.copy()).cuda().T
        wt = None if weights is None else torch.from_numpy(weights).cuda()
        packed = D.predict_gram_shard(plan, xt, Zt, w_t=wt).cpu().numpy()
        within("test_predict_gram_host_path:G_vs_device_path", relmax(Gm.ravel(order="F"), packed[:K * K]), 1e-13)
        within("test_predict_gram_host_path:C_vs_device_path", relmax(Cm.ravel(order="F"), packed[K * K:K * K + K * T]), 1e-13)
        assert np.array_equal(Gm.ravel(order="F"), packed[:K * K])
        assert np.array_equal(Cm.ravel(order="F"), packed[K * K:K * K + K * T])
        assert zz[0] == packed[-1]
        del xt, Zt
    # an odd row count: the device copies get a padded leading dimension, so only closeness is required
    xo, Zo = np.asfortranarray(x[:-1]), np.asfortranarray(Z[:-1])
    Go, Co, zo = np.empty((K, K), order="F"), np.empty((K, T), order="F"), np.zeros(1)
    check(lib.frk_plan_predict_gram(plan._h, ptr(xo), n2 - 1, ptr(Zo), T, None, ptr(Go), ptr(Co), ptr(zo)))
    packed = D.predict_gram_shard(plan, torch.from_numpy(xo.T.copy()).cuda().T, torch.from_numpy(Zo.T.copy()).cuda().T).cpu().numpy()
    within("test_predict_gram_host_path:G_odd_rows", relmax(Go.ravel(order="F"), packed[:K * K]), 1e-13)
    within("test_predict_gram_host_path:C_odd_rows", relmax(Co.ravel(order="F"), packed[K * K:K * K + K * T]), 1e-12)
    del xo, Zo
    # pinned Z takes the same route with direct asynchronous copies
    Zp = torch.from_numpy(Z.T.copy()).pin_memory()
    Gp, Cp, zp = np.empty((K, K), order="F"), np.empty((K, T), order="F"), np.zeros(1)
    check(lib.frk_plan_predict_gram(plan._h, ptr(x), n2, C.c_void_p(Zp.data_ptr()), T, ptr(w), ptr(Gp), ptr(Cp), ptr(zp)))
    assert np.array_equal(Gp, Gm) and np.array_equal(Cp, Cm) and zp[0] == zz[0]
    # sanity of the statistics themselves (parity of predict -> Gram is test_gram_dev_and_fused_predict_gram's job):
    # the first basis column is the constant 1, so against Z = 1 its cross-product is the row count
    ones = np.asfortranarray(np.ones((n2, 1)))
    G1, C1, z1 = np.empty((K, K), order="F"), np.empty((K, 1), order="F"), np.zeros(1)
    check(lib.frk_plan_predict_gram(plan._h, ptr(x), n2, ptr(ones), 1, None, ptr(G1), ptr(C1), ptr(z1)))
    assert z1[0] == float(n2) and abs(C1[0, 0] - n2) <= 1e-9 * n2 and abs(G1[0, 0] - n2) <= 1e-9 * n2
