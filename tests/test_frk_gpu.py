"""K4/K5 parity on the GPU: Gram reductions, the K x K cMLE stage, autoFRK() and predict.FRK() vs the oracle."""
import numpy as np
import pytest

from util import TOL, colrel, relmax, within

pytestmark = pytest.mark.gpu


def _data(n, K, T, seed, rank=5):
    rng = np.random.default_rng(seed)
    F = rng.standard_normal((n, K))
    F[:, 0] = 1.0
    wts = np.zeros((K, T))
    wts[:rank] = rng.standard_normal((rank, T)) * 3.0
    Z = F @ wts + rng.standard_normal((n, T))
    return F, Z


@pytest.mark.parametrize("n,K,T", [(5000, 30, 3), (20011, 130, 5), (4096, 128, 128), (30000, 500, 100), (777, 257, 1),
                                   (3000, 40, 0), (20011, 64, 4), (9999, 100, 1), (15000, 200, 20), (8000, 49, 33),
                                   (6000, 446, 1), (5000, 97, 40), (4000, 65, 0), (2500, 1000, 7)])
def test_gram_stats_matches_numpy(n, K, T):
    from autofrk_b200 import frk as G

    F, Z = _data(n, K, max(T, 1), seed=n + K)
    if T == 0:
        Gm, Cm, zz = G.gram_stats(F)
        assert Cm is None and zz == 0.0
    else:
        Gm, Cm, zz = G.gram_stats(F, Z)
        within("test_gram_stats_matches_numpy:Cm", relmax(Cm, F.T @ Z), 1e-12)
        assert abs(zz - np.sum(Z * Z)) / np.sum(Z * Z) < 1e-13
    within("test_gram_stats_matches_numpy:Gm", relmax(Gm, F.T @ F), 1e-12)
    assert np.array_equal(Gm, Gm.T)          # exactly symmetric by construction


def test_gram_dev_and_fused_predict_gram():
    """device-pointer API (row shards + packed buffer) and the fused predict -> Gram path (F never materialised)."""
    import ctypes as C
    import torch
    import oracle as O
    from oracle.mrts_oracle import xobs_diag_of
    from autofrk_b200 import _lib, mrts as GM

    rng = np.random.default_rng(0)
    d, m, K, T, n = 2, 300, 60, 4, 50_001
    Xu = rng.random((m, d))
    fit = O.mrtsrcpp(Xu, xobs_diag_of(Xu, m), K - d - 1)
    obj = GM.MrtsBasis(fit["X"], fit["UZ"], Xu, fit["nconst"], fit["BBBH"])
    plan = obj.plan()
    x = torch.rand((d, n), dtype=torch.float64, device="cuda").T
    Z = torch.randn((T, n), dtype=torch.float64, device="cuda").T
    F = plan.predict_dev(x, mode=_lib.FRK_OUT_BASIS)
    plen = _lib.lib.frk_gram_packed_len(K, T)
    assert plen == K * K + K * T + 1
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    # (1) materialised F, two row shards reduced like the multi-GPU path (sum of packed buffers)
    packed = torch.zeros(plen, dtype=torch.float64, device="cuda")
    for r0, r1 in ((0, 20_000), (20_000, n)):
        part = torch.empty(plen, dtype=torch.float64, device="cuda")
        _lib.check(_lib.lib.frk_gram_stats_dev(C.c_void_p(F[r0:].data_ptr()), r1 - r0, K, n, C.c_void_p(Z[r0:].data_ptr()),
                                               T, n, None, C.c_void_p(part.data_ptr()), st))
        packed += part
    torch.cuda.synchronize()
    Fh, Zh = F.cpu().numpy(), Z.cpu().numpy()
    Gd = packed[:K * K].cpu().numpy().reshape(K, K, order="F")
    Cd = packed[K * K:K * K + K * T].cpu().numpy().reshape(K, T, order="F")
    assert relmax(Gd, Fh.T @ Fh) < 1e-12 and relmax(Cd, Fh.T @ Zh) < 1e-12
    assert abs(packed[-1].item() - np.sum(Zh * Zh)) / np.sum(Zh * Zh) < 1e-13
    # (2) fused: chunks of 8192 rows, basis never stored
    fused = torch.empty(plen, dtype=torch.float64, device="cuda")
    _lib.check(_lib.lib.frk_plan_predict_gram_dev(plan._h, C.c_void_p(x.data_ptr()), n, n, C.c_void_p(Z.data_ptr()), T, n, None,
                                                  C.c_void_p(fused.data_ptr()), 8192, st))
    torch.cuda.synchronize()
    within("test_gram_dev_and_fused_predict_gram:fused.cpu.numpy", relmax(fused.cpu().numpy(), packed.cpu().numpy()), 1e-12)


@pytest.mark.parametrize("n,Kmax,T,s", [(6000, 60, 1, 0.0), (6000, 60, 7, 0.0), (3000, 40, 64, 0.0), (5000, 33, 5, 0.3)])
def test_cMLEimat_sweep_matches_oracle_per_K(n, Kmax, T, s):
    """the K loop of selectBasis (R/autoFRK.R:305-317) from one Cholesky factor: the same negloglik as the oracle's
    literal cMLEimat(Fk[, 1:K], ...) and as the per-K eigen path, for every K, T < K and T >= K alike."""
    import oracle as O
    from autofrk_b200 import frk as G

    F, Z = _data(n, Kmax, T, seed=Kmax + T)
    Gm, Cm, zz = G.gram_stats(F, Z)
    Ks = [1, 2, 3, 5, 8, 13, 21, Kmax - 1, Kmax]
    got = G.cmle_sweep_from_stats(Gm, Cm, zz, n, T, Ks, s=s)
    for k, nll in zip(Ks, got):
        ref = O.cMLEimat(F[:, :k], Z, s, wSave=False)["negloglik"]
        eig = G.cmle_from_stats(Gm, Cm, zz, n, T, s=s, onlylogLike=True, K=k)["negloglik"]
        assert abs(nll - ref) <= TOL * abs(ref), (k, nll, ref)
        assert abs(nll - eig) <= TOL * abs(eig), (k, nll, eig)


def test_cMLEimat_sweep_falls_back_on_singular_gram():
    """a rank-deficient basis (duplicated column): the Cholesky factor is refused from that K on (NaN from the C ABI)
    and the wrapper takes the eigen path with getHalf's zero-eigenvalue rule (R/helper.R:22-24)."""
    import ctypes as C
    from autofrk_b200 import _lib, frk as G

    n, T = 4000, 3
    F, Z = _data(n, 12, T, seed=3)
    F = np.asfortranarray(np.column_stack([F[:, :8], F[:, 2], F[:, 8:]]))     # column 8 duplicates column 2
    Gm, Cm, zz = G.gram_stats(F, Z)
    Ks = np.array([4, 8, 9, 13], dtype=np.int32)
    raw = np.empty(Ks.size)
    _lib.check(_lib.lib.frk_cMLEimat_sweep(_lib.ptr(Gm), 13, _lib.ptr(Cm), 13, zz, T, n, 0.0, Ks.ctypes.data_as(C.c_void_p),
                                           Ks.size, raw.ctypes.data_as(C.c_void_p)))
    assert np.isfinite(raw[:2]).all() and np.isnan(raw[2:]).all()
    got = G.cmle_sweep_from_stats(Gm, Cm, zz, n, T, Ks)
    for k, nll in zip(Ks, got):
        eig = G.cmle_from_stats(Gm, Cm, zz, n, T, onlylogLike=True, K=int(k))["negloglik"]
        assert abs(nll - eig) <= TOL * abs(eig)


@pytest.mark.parametrize("n,K,T,s", [(4000, 12, 5, 0.0), (20000, 60, 20, 0.0), (3000, 40, 1, 0.0), (5000, 25, 8, 0.5),
                                     (3000, 8, 12, 0.0), (3000, 10, 10, 0.2)])
def test_cMLEimat_matches_oracle(n, K, T, s):
    import oracle as O
    from autofrk_b200 import frk as G

    F, Z = _data(n, K, T, seed=K)
    ref = O.cMLEimat(F, Z, s, wSave=True)            # the literal sequence of n-sized products
    got = G.cMLEimat(F, Z, s, wSave=True)
    assert abs(got["v"] - ref["v"]) <= TOL * abs(ref["v"])
    assert abs(got["negloglik"] - ref["negloglik"]) <= TOL * abs(ref["negloglik"])
    within("test_cMLEimat_matches_oracle:got[M]", relmax(got["M"], ref["M"]), TOL)
    within("test_cMLEimat_matches_oracle:got[w]", relmax(got["w"], ref["w"]), 5e-9)           # the reference's own cancellation ~ dhat/v   # observed 4.6e-10 on B200
    # V: the literal formula is numerically unstable; strict against the closed form, loose vs literal
    Gs, Cs, zz = O.gram_stats(F, Z)
    closed = O.cmle_from_stats(Gs, Cs, zz, n, T, s=s, wSave=True)
    within("test_cMLEimat_matches_oracle:got[V]_vs_closed_form", relmax(got["V"], closed["V"]), 1e-9)
    within("test_cMLEimat_matches_oracle:got[V]_vs_literal", relmax(got["V"], ref["V"]), 5e-4)   # observed 3.0e-05 on B200: the literal V cancels
    assert np.array_equal(got["M"], got["M"].T)
    only = G.cMLEimat(F, Z, s)                         # onlylogLike = !wSave
    assert set(only) == {"negloglik"} and only["negloglik"] == got["negloglik"]


def test_getHalf_getEigen_cMLE():
    import oracle as O
    from autofrk_b200 import frk as G

    n, K, T = 6000, 35, 6
    F, Z = _data(n, K, T, seed=9)
    half_ref = O.getHalf(F, F)
    half = G.getHalf(F, F)
    within("test_getHalf_getEigen_cMLE:half", relmax(half, half_ref), TOL)
    iDF = F * np.linspace(0.5, 2.0, n)[:, None]        # a diagonal-D weighting: t(Fk) %*% iDFk path
    within("test_getHalf_getEigen_cMLE:G.getHalfF", relmax(G.getHalf(F, iDF), O.getHalf(F, iDF)), 1e-13)   # observed 1.0e-15 on B200
    e_ref, e = O.getEigen(F.T @ F), G.getEigen(F.T @ F)
    assert relmax(e["value"], e_ref["value"]) < 1e-12 and np.all(np.diff(e["value"]) <= 0)
    # cMLE given half and JSJ (selectBasis' inner call, R/autoFRK.R:306-316)
    t = half_ref @ F.T @ Z
    JSJ = (t @ t.T) / T
    JSJ = (JSJ + JSJ.T) / 2
    trS = np.sum(Z * Z) / T
    r1, g1 = O.cMLE(F, T, trS, half_ref, JSJ), G.cMLE(F, T, trS, half_ref, JSJ)
    assert abs(g1["negloglik"] - r1["negloglik"]) <= TOL * abs(r1["negloglik"])
    r2, g2 = O.cMLE(F, T, trS, half_ref, JSJ, wSave=True), G.cMLE(F, T, trS, half_ref, JSJ, wSave=True)
    assert abs(g2["v"] - r2["v"]) <= TOL * abs(r2["v"]) and relmax(g2["M"], r2["M"]) < TOL
    assert g2["L"].shape == r2["L"].shape
    within("test_getHalf_getEigen_cMLE:g2[L]g2[L].TF[", relmax(g2["L"] @ g2["L"].T @ F[:, :3], r2["L"] @ r2["L"].T @ F[:, :3]), 1e-13)   # L up to column signs   # observed 2.5e-15 on B200
    r3, g3 = O.cMLE(F, T, trS, half_ref, JSJ, vfixed=0.7, wSave=True), G.cMLE(F, T, trS, half_ref, JSJ, vfixed=0.7, wSave=True)
    assert g3["v"] == 0.7 and relmax(g3["M"], r3["M"]) < TOL


def test_sol_v_branches_through_cMLE():
    """the piecewise rule for v (R/autoFRK.R:335-351), driven through crafted diagonal JSJ matrices."""
    import oracle as O
    from autofrk_b200 import frk as G

    rng = np.random.default_rng(4)
    for n, d, trS in [(50, [0.5, 0.2, 0.1], 100.0),      # max(d) < trS/n  -> first branch
                      (50, [30.0, 5.0, 0.1], 60.0),      # general branch, L in the middle
                      (3, [9.0, 5.0, 2.0], 17.0),        # k == n corner (ks[n] <- n-1, L == n)
                      (40, [8.0, 8.0, 8.0, 1e-9], 30.0)]:
        k = len(d)
        F = rng.standard_normal((n, k))
        JSJ = np.diag(np.array(d, dtype=float))
        half = np.eye(k)
        ref = O.cMLE(F, 2, trS, half, JSJ, wSave=True)
        got = G.cMLE(F, 2, trS, half, JSJ, wSave=True)
        assert abs(got["v"] - ref["v"]) <= 1e-12 * max(1.0, abs(ref["v"])), (n, d)
        assert abs(got["negloglik"] - ref["negloglik"]) <= 1e-12 * abs(ref["negloglik"])
        within("test_sol_v_branches_through_cMLE:got[M]", relmax(got["M"], ref["M"]), 1e-12)


def test_autoFRK_config1_and_predict_FRK():
    """BASELINE configs[0]: autoFRK(Data=z, loc=s) on 2,000 random 2-D locations, 1 replicate, default mrts basis."""
    import oracle as O
    from autofrk_b200 import frk as G

    rng = np.random.default_rng(0)
    n = 2000
    loc = rng.random((n, 2))
    f = np.sin(4 * loc[:, 0]) * np.cos(3 * loc[:, 1]) + 0.5 * loc[:, 0]
    z = (f * 3.0 + rng.standard_normal(n)).reshape(-1, 1)
    ref = O.autoFRK(z, loc)
    got = G.autoFRK(z, loc)
    Kr, Kg = ref["G"].X.shape[1], got["G"].X.shape[1]
    assert np.array_equal(got["G"].Kgrid, ref["G"].extra["Kgrid"]) and len(got["G"].Kgrid) == 30 and got["G"].Kgrid.max() == 446
    within("test_autoFRK_config1_and_predict_FRK:got[G].AIC", relmax(got["G"].AIC, ref["G"].extra["AIC"]), 1e-11)   # observed 1.1e-13 on B200
    assert Kg == Kr
    assert abs(got["s"] - ref["s"]) <= 1e-8 * abs(ref["s"])
    assert abs(got["negloglik"] - ref["negloglik"]) <= 1e-8 * abs(ref["negloglik"])
    # fitted surface is invariant to the eigenvector signs: compare G M G' action and the kriging predictor
    newloc = rng.random((1500, 2))
    pr = O.predict_FRK(ref, newloc=newloc, se_report=True)
    pg = G.predict_FRK(got, newloc=newloc, se_report=True)
    within("test_autoFRK_config1_and_predict_FRK:pg[pred.value]", relmax(pg["pred.value"], pr["pred.value"]), 1e-11)   # observed 4.6e-14 on B200
    within("test_autoFRK_config1_and_predict_FRK:pg[se]", relmax(pg["se"], pr["se"]), 1e-7)   # observed 4.1e-09 on B200
    # newloc = NULL: basis <- object$G
    p0r, p0g = O.predict_FRK(ref), G.predict_FRK(got)
    within("test_autoFRK_config1_and_predict_FRK:p0g[pred.value]", relmax(p0g["pred.value"], p0r["pred.value"]), 1e-11)   # observed 3.9e-14 on B200


def test_autoFRK_multi_replicate_with_given_basis():
    """T = 20 replicates, user-supplied G = mrts(knot, K, loc) (the large-config calling convention)."""
    import oracle as O
    from autofrk_b200 import frk as G, mrts as GM

    rng = np.random.default_rng(1)
    n, T, K = 3000, 20, 40
    loc = rng.random((n, 2))
    knot = loc[:400]
    Fo = O.mrts(knot, K, loc)
    Fg = GM.MrtsBasis(Fo.X, Fo.UZ, Fo.Xu, Fo.nconst, Fo.BBBH)       # same basis on both sides: sign-free comparison
    wt = rng.standard_normal((2, T)) * np.array([[5.0], [3.0]])
    Z = Fo.X[:, 3:5] @ wt + rng.standard_normal((n, T))
    ref = O.autoFRK(Z, loc, G=Fo)
    got = G.autoFRK(Z, loc, G=Fg)
    assert abs(got["s"] - ref["s"]) <= TOL * abs(ref["s"])
    within("test_autoFRK_multi_replicate_with_given_basis:got[M]", relmax(got["M"], ref["M"]), TOL)
    within("test_autoFRK_multi_replicate_with_given_basis:got[w]", relmax(got["w"], ref["w"]), 1e-9)
    newloc = rng.random((2000, 2))
    pr = O.predict_FRK(ref, newloc=newloc, se_report=True)
    pg = G.predict_FRK(got, newloc=newloc, se_report=True)
    within("test_autoFRK_multi_replicate_with_given_basis:pg[pred.value]", relmax(pg["pred.value"], pr["pred.value"]), 1e-9)
    assert pg["se"].shape == pr["se"].shape == (2000, T)
    # se is checked against the closed-form V (the literal V is unstable)
    Gs, Cs, zz = O.gram_stats(Fo.X, Z)
    closed = O.cmle_from_stats(Gs, Cs, zz, n, T, wSave=True)
    B = O.predict_mrts(Fo, newloc)
    se_closed = np.sqrt(np.maximum(0, np.sum((B @ closed["V"]) * B, axis=1)))
    within("test_autoFRK_multi_replicate_with_given_basis:pg[se][", relmax(pg["se"][:, 0], se_closed), 1e-11)   # observed 5.4e-13 on B200


def test_weighted_gram_and_cMLEsp_diagonal_D():
    """row-weighted reductions (G = F' iD F, ...) and the diagonal-D fit cMLEsp (R/autoFRK.R:528-558)."""
    import oracle as O
    from autofrk_b200 import frk as G

    n, K, T = 9001, 37, 4
    F, Z = _data(n, K, T, seed=21)
    rng = np.random.default_rng(5)
    w = rng.random(n) + 0.25
    Gm, Cm, zz = G.gram_stats(F, Z, weights=w)
    within("test_weighted_gram_and_cMLEsp_diagonal_D:Gm", relmax(Gm, F.T @ (w[:, None] * F)), 1e-12)
    within("test_weighted_gram_and_cMLEsp_diagonal_D:Cm", relmax(Cm, F.T @ (w[:, None] * Z)), 1e-12)
    assert abs(zz - np.sum(w[:, None] * Z * Z)) / zz < 1e-13
    mask = (rng.random(n) > 0.3).astype(float)             # 0/1 weights = NA mask of one replicate (R/autoFRK.R:609-615)
    Gk, Ck, _ = G.gram_stats(F, Z[:, :1], weights=mask)
    o = mask > 0
    assert relmax(Gk, F[o].T @ F[o]) < 1e-12 and relmax(Ck, F[o].T @ Z[o, :1]) < 1e-12
    for n2, K2, T2 in ((7003, 75, 3), (12001, 205, 37)):   # tensor-core job kernel (k-split and 16-tile jobs), weighted
        F2, Z2 = _data(n2, K2, T2, seed=K2)
        w2 = rng.random(n2) + 0.25
        G2, C2, zz2 = G.gram_stats(F2, Z2, weights=w2)
        assert relmax(G2, F2.T @ (w2[:, None] * F2)) < 1e-12 and relmax(C2, F2.T @ (w2[:, None] * Z2)) < 1e-12
        assert abs(zz2 - np.sum(w2[:, None] * Z2 * Z2)) / zz2 < 1e-13
    D = 1.0 / w
    ref = O.cMLEsp(F, Z, D, wSave=True)
    got = G.cMLEsp(F, Z, D, wSave=True)
    assert abs(got["s"] - ref["s"]) <= TOL * abs(ref["s"])
    assert abs(got["negloglik"] - ref["negloglik"]) <= TOL * abs(ref["negloglik"])
    within("test_weighted_gram_and_cMLEsp_diagonal_D:got[M]", relmax(got["M"], ref["M"]), TOL)
    within("test_weighted_gram_and_cMLEsp_diagonal_D:got[w]", relmax(got["w"], ref["w"]), 1e-9)
    within("test_weighted_gram_and_cMLEsp_diagonal_D:got[V]", relmax(got["V"], ref["V"]), 2e-4)                # literal V: catastrophic cancellation   # observed 7.1e-06 on B200
    # indeMLE routes a non-constant diagonal D to cMLEsp and a constant one to cMLEimat (R/autoFRK.R:775-806)
    r1 = G.indeMLE(Z, F, D=D, wSave=True)
    within("test_weighted_gram_and_cMLEsp_diagonal_D:r1[M]", relmax(r1["M"], ref["M"]), TOL)
    r2, r3 = G.indeMLE(Z, F, D=np.full(n, 2.0), wSave=True), G.indeMLE(Z, F, wSave=True)
    assert np.array_equal(r2["M"], r3["M"])


def test_invCz_ZinvC_getLikelihood():
    """the Woodbury helpers named in the north star, against the oracle and against dense solves."""
    import oracle as O
    from autofrk_b200 import frk as G

    rng = np.random.default_rng(17)
    n, K, T = 6001, 14, 3
    L = rng.standard_normal((n, K))
    z = rng.standard_normal((n, T))
    Rd = rng.random(n) + 0.5
    got = G.invCz(Rd, L, z)
    within("test_invCz_ZinvC_getLikelihood:got", relmax(got, O.invCz(Rd, L, z)), TOL)
    small = slice(0, 300)
    dense = np.linalg.solve(np.diag(Rd[small]) + L[small] @ L[small].T, z[small])
    within("test_invCz_ZinvC_getLikelihood:G.invCzRd[small]", relmax(G.invCz(Rd[small], L[small], z[small]), dense), TOL)
    within("test_invCz_ZinvC_getLikelihood:G.invCz0.7", relmax(G.invCz(0.7, L, z[:, :1]), O.invCz(np.full(n, 0.7), L, z[:, :1])), TOL)      # scalar R = s * I
    within("test_invCz_ZinvC_getLikelihood:G.ZinvCRd[small]", relmax(G.ZinvC(Rd[small], L[small], z[small].T), O.ZinvC(np.diag(Rd[small]), L[small], z[small].T)), TOL)
    # getLikelihood: complete data equals cMLEimat's negloglik; with NAs equals the oracle's masked evaluation
    F, Z = _data(4000, 9, 4, seed=3)
    fit = O.cMLEimat(F, Z, 0.0, wSave=True)
    ll = G.getLikelihood(Z, F, fit["M"], fit["v"], np.ones(4000))
    assert abs(ll - fit["negloglik"]) <= 1e-9 * abs(fit["negloglik"])
    Zna = Z.copy()
    Zna[rng.random(Z.shape) < 0.2] = np.nan
    Zna[:, 3] = np.nan
    Zna[7, 3] = Z[7, 3]                                         # a replicate with ONE observation: R/helper.R:48-51 quirk
    Dd = rng.random(4000) + 0.5
    ref = O.getLikelihood(Zna, F, fit["M"], fit["v"], Dd)
    got = G.getLikelihood(Zna, F, fit["M"], fit["v"], Dd)
    assert abs(got - ref) <= 1e-9 * abs(ref), (got, ref)


def test_predict_FRK_with_new_observations():
    """predict.FRK(obsData=, obsloc=) (R/autoFRK.R:1248-1274): kriging with fresh observations at new sites."""
    import oracle as O
    from autofrk_b200 import frk as G, mrts as GM

    rng = np.random.default_rng(2)
    n, K = 2500, 30
    loc = rng.random((n, 2))
    Fo = O.mrts(loc[:300], K, loc)
    Fg = GM.MrtsBasis(Fo.X, Fo.UZ, Fo.Xu, Fo.nconst, Fo.BBBH)
    z = (Fo.X[:, 3:6] @ np.array([4.0, -3.0, 2.0]) + rng.standard_normal(n)).reshape(-1, 1)
    ref, got = O.autoFRK(z, loc, G=Fo), G.autoFRK(z, loc, G=Fg)
    obsloc = rng.random((900, 2))
    Bo = O.predict_mrts(Fo, obsloc)
    obs = Bo[:, 3:6] @ np.array([4.0, -3.0, 2.0]) + rng.standard_normal(900)
    obs[::17] = np.nan                                           # missing new observations are dropped (:1249)
    newloc = rng.random((1200, 2))
    pr = O.predict_FRK(ref, obsData=obs, obsloc=obsloc, newloc=newloc, se_report=True)
    pg = G.predict_FRK(got, obsData=obs, obsloc=obsloc, newloc=newloc, se_report=True)
    within("test_predict_FRK_with_new_observations:pg[pred.value]", relmax(pg["pred.value"], pr["pred.value"]), 1e-10)   # observed 6.9e-12 on B200
    within("test_predict_FRK_with_new_observations:pg[se]", relmax(pg["se"], pr["se"]), 5e-6)   # observed 2.4e-07 on B200
    # obsloc = NULL: new data at the training sites
    obs2 = z.ravel() + 0.1 * rng.standard_normal(n)
    pr2 = O.predict_FRK(ref, obsData=obs2, se_report=True)
    pg2 = G.predict_FRK(got, obsData=obs2, se_report=True)
    within("test_predict_FRK_with_new_observations:pg2[pred.value]", relmax(pg2["pred.value"], pr2["pred.value"]), 2e-9)   # observed 1.0e-10 on B200
    within("test_predict_FRK_with_new_observations:pg2[se]", relmax(pg2["se"], pr2["se"]), 5e-5)              # V = M - (...) cancels on both sides   # observed 2.1e-06 on B200
    with pytest.raises(ValueError, match="Dimensions of obsloc and obsData"):
        G.predict_FRK(got, obsData=obs[:10], obsloc=obsloc)


def test_EM0miss_matches_oracle():
    """the missing-data EM (R/autoFRK.R:560-737) from per-replicate masked statistics."""
    import oracle as O
    from autofrk_b200 import frk as G

    rng = np.random.default_rng(6)
    n, K, T = 2500, 11, 7
    F, Z = _data(n, K, T, seed=31, rank=3)
    Z[rng.random(Z.shape) < 0.3] = np.nan
    D = rng.random(n) + 0.5
    for Dd, vfix in ((np.ones(n), None), (D, None), (np.ones(n), 1.3)):
        ref = O.EM0miss(F, Z, Dd, maxit=50, avgtol=1e-6, wSave=True, vfixed=vfix)
        got = G.EM0miss(F, Z, Dd, maxit=50, avgtol=1e-6, wSave=True, num_report=False, vfixed=vfix)
        assert got["niter"] == ref["niter"]
        assert abs(got["s"] - ref["s"]) <= 1e-9 * abs(ref["s"])
        within("test_EM0miss_matches_oracle:got[M]", relmax(got["M"], ref["M"]), 1e-12)   # observed 1.4e-14 on B200
        within("test_EM0miss_matches_oracle:got[w]", relmax(got["w"], ref["w"]), 1e-12)   # observed 1.6e-14 on B200
        within("test_EM0miss_matches_oracle:got[V]", relmax(got["V"], ref["V"]), 1e-10)   # observed 3.1e-12 on B200
        assert abs(got["negloglik"] - ref["negloglik"]) <= 1e-9 * abs(ref["negloglik"])
    # indeMLE dispatches data with NAs to EM0miss (R/autoFRK.R:829-841); an all-NA replicate keeps a zero column in w
    Z2 = Z.copy()
    Z2[:, 2] = np.nan
    r = G.indeMLE(Z2, F, wSave=True)
    assert r["w"].shape == (K, T) and np.all(r["w"][:, 2] == 0.0)
    Zk = Z2[:, [0, 1, 3, 4, 5, 6]]
    rows = ~np.all(np.isnan(Zk), axis=1)                   # indeMLE drops never-observed rows first (R/autoFRK.R:754-770)
    ref2 = O.EM0miss(F[rows], Zk[rows], np.ones(int(rows.sum())), wSave=True)
    within("test_EM0miss_matches_oracle:r[M]", relmax(r["M"], ref2["M"]), 1e-12)   # observed 4.3e-15 on B200
    assert np.array_equal(r["pinfo"]["pick"], np.nonzero(rows)[0])


def test_knn_imputation_and_autoFRK_with_missing_values():
    """method = "fast" with NAs: kNN mean imputation (R/autoFRK.R:135-148) then the complete-data fit."""
    import oracle as O
    from autofrk_b200 import frk as G

    rng = np.random.default_rng(8)
    for d in (1, 2, 3):
        n, T = 3000, 4
        loc = rng.random((n, d))
        Z = np.sin(3 * loc[:, :1]) @ np.ones((1, T)) + 0.1 * rng.standard_normal((n, T))
        Z[rng.random(Z.shape) < 0.15] = np.nan
        for k in (1, 3, 5):
            ref, got = O.knn_impute(Z, loc, k), G.knn_impute(Z, loc, k)
            assert not np.isnan(got).any()
            assert np.max(np.abs(got - ref)) < 1e-14
    n = 1500
    loc = rng.random((n, 2))
    z = (3.0 * np.sin(4 * loc[:, 0]) * np.cos(3 * loc[:, 1]) + rng.standard_normal(n)).reshape(-1, 1)
    z[rng.random(n) < 0.1] = np.nan
    ref = O.autoFRK(z, loc, maxK=40)
    got = G.autoFRK(z, loc, maxK=40)
    assert got["G"].X.shape == ref["G"].X.shape
    assert abs(got["s"] - ref["s"]) <= 1e-8 * abs(ref["s"])
    pr, pg = O.predict_FRK(ref, newloc=loc[:200]), G.predict_FRK(got, newloc=loc[:200])
    within("test_knn_imputation_and_autoFRK_with_missing_values:pg[pred.value]", relmax(pg["pred.value"], pr["pred.value"]), 1e-11)   # observed 4.5e-13 on B200
    # method = "EM": AIC sweep and final fit through EM0miss
    gotem = G.autoFRK(z, loc, Kseq=[5, 9, 14], method="EM")
    assert gotem["G"].X.shape[1] in (5, 9, 14) and np.isfinite(gotem["negloglik"]) and gotem["w"].shape[1] == 1
    with pytest.raises(_frk_error(), match="fewer observed values"):
        G.knn_impute(np.array([[1.0], [np.nan], [np.nan]]), np.arange(3.0).reshape(-1, 1), 3)


def _frk_error():
    from autofrk_b200._lib import FrkError
    return FrkError


@pytest.mark.parametrize("path", ["mid", "narrow", "tma", "cpasync"])
@pytest.mark.parametrize("n,K,T,weighted", [(100_003, 32, 4, False), (50_000, 5, 1, True), (77_777, 48, 8, False),
                                             (20_001, 17, 0, False), (64_000, 40, 16, True), (999, 3, 2, False),
                                             (70_001, 64, 4, False), (33_333, 100, 1, True), (41_000, 128, 16, False),
                                             (9_000, 128, 32, True), (130, 72, 20, False), (31, 100, 3, True),
                                             (50_000, 96, 0, False), (25_000, 40, 100, False)])
def test_gram_small_K_variants(n, K, T, weighted, path, monkeypatch):
    """The shapes default autoFRK() selects (K <= ~130), through the device entry point: the register-resident kernel
    (gram_mid.cu: 1, 2 and 4 row-stage sets, ragged rows and columns, row weights, zz folded in) and, forced through
    FRK_GRAM_PATH, the kernels it stands in front of (gram_narrow.cu for K <= 48, the job kernel, the cp.async one)."""
    import ctypes as C
    import torch
    from autofrk_b200 import _lib

    monkeypatch.setenv("FRK_GRAM_PATH", path)

    g = torch.Generator(device="cuda").manual_seed(n)
    ld = n + (n & 1)
    F = torch.randn((K, ld), dtype=torch.float64, device="cuda", generator=g).T[:n]
    Z = torch.randn((max(T, 1), ld), dtype=torch.float64, device="cuda", generator=g).T[:n]
    w = (torch.rand(ld, dtype=torch.float64, device="cuda", generator=g) + 0.5) if weighted else None
    plen = _lib.lib.frk_gram_packed_len(K, T)
    packed = torch.full((plen,), float("nan"), dtype=torch.float64, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(_lib.lib.frk_gram_stats_dev(C.c_void_p(F.data_ptr()), n, K, ld, C.c_void_p(Z.data_ptr()), T, ld,
                                           None if w is None else C.c_void_p(w.data_ptr()), C.c_void_p(packed.data_ptr()), st))
    torch.cuda.synchronize()
    Fh, Zh = F.cpu().numpy(), Z.cpu().numpy()
    wh = np.ones(n) if w is None else w[:n].cpu().numpy()
    G = packed[:K * K].cpu().numpy().reshape(K, K, order="F")
    assert relmax(G, Fh.T @ (wh[:, None] * Fh)) < 1e-12 and np.array_equal(G, G.T)
    if T:
        Cm = packed[K * K:K * K + K * T].cpu().numpy().reshape(K, T, order="F")
        within("test_gram_small_K_variants:Cm", relmax(Cm, Fh.T @ (wh[:, None] * Zh[:, :T])), 1e-12)
        zz = np.sum(wh[:, None] * Zh[:, :T] ** 2)
        assert abs(packed[-1].item() - zz) / zz < 1e-12
    else:
        assert packed[-1].item() == 0.0


def test_plan_predict_frk_folded_path_edge_cases():
    """frk_plan_predict_frk (coefficients folded through the projection) against the explicit products
    basis %*% w and sqrt(pmax(0, rowSums((basis %*% V) * basis))) (R/autoFRK.R:1216,1220): indefinite V (the signed sum
    of squares must reproduce the clamped bilinear form), full-rank V, zero V, no V, one location, many replicates,
    chunked pipeline (n2 above one chunk), d = 1 and d = 3."""
    import ctypes as C
    from autofrk_b200 import _lib, mrts as GM

    rng = np.random.default_rng(31)

    def run(basis, x, w, V):
        n2, T = x.shape[0], w.shape[1]
        yhat = np.empty((n2, T), order="F")
        se = None if V is None else np.empty(n2)
        _lib.check(_lib.lib.frk_plan_predict_frk(basis.plan()._h, _lib.ptr(x), n2, _lib.ptr(w), T,
                                                 None if V is None else _lib.ptr(V), _lib.ptr(yhat),
                                                 None if se is None else _lib.ptr(se)))
        return yhat, se

    for d, m, K, n2, T in ((2, 200, 40, 3001, 3), (1, 60, 12, 1, 1), (3, 150, 30, 777, 5), (2, 120, 25, 400_000, 2)):
        basis = GM.mrts(rng.random((m, d)), K)
        x = np.asfortranarray(rng.random((n2, d)))
        F = GM.predict_mrts(basis, x) if n2 <= 5000 else None
        w = np.asfortranarray(rng.standard_normal((K, T)))
        R = rng.standard_normal((K, 3))
        cases = {"psd_rank3": R @ R.T, "full_rank": (lambda A: A @ A.T)(rng.standard_normal((K, K))),
                 "indefinite": R @ np.diag([1.0, -2.0, 0.5]) @ R.T, "zero": np.zeros((K, K)), "none": None}
        for name, V in cases.items():
            Vf = None if V is None else np.asfortranarray(V)
            yhat, se = run(basis, x, w, Vf)
            if F is None:                                         # large n2: check three scattered row blocks
                idx = np.r_[0:500, n2 // 2:n2 // 2 + 500, n2 - 500:n2]
                Fi = GM.predict_mrts(basis, np.asfortranarray(x[idx]))
            else:
                idx, Fi = np.arange(n2), F
            ref_y = Fi @ w
            assert np.abs(yhat[idx] - ref_y).max() <= 1e-11 * max(1.0, np.abs(ref_y).max()), (d, name)
            if V is not None:
                q = np.sum((Fi @ V) * Fi, axis=1)
                ref_se = np.sqrt(np.maximum(0.0, q))
                scale = max(1e-300, np.abs(q).max())
                # compare squares: where the bilinear form is clamped at 0 the reference's own rounding decides the sign
                assert np.abs(se[idx] ** 2 - ref_se ** 2).max() <= 1e-10 * scale, (d, name)
                if name == "zero":
                    assert np.all(se == 0.0)


def test_gram_random_shapes_against_numpy():
    """40 seeded random shapes (49 <= K <= 700, 0 <= T <= 130, ragged n, weights on and off) through the tensor-core
    job kernel: every block pair, every trimmed edge tile and every split work unit must land in the right place."""
    from autofrk_b200 import frk as G

    rng = np.random.default_rng(2024)
    for it in range(40):
        K = int(rng.integers(49, 701))
        T = int(rng.integers(0, 131))
        n = int(rng.integers(100, 3001))
        F = np.asfortranarray(rng.standard_normal((n, K)))
        Z = np.asfortranarray(rng.standard_normal((n, max(T, 1))))
        w = (rng.random(n) + 0.5) if it % 3 == 0 else None
        Fw = F if w is None else w[:, None] * F
        if T == 0:
            Gm, Cm, zz = G.gram_stats(F, weights=w)
            assert Cm is None
        else:
            Gm, Cm, zz = G.gram_stats(F, Z, weights=w)
            assert relmax(Cm, Fw.T @ Z) < 1e-12, (K, T, n)
            ref_zz = np.sum(Z * Z) if w is None else np.sum(w[:, None] * Z * Z)
            assert abs(zz - ref_zz) / ref_zz < 1e-13, (K, T, n)
        assert relmax(Gm, Fw.T @ F) < 1e-12, (K, T, n)
        assert np.array_equal(Gm, Gm.T), (K, T, n)


def test_cMLEimat_scale_equivariance():
    """oracle-free property of the closed-form MLE: Data -> a * Data scales v and M by a^2, w by a, V by a^2 and shifts
    negloglik by n T log(a^2); the basis may be rescaled column-wise (F -> F diag(c)) with M -> C^-1 M C^-1, w -> C^-1 w
    and an unchanged likelihood."""
    from autofrk_b200 import frk as G

    n, K, T = 6000, 30, 4
    F, Z = _data(n, K, T, seed=99)
    base = G.cMLEimat(F, Z, 0.0, wSave=True)
    a = 3.7
    sc = G.cMLEimat(F, a * Z, 0.0, wSave=True)
    assert abs(sc["v"] - a * a * base["v"]) <= 1e-10 * sc["v"]
    assert relmax(sc["M"], a * a * base["M"]) < 1e-10 and relmax(sc["V"], a * a * base["V"]) < 1e-9
    within("test_cMLEimat_scale_equivariance:sc[w]", relmax(sc["w"], a * base["w"]), 1e-10)
    assert abs(sc["negloglik"] - (base["negloglik"] + n * T * np.log(a * a))) <= 1e-10 * abs(sc["negloglik"])
    c = np.random.default_rng(5).random(K) + 0.5
    rs = G.cMLEimat(np.asfortranarray(F * c), Z, 0.0, wSave=True)
    assert abs(rs["v"] - base["v"]) <= 1e-10 * base["v"]
    assert abs(rs["negloglik"] - base["negloglik"]) <= 1e-10 * abs(base["negloglik"])
    within("test_cMLEimat_scale_equivariance:rs[M]", relmax(rs["M"], base["M"] / np.outer(c, c)), 1e-13)   # observed 2.6e-15 on B200
    within("test_cMLEimat_scale_equivariance:rs[w]", relmax(rs["w"], base["w"] / c[:, None]), 1e-13)   # observed 2.1e-15 on B200


def test_predict_FRK_argument_contract():
    """Corners of predict.FRK the reference treats specially (R/autoFRK.R:1214-1245): a model fitted with missing data has
    per-replicate standard errors (empty replicates dropped by the fit make R fail on miss[, tt]); obsloc without obsData leaves yhat
    undefined in R; w / V that do not match the basis are rejected before the device reads them."""
    from autofrk_b200 import frk as G, mrts as GM

    rng = np.random.default_rng(8)
    n, K = 600, 12
    loc = rng.random((n, 2))
    Fg = GM.mrts(loc[:100], K, loc)
    z = (Fg.X[:, 3:5] @ np.array([2.0, -1.0]) + 0.3 * rng.standard_normal(n)).reshape(-1, 1)
    obj = G.autoFRK(z, loc, G=Fg)
    newloc = rng.random((50, 2))
    ok = G.predict_FRK(obj, newloc=newloc, se_report=True)
    assert ok["pred.value"].shape == (50, 1) and ok["se"].shape == (50, 1)
    with pytest.raises(ValueError, match="obsloc was given without obsData"):
        G.predict_FRK(obj, obsloc=newloc)
    bad = dict(obj)
    bad["w"] = obj["w"][:-1]
    with pytest.raises(ValueError, match="Dimensions of basis and object\\$w"):
        G.predict_FRK(bad, newloc=newloc)
    bad = dict(obj)
    bad["V"] = obj["V"][:-1, :-1]
    with pytest.raises(ValueError, match="object\\$V"):
        G.predict_FRK(bad, newloc=newloc, se_report=True)
    miss = dict(obj)
    miss["missing"] = {"miss": np.zeros((n, 1))}
    assert G.predict_FRK(miss, newloc=newloc)["pred.value"].shape == (50, 1)       # yhat is the same formula (:1216)
    pm = G.predict_FRK(miss, newloc=newloc, se_report=True)                        # per-replicate V (:1224-1245)
    assert pm["se"].shape == (50, 1) and np.all(np.isfinite(pm["se"])) and np.all(pm["se"] >= 0)
    short = dict(obj)
    short["w"] = np.hstack([obj["w"], np.zeros((K, 1))])                          # a dropped empty replicate keeps its column in w
    short["missing"] = {"miss": np.zeros((n, 1))}
    with pytest.raises(IndexError, match="subscript out of bounds"):
        G.predict_FRK(short, newloc=newloc, se_report=True)


def test_predict_FRK_per_replicate_se_after_EM_fit():
    """predict.FRK(se.report = TRUE) on a model fitted through EM0miss (R/autoFRK.R:1224-1245): one V_t per replicate
    from the rows observed in that replicate, pinfo$pick and pinfo$D; literal restatement in the oracle."""
    import oracle as O
    from autofrk_b200 import frk as G, mrts as GM

    rng = np.random.default_rng(12)
    n, T, K = 1400, 5, 12
    loc = rng.random((n, 2))
    Fo = O.mrts(loc[:150], K, loc)
    Fg = GM.MrtsBasis(Fo.X, Fo.UZ, Fo.Xu, Fo.nconst, Fo.BBBH)
    Z = Fo.X[:, 3:6] @ (rng.standard_normal((3, T)) * 4.0) + rng.standard_normal((n, T))
    Z[rng.random(Z.shape) < 0.35] = np.nan
    Z[5:40, :] = np.nan                                        # never-observed rows are dropped from the fit (pick)
    ref = O.autoFRK(Z, loc, G=Fo, method="EM")
    kept = int((~np.all(np.isnan(Z), axis=1)).sum())
    assert kept <= n - 35 and ref["missing"]["miss"].shape == (kept, T) and ref["pinfo"]["pick"].size == kept
    got = G.autoFRK(Z, loc, G=Fg, method="EM")
    assert np.array_equal(got["pinfo"]["pick"], ref["pinfo"]["pick"])
    assert np.array_equal(np.asarray(got["missing"]["miss"]) == 1, ref["missing"]["miss"] == 1)
    within("test_predict_FRK_per_replicate_se_after_EM_fit:got[M]", relmax(got["M"], ref["M"]), 1e-13)   # observed 6.3e-16 on B200
    newloc = rng.random((700, 2))
    # the same fitted quantities on both sides: isolates predict.FRK
    same = dict(ref)
    same["G"] = Fg
    pr = O.predict_FRK(ref, newloc=newloc, se_report=True)
    assert pr["se"].shape == (700, T) and np.ptp(pr["se"], axis=1).max() > 0     # the columns really differ
    for kind, pg in (("fused", G.predict_FRK(same, newloc=newloc, se_report=True)),
                     ("basis", G.predict_FRK(same, newloc=newloc, basis=O.predict_mrts(Fo, newloc), se_report=True))):
        within(f"test_predict_FRK_per_replicate_se_after_EM_fit:{kind}[pred.value]",
               relmax(pg["pred.value"], pr["pred.value"]), 1e-13)   # observed 1.4e-15 on B200
        # V_t = M - (...) cancels on both sides, as in test_predict_FRK_with_new_observations
        within(f"test_predict_FRK_per_replicate_se_after_EM_fit:{kind}[se]", relmax(pg["se"], pr["se"]), 5e-6)   # observed 3.7e-07 on B200
    # newloc = NULL: basis = object$G (all rows, observed or not)
    pr0, pg0 = O.predict_FRK(ref, se_report=True), G.predict_FRK(same, se_report=True)
    assert pg0["se"].shape == (n, T)
    within("test_predict_FRK_per_replicate_se_after_EM_fit:default[se]", relmax(pg0["se"], pr0["se"]), 5e-6)   # observed 4.0e-07 on B200
    # and end to end on the product's own fit
    pe = G.predict_FRK(got, newloc=newloc, se_report=True)
    within("test_predict_FRK_per_replicate_se_after_EM_fit:own_fit[se]", relmax(pe["se"], pr["se"]), 5e-6)   # observed 4.3e-07 on B200
    # a non-identity diagonal D enters through pinfo$D
    D = rng.random(n) + 0.5
    refD = O.indeMLE(Z, Fo.X, wSave=True, Ddiag=D)
    refD["G"] = Fo
    sameD = dict(refD)
    sameD["G"] = Fg
    prD, pgD = O.predict_FRK(refD, newloc=newloc, se_report=True), G.predict_FRK(sameD, newloc=newloc, se_report=True)
    within("test_predict_FRK_per_replicate_se_after_EM_fit:D[se]", relmax(pgD["se"], prD["se"]), 2e-6)   # observed 1.0e-07 on B200
