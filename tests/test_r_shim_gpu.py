"""The R `.Call` shim EXECUTED on the GPU box against a mock of R's C API (tests/r_stub/mock_r.c): every registered
routine is called through its registered name and arity with R-shaped arguments (REALSXP matrices with dims), and the
returned list layouts / values are compared with the ctypes path and with the reference-source vectors.  What this
leaves unverified is only the link against a real libR."""
import json
from pathlib import Path

import numpy as np
import pytest

from r_shim_util import MockR
from util import TOL, colrel, relmax, within

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"
V = np.load(GOLD / "ref_vectors.npz")
META = json.loads((GOLD / "ref_vectors.json").read_text())


def sign_align(A, B):
    s = np.sign(np.sum(A * B, axis=0))
    s[s == 0] = 1.0
    return A * s


@pytest.fixture(scope="module")
def R():
    r = MockR()
    yield r
    r.free()


def test_four_reference_routines_through_dotcall(R):
    """_autoFRK_mrtsrcpp / _predict0 / _predict / _getASCeigens with the list members of src/autoFRK.cpp:238-241,
    275-279, 320-324, 30-31, checked against the vectors produced by the reference's own source."""
    t, d, ks = "d2", 2, 40
    Xu, xd, xnew = V[f"{t}_Xu"], V[f"{t}_xobs_diag"], V[f"{t}_xnew"]
    m, K = Xu.shape[0], ks + d + 1
    fit = R.call("_autoFRK_mrtsrcpp", R.mat(Xu), R.mat(xd), R.ivec([ks]))
    assert list(fit) == ["X", "UZ", "BBBH", "nconst"]
    assert fit["X"].shape == (m, K) and fit["UZ"].shape == (m + d + 1, K) and fit["BBBH"].shape == (d + 1, m) and fit["nconst"].shape == (d,)
    assert colrel(sign_align(fit["UZ"], V[f"{t}_UZ"]), V[f"{t}_UZ"]) < TOL
    assert relmax(fit["BBBH"], V[f"{t}_BBBH"]) < 1e-12 and relmax(fit["nconst"], V[f"{t}_nconst"]) < 1e-13
    # predict with the reference's fit fed in; k = K as predict.mrts passes it (R/autoFRK.R:1458)
    p = R.call("_autoFRK_mrtsrcpp_predict", R.mat(Xu), R.mat(xd), R.mat(xnew), R.mat(V[f"{t}_BBBH"]), R.mat(V[f"{t}_UZ"]),
               R.vec(V[f"{t}_nconst"]), R.ivec([K]))
    assert list(p) == ["X", "UZ", "BBBH", "nconst", "X1"]
    assert np.array_equal(p["X"], Xu)                                          # X = Xu (src/autoFRK.cpp:320)
    assert p["X1"].shape == (xnew.shape[0], K) and np.all(p["X1"][:, ks:] == 0.0)
    assert colrel(p["X1"][:, :ks], V[f"{t}_X1_kK"][:, :ks]) < TOL
    p0 = R.call("_autoFRK_mrtsrcpp_predict0", R.mat(Xu), R.mat(xd), R.mat(xnew), R.ivec([ks]))
    assert list(p0) == ["X", "UZ", "BBBH", "nconst", "X1"] and p0["X1"].shape == (xnew.shape[0], ks)
    assert colrel(sign_align(p0["X1"], V[f"{t}_X1_predict0"]), V[f"{t}_X1_predict0"]) < TOL
    e = R.call("_autoFRK_getASCeigens", R.mat(V["eig_A"]))
    assert list(e) == ["value", "vector"] and relmax(e["value"], V["eig_value"]) < 1e-13


def test_errors_leave_through_rf_error_with_the_reference_messages(R):
    """Argument errors reach R as an error condition carrying the compiled reference's text (tests/golden/ref_vectors.json)."""
    msg = META["error_messages"]
    rng = np.random.default_rng(5)
    Xu = rng.random((20, 2))
    xd = np.diag(1.0 / Xu.std(axis=0))
    xn = rng.random((7, 2))
    fit = R.call("_autoFRK_mrtsrcpp", R.mat(Xu), R.mat(xd), R.ivec([5]))
    cases = {
        "mrts_k_lt_n": lambda: R.call("_autoFRK_mrtsrcpp", R.mat(Xu), R.mat(xd), R.ivec([20])),
        "mrtsrcpp_xobs": lambda: R.call("_autoFRK_mrtsrcpp", R.mat(Xu), R.mat(xd[:1, :1]), R.ivec([5])),
        "predict0_xnew": lambda: R.call("_autoFRK_mrtsrcpp_predict0", R.mat(Xu), R.mat(xd), R.mat(rng.random((7, 3))), R.ivec([5])),
        "predict_BBBH": lambda: R.call("_autoFRK_mrtsrcpp_predict", R.mat(Xu), R.mat(xd), R.mat(xn), R.mat(fit["BBBH"][:, :19]),
                                       R.mat(fit["UZ"]), R.vec(fit["nconst"]), R.ivec([8])),
        "predict_UZ": lambda: R.call("_autoFRK_mrtsrcpp_predict", R.mat(Xu), R.mat(xd), R.mat(xn), R.mat(fit["BBBH"]),
                                     R.mat(fit["UZ"]), R.vec(fit["nconst"]), R.ivec([9])),
        "predict_nconst": lambda: R.call("_autoFRK_mrtsrcpp_predict", R.mat(Xu), R.mat(xd), R.mat(xn), R.mat(fit["BBBH"]),
                                         R.mat(fit["UZ"]), R.vec(fit["nconst"][:1]), R.ivec([8])),
    }
    for name, f in cases.items():
        with pytest.raises(RuntimeError) as ei:
            f()
        assert msg[name] in str(ei.value), (name, str(ei.value))
    # the overlay routines release their plan before raising: a failing call can be repeated indefinitely
    for _ in range(3):
        with pytest.raises(RuntimeError, match="xnew must have 2 columns"):
            R.call("_autoFRK_predict_mrts_basis", R.mat(Xu), R.mat(fit["BBBH"]), R.mat(fit["UZ"]), R.vec(fit["nconst"]), R.ivec([8]),
                   R.mat(rng.random((7, 3))))


def test_overlay_routines_match_the_python_mirror(R):
    """The routines behind r/R/overrides.R (predict.mrts body, Gram statistics, cMLEimat from statistics, the AIC sweep,
    fused predict.FRK) return what the ctypes mirror returns for the same inputs."""
    from autofrk_b200 import frk as GF, mrts as GM

    rng = np.random.default_rng(11)
    d, m, K, n = 2, 150, 40, 5000
    knot = rng.random((m, d))
    obj = GM.mrts(knot, K)                               # carries the attributes mrts() attaches (R/autoFRK.R:1115-1120)
    attrs = (R.mat(obj.Xu), R.mat(obj.BBBH), R.mat(obj.UZ), R.vec(obj.nconst), R.ivec([K]))
    x = rng.random((n, d))
    B = R.call("_autoFRK_predict_mrts_basis", *attrs, R.mat(x))
    Bpy = GM.predict_mrts(obj, x)
    assert B.shape == (n, K) and np.array_equal(B, Bpy)
    Z = B[:, :5] @ rng.standard_normal((5, 3)) + rng.standard_normal((n, 3))
    st = R.call("_autoFRK_gram_stats", R.mat(B), R.mat(Z))
    assert list(st) == ["G", "C", "zz"]
    Gm, Cm, zz = GF.gram_stats(B, Z)
    assert np.array_equal(st["G"], Gm) and np.array_equal(st["C"], Cm) and st["zz"][0] == zz
    fused = R.call("_autoFRK_predict_gram", *attrs, R.mat(x), R.mat(Z))
    assert relmax(fused["G"], Gm) < 1e-12 and relmax(fused["C"], Cm) < 1e-12 and abs(fused["zz"][0] - zz) < 1e-12 * zz
    half = R.call("_autoFRK_getHalf_stats", R.mat(Gm))
    assert relmax(half, GF.getHalf(B, B)) < 1e-12
    fit = R.call("_autoFRK_cMLEimat_stats", R.mat(Gm), R.mat(Cm), R.vec([zz]), R.vec([n]), R.vec([0.0]), R.ivec([1], logical=True))
    ref = GF.cMLEimat(B, Z, 0.0, wSave=True)
    assert list(fit) == ["v", "M", "s", "negloglik", "w", "V"]
    assert abs(fit["v"][0] - ref["v"]) < 1e-13 * abs(ref["v"]) and relmax(fit["M"], ref["M"]) < 1e-12
    assert relmax(fit["w"], ref["w"]) < 1e-12 and relmax(fit["V"], ref["V"]) < 1e-10
    Ks = [5, 12, 23, 40]
    nll = R.call("_autoFRK_cMLEimat_sweep", R.mat(Gm), R.mat(Cm), R.vec([zz]), R.ivec([3]), R.vec([n]), R.vec([0.0]), R.ivec(Ks))
    for k, v in zip(Ks, nll):
        want = GF.cMLEimat(np.asfortranarray(B[:, :k]), Z, 0.0, wSave=False)["negloglik"]
        assert abs(v - want) < 1e-10 * abs(want), (k, v, want)
    pr = R.call("_autoFRK_predict_frk", *attrs, R.mat(x[:777]), R.mat(ref["w"]), R.mat(ref["V"]))
    assert list(pr) == ["yhat", "se"] and pr["yhat"].shape == (777, 3) and pr["se"].shape == (777,)
    want = GF.predict_FRK({"w": ref["w"], "V": ref["V"], "G": obj}, newloc=x[:777], se_report=True)
    assert np.array_equal(pr["yhat"], want["pred.value"]) and np.array_equal(pr["se"], want["se"][:, 0])
    pr2 = R.call("_autoFRK_predict_frk", *attrs, R.mat(x[:10]), R.mat(ref["w"]), R.nil())
    assert pr2["se"] is None and np.array_equal(pr2["yhat"], pr["yhat"][:10])


def test_masked_grams_and_krige_V_behind_the_per_replicate_se(R):
    """The two routines the overlay's predict.FRK uses for a model fitted with missing data (R/autoFRK.R:1224-1245):
    every replicate's t(G_t) solve(De_t) G_t from one pass, and V_t = M - t(G_t M) invCz(s De_t, L_t, G_t M) as K x K
    algebra -- against the dense definitions."""
    rng = np.random.default_rng(21)
    n, K, T = 2500, 9, 4
    F = rng.standard_normal((n, K))
    obs = rng.random((n, T)) > 0.3
    D = rng.random(n) + 0.5
    GG = R.call("_autoFRK_masked_grams", R.mat(F), R.mat(obs.astype(float)), R.vec(D))
    assert GG.shape == (K * K, T)
    for t in range(T):
        o = obs[:, t]
        want = (F[o].T / D[o]) @ F[o]
        assert relmax(GG[:, t].reshape((K, K), order="F"), want) < 1e-12
    A = rng.standard_normal((K, K))
    M = A @ A.T / K + 0.1 * np.eye(K)
    s = 0.7
    Vt = R.call("_autoFRK_krige_V", R.mat(GG[:, 0].reshape((K, K), order="F") / s), R.mat(M))
    o = obs[:, 0]
    G0 = F[o]
    Cm = np.diag(s * D[o]) + G0 @ M @ G0.T                      # M is positive definite here: M+ = M
    want = M - M @ G0.T @ np.linalg.solve(Cm, G0 @ M)
    # V = M - (...) cancels (|V| ~ s / n against |M| ~ 1, and t(G) C^-1 G itself is a difference of terms ~ n / s:
    # SURVEY 8c (3)), in the reference's formula as much as here
    assert Vt.shape == (K, K)
    within("test_masked_grams_and_krige_V:V_vs_dense", relmax(Vt, want), 2e-7)   # observed 1.6e-08 on B200
    with pytest.raises(RuntimeError, match="one row per row of Fk"):
        R.call("_autoFRK_masked_grams", R.mat(F), R.mat(obs[:-1].astype(float)), R.vec(D))
    with pytest.raises(RuntimeError, match="must be K x K"):
        R.call("_autoFRK_krige_V", R.mat(M[:, :-1]), R.mat(M))
