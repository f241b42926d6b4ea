"""CPU-only self-consistency checks of the oracle (its pin to the reference's own source is tests/test_oracle_ref.py).

Two independent codings (numpy vs plain C), 50-digit arithmetic on a small case, the algebraic identities
of SURVEY.md section 4, and the committed golden vectors."""
import ctypes as C
import json
from pathlib import Path

import numpy as np
import pytest

import oracle as O
from oracle.mrts_oracle import tpm2, tpm_predict, xobs_diag_of
from util import colrel, relmax

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def clib():
    from oracle.build_oracle import build_oracle

    lib = C.CDLL(str(build_oracle()))
    lib.oracle_tpm_predict.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int64]
    lib.oracle_tpm2.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p]
    lib.oracle_predict_x1.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64,
                                      C.c_int, C.c_void_p]
    return lib


@pytest.mark.parametrize("d", [1, 2, 3])
def test_c_and_numpy_restatements_agree(clib, d):
    rng = np.random.default_rng(d)
    m, n2 = 57, 203
    Xu = np.asfortranarray(rng.random((m, d)))
    xn = np.asfortranarray(rng.random((n2, d)))
    xn[:4] = Xu[:4]
    L = np.empty((n2, m), order="F")
    assert clib.oracle_tpm_predict(xn.ctypes.data, n2, n2, Xu.ctypes.data, m, d, L.ctypes.data, n2) == 0
    ref = tpm_predict(xn, Xu, d)
    assert np.max(np.abs(L - ref)) <= 4e-16 * np.max(np.abs(ref))
    assert np.all(L[:4][np.arange(4), np.arange(4)] == 0.0)          # r == 0
    H = np.empty((m, m), order="F")
    assert clib.oracle_tpm2(Xu.ctypes.data, m, d, H.ctypes.data) == 0
    assert np.max(np.abs(H - tpm2(Xu, d))) <= 4e-16 * np.max(np.abs(H))
    assert np.array_equal(H, H.T) and np.all(np.diag(H) == 0)
    kstar = 12
    fit = O.mrtsrcpp(Xu, xobs_diag_of(Xu, m), kstar)
    UZ, BBBH = np.asfortranarray(fit["UZ"]), np.asfortranarray(fit["BBBH"])
    X1 = np.empty((n2, kstar), order="F")
    assert clib.oracle_predict_x1(Xu.ctypes.data, m, d, xn.ctypes.data, n2, BBBH.ctypes.data, UZ.ctypes.data, UZ.shape[0],
                                  kstar, X1.ctypes.data) == 0
    ref1 = O.mrtsrcpp_predict(Xu, None, xn, BBBH, UZ, fit["nconst"], kstar)["X1"]
    assert colrel(X1, ref1) < 1e-11
    assert clib.oracle_tpm_predict(xn.ctypes.data, n2, n2, Xu.ctypes.data, m, 4, L.ctypes.data, n2) == 1


def test_kernel_against_50_digit_arithmetic():
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 50
    rng = np.random.default_rng(0)
    for d in (1, 2, 3):
        Xu, xn = rng.random((6, d)), rng.random((5, d))
        got = tpm_predict(xn, Xu, d)
        for i in range(5):
            for j in range(6):
                r = mp.sqrt(sum((mp.mpf(xn[i, c]) - mp.mpf(Xu[j, c])) ** 2 for c in range(d)))
                exact = {1: r ** 3 / 12, 2: r * r * mp.log(r) / (8 * mp.pi), 3: -r / 8}[d]
                # a few ulp of the O(0.1) kernel scale (near r = 1 the d=2 kernel crosses zero, so not relative)
                assert abs(got[i, j] - float(exact)) <= 5e-17


@pytest.mark.parametrize("d,m,kstar", [(1, 30, 8), (2, 200, 50), (3, 150, 40)])
def test_identities_of_the_reference_algorithm(d, m, kstar):
    """predict at the knots reproduces the fit; eigen-columns orthonormal and orthogonal to [1, Xu]."""
    rng = np.random.default_rng(m)
    Xu = rng.random((m, d))
    xd = xobs_diag_of(Xu, m)
    fit = O.mrtsrcpp(Xu, xd, kstar)
    K = kstar + d + 1
    assert fit["X"].shape == (m, K) and fit["UZ"].shape == (m + d + 1, K) and fit["BBBH"].shape == (d + 1, m)
    at = O.mrtsrcpp_predict(Xu, xd, Xu, fit["BBBH"], fit["UZ"], fit["nconst"], K)
    assert np.all(at["X1"][:, kstar:] == 0.0)                                  # quirk 1: trailing d+1 zero columns
    assert colrel(at["X1"][:, :kstar], fit["X"][:, d + 1:]) < 1e-10
    g = fit["X"][:, d + 1:]
    assert np.max(np.abs(g.T @ g / m - np.eye(kstar))) < 1e-12
    B = np.column_stack([np.ones(m), Xu])
    assert np.max(np.abs(B.T @ g)) < 1e-9
    assert at["X"] is not None and np.array_equal(at["X"], Xu)                 # quirk 4
    assert np.allclose(fit["nconst"], Xu.std(axis=0))                          # population sd (quirk 5)
    p0 = O.mrtsrcpp_predict0(Xu, xd, Xu[:10], kstar)
    assert colrel(p0["X1"], at["X1"][:10, :kstar]) < 1e-12
    # R front-ends
    obj = O.mrts(Xu, K)
    assert colrel(O.predict_mrts(obj, Xu), obj.X) < 1e-10
    with pytest.raises(ValueError, match="k must be less than n"):
        O.mrtsrcpp(Xu, xd, m)
    with pytest.raises(ValueError, match="xobs_diag must be"):
        O.mrtsrcpp(Xu, np.eye(d + 1), kstar)
    with pytest.raises(ValueError, match="UZ dimensions are inconsistent"):
        O.mrtsrcpp_predict(Xu, xd, Xu, fit["BBBH"], fit["UZ"], fit["nconst"], K + 1)


def test_fit_depends_on_data_only_through_sufficient_statistics():
    """literal cMLEimat (R/autoFRK.R:399-480) == the (G, C, zz) restatement the CUDA path evaluates."""
    rng = np.random.default_rng(3)
    n, K, T = 4000, 12, 5
    F = rng.standard_normal((n, K))
    F[:, 0] = 1.0
    wts = np.zeros((K, T))
    wts[:3] = rng.standard_normal((3, T)) * 3
    Z = F @ wts + rng.standard_normal((n, T))
    lit = O.cMLEimat(F, Z, 0.0, wSave=True)
    st = O.cmle_from_stats(*O.gram_stats(F, Z), n, T, wSave=True)
    assert abs(lit["v"] - st["v"]) < 1e-12 * abs(st["v"])
    assert abs(lit["negloglik"] - st["negloglik"]) < 1e-12 * abs(st["negloglik"])
    assert relmax(lit["M"], st["M"]) < 1e-11
    assert relmax(lit["w"], st["w"]) < 1e-9
    assert relmax(lit["V"], st["V"]) < 1e-4              # literal V cancels catastrophically
    # invCz is the Woodbury inverse
    L = rng.standard_normal((50, 4))
    z = rng.standard_normal((50, 2))
    R = np.diag(rng.random(50) + 0.5)
    assert relmax(O.invCz(R, L, z), np.linalg.solve(R + L @ L.T, z)) < 1e-12
    assert relmax(O.invCz(np.diag(R).copy(), L, z), np.linalg.solve(R + L @ L.T, z)) < 1e-12
    assert relmax(O.ZinvC(R, L, z.T), np.linalg.solve(R + L @ L.T, z).T) < 1e-12


def test_sol_v_and_kgrid():
    assert O.sol_v([0.5, 0.2], 0.0, 100.0, 50) == 2.0
    assert O.sol_v([0.5, 0.2], 3.0, 100.0, 50) == 0.0
    v = O.sol_v([30.0, 5.0, 0.1], 0.0, 60.0, 50)
    assert abs(v - (60.0 - 35.0) / 48) < 1e-15      # L = 2
    assert O.sol_v([9.0, 5.0, 2.0], 0.0, 17.0, 3) == max((17.0 - 14.0) / 1 - 0, 0)
    from oracle.frk_oracle import kgrid
    K = kgrid(2, 447)                                     # BASELINE configs[0]: N = 2000 -> klim = 447
    assert len(K) == 30 and K[0] == 3 and K[-1] == 446


def test_getLikelihood_complete_data_equals_neg2llik():
    rng = np.random.default_rng(8)
    n, K, T = 300, 6, 3
    F = rng.standard_normal((n, K))
    Z = F[:, :2] @ rng.standard_normal((2, T)) * 2 + rng.standard_normal((n, T))
    fit = O.cMLEimat(F, Z, 0.0, wSave=True)
    ll = O.getLikelihood(Z, F, fit["M"], fit["v"], np.ones(n))
    assert abs(ll - fit["negloglik"]) < 1e-8 * abs(fit["negloglik"])


def test_golden_vectors():
    """committed fixtures (tests/golden/make_golden.py): regression pins for the oracle, since the reference has none."""
    gold = np.load(ROOT / "tests" / "golden" / "mrts_small.npz")
    meta = json.loads((ROOT / "tests" / "golden" / "mrts_small.json").read_text())
    for case in meta["cases"]:
        tag, d, kstar = case["tag"], case["d"], case["kstar"]
        Xu, xnew = gold[f"{tag}_Xu"], gold[f"{tag}_xnew"]
        fit = O.mrtsrcpp(Xu, xobs_diag_of(Xu, Xu.shape[0]), kstar)
        assert relmax(fit["BBBH"], gold[f"{tag}_BBBH"]) < 1e-12
        assert relmax(fit["nconst"], gold[f"{tag}_nconst"]) < 1e-14
        X1 = O.mrtsrcpp_predict(Xu, None, xnew, gold[f"{tag}_BBBH"], gold[f"{tag}_UZ"], gold[f"{tag}_nconst"], kstar)["X1"]
        assert colrel(X1, gold[f"{tag}_X1"]) < 1e-12


def test_fit_restatement_against_50_digit_arithmetic():
    """the whole C++ mrts() (src/autoFRK.cpp:160-217) and the projection (:316-324) re-evaluated with 50-digit mpmath
    (kernel matrix, normal equations, double centering, symmetric eigensolve, gammas, BBBH, X1) on a tiny problem."""
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 50
    rng = np.random.default_rng(42)
    for d, m, k in [(2, 12, 5), (3, 11, 4), (1, 10, 3)]:
        Xu = rng.random((m, d))
        xd = xobs_diag_of(Xu, m)
        fit = O.mrtsrcpp(Xu, xd, k)
        X = [[mp.mpf(float(v)) for v in row] for row in Xu]

        def phi(a, b):
            r = mp.sqrt(sum((a[c] - b[c]) ** 2 for c in range(d)))
            if d == 1:
                return r ** 3 / 12
            if d == 2:
                return mp.mpf(0) if r == 0 else r * r * mp.log(r) / (8 * mp.pi)
            return -r / 8

        H = mp.matrix(m, m)
        for i in range(m):
            for j in range(m):
                H[i, j] = mp.mpf(0) if i == j else phi(X[i], X[j])
        B = mp.matrix(m, d + 1)
        for i in range(m):
            B[i, 0] = 1
            for c in range(d):
                B[i, c + 1] = X[i][c]
        BBB = mp.inverse(B.T * B) * B.T
        AH = H - (H * B) * BBB
        AHA = AH - BBB.T * (B.T * AH)
        AHA = (AHA + AHA.T) / 2
        ev, EV = mp.eigsy(AHA)
        order = sorted(range(m), key=lambda i: -ev[i])[:k]
        root = mp.sqrt(m)
        gam = mp.matrix(m, k)
        for c, idx in enumerate(order):
            col = [EV[i, idx] for i in range(m)]
            big = max(range(m), key=lambda i: abs(col[i]))
            sgn = 1 if col[big] > 0 else -1                       # the oracle's sign convention (fix_signs)
            for i in range(m):
                gam[i, c] = sgn * col[i]
        rho = [ev[idx] for idx in order]
        gammas = gam - B * (BBB * gam)
        for c in range(k):
            for i in range(m):
                gammas[i, c] = gammas[i, c] / rho[c] * root
        BBBH = BBB * H
        to_np = lambda M: np.array([[float(M[i, j]) for j in range(M.cols)] for i in range(M.rows)])
        assert colrel(fit["X"][:, d + 1:], to_np(gam) * float(root)) < 1e-11
        assert colrel(fit["UZ"][:m, :k], to_np(gammas)) < 1e-10
        assert colrel(fit["BBBH"].T, to_np(BBBH).T) < 1e-11
        # projection at new points
        xn = rng.random((7, d))
        ref = O.mrtsrcpp_predict(Xu, xd, xn, fit["BBBH"], fit["UZ"], fit["nconst"], k)["X1"]
        Pn = mp.matrix(7, m)
        Bn = mp.matrix(7, d + 1)
        for i in range(7):
            xi = [mp.mpf(float(v)) for v in xn[i]]
            Bn[i, 0] = 1
            for c in range(d):
                Bn[i, c + 1] = xi[c]
            for j in range(m):
                Pn[i, j] = phi(xi, X[j])
        X1 = Pn * gammas - Bn * (BBBH * gammas)
        assert colrel(ref, to_np(X1)) < 1e-9


@pytest.mark.parametrize("d", [2, 3])
def test_rigid_motion_invariance_of_the_eigen_columns(d):
    """A property of the mathematical object, independent of any implementation: the thin-plate-spline kernel depends
    on distances only and the double centering removes the affine part, so rotating and translating knots and
    locations together leaves the non-polynomial basis columns unchanged up to sign -- in d = 3 exactly; in d = 2 the
    r^2 log r kernel is rotation- and translation-invariant as well (only scaling adds an affine term)."""
    import oracle as O
    from oracle.mrts_oracle import xobs_diag_of

    rng = np.random.default_rng(41)
    m, kstar, n2 = 80, 25, 200
    Xu = rng.random((m, d))
    xnew = rng.random((n2, d))
    Q, _ = np.linalg.qr(rng.standard_normal((d, d)))
    shift = rng.standard_normal(d)
    Xu2, xnew2 = Xu @ Q + shift, xnew @ Q + shift
    K = kstar + d + 1
    out = []
    for knots, pts in ((Xu, xnew), (Xu2, xnew2)):
        xd = xobs_diag_of(knots, m)
        fit = O.mrtsrcpp(knots, xd, kstar)
        X1 = O.mrtsrcpp_predict(knots, xd, pts, fit["BBBH"], fit["UZ"], fit["nconst"], K)["X1"][:, :kstar]
        out.append((fit["X"][:, d + 1:], X1))
    for A, B in zip(out[0], out[1]):
        s = np.sign(np.sum(A * B, axis=0))
        err = np.max(np.max(np.abs(A * s - B), axis=0) / np.max(np.abs(B), axis=0))
        assert err < 1e-8, err          # eigenvector conditioning (relative gaps ~1e-3) times 1e-13-class perturbations


def test_missing_data_predict_se_against_dense_algebra():
    """The oracle's literal restatement of predict.FRK's per-replicate V (R/autoFRK.R:1224-1245, through invCz's
    Woodbury form) against the dense definition V_t = M - M G_t' (s De_t + G_t M+ G_t')^-1 G_t M at a small size."""
    rng = np.random.default_rng(3)
    n, T, K = 160, 3, 7
    loc = rng.random((n, 2))
    Fo = O.mrts(loc[:40], K, loc)
    Z = Fo.X[:, 3:5] @ rng.standard_normal((2, T)) * 3.0 + rng.standard_normal((n, T))
    Z[rng.random(Z.shape) < 0.3] = np.nan
    Z[:4] = np.nan
    Z = np.hstack([Z, np.full((n, 1), np.nan)])                     # an empty replicate: dropped by the fit, kept in w
    D = rng.random(n) + 0.5
    obj = O.indeMLE(Z, Fo.X, wSave=True, Ddiag=D)
    obj["G"] = Fo
    assert obj["w"].shape == (K, T + 1) and np.all(obj["w"][:, T] == 0.0)
    assert obj["pinfo"]["pick"].size == obj["missing"]["miss"].shape[0] <= n - 4
    with pytest.raises(IndexError):                                  # miss[, tt] past the last kept replicate: R fails too
        O.predict_FRK(obj, newloc=loc[:5], se_report=True)
    obj["w"] = obj["w"][:, :T]
    newloc = rng.random((30, 2))
    out = O.predict_FRK(obj, newloc=newloc, se_report=True)
    B = O.predict_mrts(Fo, newloc)
    M, s = obj["M"], obj["s"]
    lam, U = np.linalg.eigh(M)
    Mp = (U * np.maximum(lam, 0.0)) @ U.T
    pick = obj["pinfo"]["pick"]
    for tt in range(T):
        o = ~(obj["missing"]["miss"][:, tt] == 1)
        Gt = Fo.X[pick][o]
        Cm = np.diag(s * D[pick][o]) + Gt @ Mp @ Gt.T
        V = M - M @ Gt.T @ np.linalg.solve(Cm, Gt @ M)
        se = np.sqrt(np.maximum(0.0, np.sum((B @ V) * B, axis=1)))
        assert np.max(np.abs(out["se"][:, tt] - se)) <= 1e-8 * np.max(se)
    assert np.allclose(out["pred.value"], B @ obj["w"], rtol=0, atol=1e-12)
