"""K2/K3 parity on the GPU: the mrts fit (knot matrix, centering, top-k eigenpairs, assembly)."""
import numpy as np
import pytest

from util import TOL, colrel, relmax, within

pytestmark = pytest.mark.gpu

FIT_CASES = [(2, 300, 100), (2, 500, 197), (3, 250, 80), (1, 40, 10), (2, 1000, 497)]


def _sign_align(A, B):
    """flip columns of A to match B (eigenvector signs are arbitrary: SURVEY.md 8c (2))."""
    s = np.sign(np.sum(A * B, axis=0))
    s[s == 0] = 1.0
    return A * s


@pytest.mark.parametrize("d,m,kstar", FIT_CASES)
def test_mrtsrcpp_fit(d, m, kstar):
    import oracle as O
    from oracle.mrts_oracle import xobs_diag_of, tpm2
    from autofrk_b200 import mrts as G

    rng = np.random.default_rng(100 + m)
    Xu = rng.random((m, d))
    xd = xobs_diag_of(Xu, m)
    ref = O.mrtsrcpp(Xu, xd, kstar)
    got = G.mrtsrcpp(Xu, xd, kstar)
    K = kstar + d + 1
    assert got["X"].shape == (m, K) and got["UZ"].shape == (m + d + 1, K) and got["BBBH"].shape == (d + 1, m)
    # deterministic pieces: 1e-10 or better
    within("test_mrtsrcpp_fit:got[nconst]", relmax(got["nconst"], ref["nconst"]), 1e-13)
    within("test_mrtsrcpp_fit:got[BBBH].T", colrel(got["BBBH"].T, ref["BBBH"].T), TOL)
    within("test_mrtsrcpp_fit:got[X][", colrel(got["X"][:, :d + 1], ref["X"][:, :d + 1]), 1e-13)
    assert np.array_equal(got["UZ"][m:, :], ref["UZ"][m:, :])          # unit + xobs_diag block, exact
    assert np.all(got["UZ"][:m, kstar:] == 0.0)
    # eigen-columns: per column up to sign at 1e-8 (LAPACK-vs-LAPACK differences scale with 1/eigengap)
    gam_ref = ref["X"][:, d + 1:]
    gam = _sign_align(got["X"][:, d + 1:], gam_ref)
    within("test_mrtsrcpp_fit:gam", colrel(gam, gam_ref), 1e-8)
    # invariants at 1e-10: orthonormality, orthogonality to [1, Xu], eigen-residual against the oracle's AHA
    g = got["X"][:, d + 1:] / np.sqrt(m)
    assert np.max(np.abs(g.T @ g - np.eye(kstar))) < TOL
    B = np.column_stack([np.ones(m), Xu])
    assert np.max(np.abs(B.T @ g)) < TOL * np.sqrt(m)
    H = tpm2(Xu, d)
    P = B @ np.linalg.solve(B.T @ B, B.T)
    AHA = (np.eye(m) - P) @ H @ (np.eye(m) - P)
    rho = np.sum(g * (AHA @ g), axis=0)
    resid = np.max(np.abs(AHA @ g - g * rho)) / np.max(np.abs(rho))
    assert resid < TOL
    assert np.all(np.diff(rho) <= 1e-12 * np.max(np.abs(rho)))        # descending
    assert np.all(rho > 0)
    # gammas block consistent with the library's own eigenvectors: UZ = (I-P) gamma / rho * sqrt(m)
    gs = (g - B @ np.linalg.solve(B.T @ B, B.T @ g)) / rho * np.sqrt(m)
    within("test_mrtsrcpp_fit:got[UZ][m", colrel(got["UZ"][:m, :kstar], gs), 1e-11)   # observed 5.6e-13 on B200


def test_identity_predict_at_knots_equals_fit():
    """known-answer identity: predict(mrts(knot,k), knot)[, d+2:k] == mrts(knot,k)[, d+2:k]."""
    from autofrk_b200 import mrts as G

    rng = np.random.default_rng(3)
    for d, m, K in [(2, 400, 150), (3, 300, 120)]:
        knot = rng.random((m, d))
        obj = G.mrts(knot, K)
        assert obj.X.shape == (m, K)
        at = G.predict_mrts(obj, knot)
        within("test_identity_predict_at_knots_equals_fit:at", colrel(at, obj.X), TOL)
        # mrts(knot, k, x) == predict(mrts(knot, k), x)
        x = rng.random((777, d))
        a = G.mrts(knot, K, x).X
        b = G.predict_mrts(obj, x)
        within("test_identity_predict_at_knots_equals_fit:a", colrel(a, b), TOL)


def test_mrts_frontend_matches_oracle_subspace():
    """mrts() front-end end to end vs the oracle (columns up to sign)."""
    import oracle as O
    from autofrk_b200 import mrts as G

    rng = np.random.default_rng(11)
    d, m, K = 2, 350, 90
    knot = rng.random((m, d))
    knot = np.vstack([knot, knot[:7]])            # duplicated rows: unique() path, x defaults to xobs (:1060-1063)
    x = rng.random((500, d))
    ref = O.mrts(knot, K, x)
    got = G.mrts(knot, K, x)
    assert got.X.shape == ref.X.shape == (500, K)
    within("test_mrts_frontend_matches_oracle_subspace:got.X[", colrel(got.X[:, :d + 1], ref.X[:, :d + 1]), 1e-13)
    within("test_mrts_frontend_matches_oracle_subspace:_sign_aligngot.X[", colrel(_sign_align(got.X[:, d + 1:], ref.X[:, d + 1:]), ref.X[:, d + 1:]), 1e-10)   # observed 2.9e-12 on B200
    ref0 = O.mrts(knot, K)
    got0 = G.mrts(knot, K)
    assert got0.X.shape == ref0.X.shape == (m + 7, K)   # duplicates: evaluated at all original rows
    # k == d+1: polynomial-only basis, host-side R fallback (:1091-1096)
    p = G.mrts(knot[:m], d + 1)
    assert p.X.shape == (m, d + 1)
    with pytest.raises(ValueError, match="k-1 can not be smaller"):
        G.mrts(knot, d)


def test_fit_error_messages():
    from autofrk_b200 import mrts as G
    from autofrk_b200._lib import FrkError

    Xu = np.random.default_rng(0).random((20, 2))
    xd = np.eye(2)
    with pytest.raises(FrkError, match="k must be positive"):
        G.mrtsrcpp(Xu, xd, 0)
    with pytest.raises(FrkError, match=r"k must be less than n \(got k=20, n=20\)"):
        G.mrtsrcpp(Xu, xd, 20)
    with pytest.raises(FrkError, match="xobs_diag must be 2 x 2"):
        G.mrtsrcpp(Xu, np.eye(3), 5)
    with pytest.raises(FrkError, match="xnew must have 2 columns"):
        G.mrtsrcpp_predict0(Xu, xd, np.zeros((4, 3)), 5)


def test_getASCeigens():
    import oracle as O
    from autofrk_b200 import mrts as G

    rng = np.random.default_rng(5)
    A = rng.standard_normal((300, 120))
    A = A.T @ A
    ref = O.getASCeigens(A)
    got = G.getASCeigens(A)
    within("test_getASCeigens:got[value]", relmax(got["value"], ref["value"]), 1e-12)
    V = got["vector"]
    assert np.max(np.abs(V.T @ V - np.eye(120))) < 1e-12
    assert np.max(np.abs(A @ V - V * got["value"])) / got["value"].max() < 1e-12


@pytest.mark.parametrize("d", [2, 3])
def test_rigid_motion_invariance_on_the_gpu(d):
    """size-independent property (no oracle): rotating + translating knots and 50,000 locations together leaves the
    non-polynomial basis columns unchanged up to sign (distance-only kernel, affine part centred out)."""
    from autofrk_b200 import mrts as G

    rng = np.random.default_rng(43)
    m, K, n2 = 300, 60, 50_000
    Xu = rng.random((m, d))
    x = rng.random((n2, d))
    Q, _ = np.linalg.qr(rng.standard_normal((d, d)))
    shift = rng.standard_normal(d)
    A = G.predict_mrts(G.mrts(Xu, K), x)[:, d + 1:]
    B = G.predict_mrts(G.mrts(Xu @ Q + shift, K), x @ Q + shift)[:, d + 1:]
    s = np.sign(np.sum(A * B, axis=0))
    err = np.max(np.max(np.abs(A * s - B), axis=0) / np.max(np.abs(B), axis=0))
    within("test_rigid_motion_invariance_on_the_gpu:err", err, 1e-10)   # observed 2.3e-12 on B200
