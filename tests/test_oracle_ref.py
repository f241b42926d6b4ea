"""Pins the oracle to the reference's OWN source (CPU only).

`oracle/_ref/libautofrk_ref.so` is /root/reference/src/autoFRK.cpp compiled unmodified behind stand-in headers
(oracle/stub/); `tests/golden/ref_vectors.npz` holds its outputs (tests/golden/make_golden_ref.py).  Checked here:
  * the travelled library reproduces the committed vectors bit for bit (same source, same compiler flags);
  * the plain-C restatement of the scalar kernels equals the reference's arithmetic bit for bit (0 ulp);
  * the numpy restatement is within a few ulp of it (numpy's vectorised log/pow are not glibc's);
  * the numpy restatement of the fit / predict wrappers matches the reference's sequence of operations to rounding
    level (per-column sign for eigenvector columns), its argument checks and its messages.
Tolerances are ~10x the observed error, which is recorded beside each.
"""
import ctypes as C
import json
from pathlib import Path

import numpy as np
import pytest

import oracle as O
from oracle import ref as R
from util import colrel, relmax

GOLD = Path(__file__).resolve().parent / "golden"
V = np.load(GOLD / "ref_vectors.npz")
META = json.loads((GOLD / "ref_vectors.json").read_text())

needs_ref = pytest.mark.skipif(not R.available(), reason="oracle/_ref not built (needs /root/reference at build time)")


def ulps(a, b):
    """max |a-b| in units of the spacing of b (0 where both are exactly equal, incl. both zero)."""
    a, b = np.asarray(a), np.asarray(b)
    diff = np.abs(a - b)
    return float(np.max(np.where(diff == 0, 0.0, diff / np.spacing(np.maximum(np.abs(a), np.abs(b))))))


def sign_align(A, B):
    s = np.sign(np.sum(A * B, axis=0))
    s[s == 0] = 1.0
    return A * s


@pytest.fixture(scope="module")
def clib():
    from oracle.build_oracle import build_oracle

    lib = C.CDLL(str(build_oracle()))
    lib.oracle_tpm_predict.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int64]
    lib.oracle_tpm2.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p]
    return lib


@needs_ref
def test_ref_library_reproduces_the_committed_vectors():
    for c in META["kernel_cases"]:
        t, d = c["tag"], c["d"]
        assert np.array_equal(R.tpm_predict(V[f"{t}_Pnew"], V[f"{t}_P"], d), V[f"{t}_tpm_predict"])
        assert np.array_equal(R.tpm2(V[f"{t}_P"], d), V[f"{t}_tpm2_upper"])
    for c in META["fit_cases"]:
        t, ks = c["tag"], c["kstar"]
        fit = R.mrtsrcpp(V[f"{t}_Xu"], V[f"{t}_xobs_diag"], ks)
        for nm in ("X", "UZ", "BBBH", "nconst"):
            assert np.array_equal(fit[nm], V[f"{t}_{nm}"]), (t, nm)
    e = R.getASCeigens(V["eig_A"])
    assert np.array_equal(e["value"], V["eig_value"]) and np.array_equal(e["vector"], V["eig_vector"])


@pytest.mark.parametrize("case", META["kernel_cases"], ids=lambda c: c["tag"])
def test_scalar_kernels_equal_the_reference_arithmetic(clib, case):
    """src/autoFRK.cpp:60-153.  C restatement: bit-exact.  numpy: observed 2 (d=1), 2-3 (d=2), 0 (d=3) ulp."""
    t, d, m, n2 = case["tag"], case["d"], case["m"], case["n2"]
    P, Pn = np.asfortranarray(V[f"{t}_P"]), np.asfortranarray(V[f"{t}_Pnew"])
    want = V[f"{t}_tpm_predict"]
    L = np.full((n2, m), np.nan, order="F")
    assert clib.oracle_tpm_predict(Pn.ctypes.data, n2, n2, P.ctypes.data, m, d, L.ctypes.data, n2) == 0
    assert np.array_equal(L, want)                                   # 0 ulp, r == 0 entries included
    H = np.full((m, m), np.nan, order="F")
    assert clib.oracle_tpm2(P.ctypes.data, m, d, H.ctypes.data) == 0
    up = V[f"{t}_tpm2_upper"]
    assert np.array_equal(H, up + up.T)                              # :189 symmetrisation
    got = O.tpm_predict(Pn, P, d)
    assert ulps(got, want) <= 8
    assert np.all(got[want == 0] == 0) and np.all(want[:4, :4].diagonal() == 0)   # coincident points: exact zeros
    assert ulps(O.tpm2(P, d), up + up.T) <= 8
    assert np.all(np.tril(up) == 0)                                  # the reference writes the strict upper triangle


@pytest.mark.parametrize("case", META["fit_cases"], ids=lambda c: c["tag"])
def test_fit_restatement_matches_the_reference_sequence(case):
    """src/autoFRK.cpp:160-242 with Jacobi (stand-in) vs LAPACK (oracle) eigenpairs.  Observed: BBBH 2e-15, nconst 0;
    X / UZ columns up to sign 1.1e-12 (d=1), 4.6e-13 (d=2), 5.6e-13 (d=3) and 6.5e-11 for the man/mrts.Rd case, which
    runs at the k* = m-d-1 limit (small rho amplifies rounding in gammas = .../rho, |gammas| up to 1.7e6)."""
    t, d, ks = case["tag"], case["d"], case["kstar"]
    Xu, xd = V[f"{t}_Xu"], V[f"{t}_xobs_diag"]
    fit = O.mrtsrcpp(Xu, xd, ks)
    assert relmax(fit["BBBH"], V[f"{t}_BBBH"]) < 5e-14
    assert relmax(fit["nconst"], V[f"{t}_nconst"]) < 5e-14
    tol = {"d1": 1e-11, "d2": 5e-12, "d3": 5e-12, "rd": 1e-9}[t]
    for nm in ("X", "UZ"):
        want = V[f"{t}_{nm}"]
        assert fit[nm].shape == want.shape
        assert colrel(sign_align(fit[nm], want), want) < tol, nm
    # structure that no rounding touches: first column of ones, the UZ corner blocks (:212-215)
    m = Xu.shape[0]
    UZ = V[f"{t}_UZ"]
    assert np.all(V[f"{t}_X"][:, 0] == 1.0) and UZ[m, ks] == 1.0
    assert np.array_equal(UZ[m + 1:, ks + 1:], xd) and np.all(UZ[:m, ks:] == 0) and np.all(UZ[m:, :ks] == 0)
    assert np.array_equal(fit["UZ"][m:, :], UZ[m:, :])


@pytest.mark.parametrize("case", META["fit_cases"], ids=lambda c: c["tag"])
def test_predict_restatement_with_reference_fit_fed_in(case):
    """src/autoFRK.cpp:287-326 with the reference's own UZ/BBBH as inputs: deterministic, no sign ambiguity.
    Observed column-max-relative error 4.7e-13 (d=1), 2.7e-14 (d=2), 2.0e-14 (d=3), 2.9e-11 (Rd case); predict0 (fit
    redone, per-column sign) 1.5e-12 / 4.2e-13 / 5.6e-13 / 7.7e-11."""
    t, d, ks = case["tag"], case["d"], case["kstar"]
    K = ks + d + 1
    got = O.mrtsrcpp_predict(V[f"{t}_Xu"], V[f"{t}_xobs_diag"], V[f"{t}_xnew"], V[f"{t}_BBBH"], V[f"{t}_UZ"],
                             V[f"{t}_nconst"], K)
    want = V[f"{t}_X1_kK"]
    assert got["X1"].shape == want.shape == (case["n2"], K)
    assert np.all(want[:, ks:] == 0) and np.all(got["X1"][:, ks:] == 0)       # k = K quirk: trailing exact zeros
    tol = {"d1": 5e-12, "d2": 5e-13, "d3": 5e-13, "rd": 5e-10}[t]
    assert colrel(got["X1"][:, :ks], want[:, :ks]) < tol
    assert np.array_equal(got["X"], V[f"{t}_Xu"])                              # returns X = Xu (:320)
    # predict0 = fit + predict in one call (:248-281)
    p0 = O.mrtsrcpp_predict0(V[f"{t}_Xu"], V[f"{t}_xobs_diag"], V[f"{t}_xnew"], ks)
    w0 = V[f"{t}_X1_predict0"]
    assert p0["X1"].shape == w0.shape == (case["n2"], ks)
    tol0 = {"d1": 2e-11, "d2": 5e-12, "d3": 5e-12, "rd": 1e-9}[t]
    assert colrel(sign_align(p0["X1"], w0), w0) < tol0


def test_getASCeigens_restatement():
    """src/autoFRK.cpp:24-32: ascending eigenvalues; observed 5e-15 / 3e-14."""
    e = O.getASCeigens(V["eig_A"])
    assert np.all(np.diff(V["eig_value"]) > 0)
    assert relmax(e["value"], V["eig_value"]) < 1e-13
    assert relmax(sign_align(e["vector"], V["eig_vector"]), V["eig_vector"]) < 1e-12


def test_argument_checks_and_messages_match_the_reference():
    """All 16 Rcpp::stop messages reachable through the entry points, as the compiled reference words them."""
    msg = META["error_messages"]
    assert all(v is not None for v in msg.values())
    rng = np.random.default_rng(5)
    Xu = rng.random((20, 2))
    from oracle.mrts_oracle import xobs_diag_of

    xd = xobs_diag_of(Xu, 20)
    xn = rng.random((7, 2))
    fit = O.mrtsrcpp(Xu, xd, 5)
    calls = {
        "tpm_predict_d": lambda: O.tpm_predict(xn, Xu, 0),
        "tpm_predict_cols": lambda: O.tpm_predict(xn[:, :1], Xu, 2),
        "mrts_k_positive": lambda: O.mrtsrcpp(Xu, xd, 0),
        "mrts_n": lambda: O.mrtsrcpp(Xu[:1], xd, 1),
        "mrts_k_lt_n": lambda: O.mrtsrcpp(Xu, xd, 20),
        "mrtsrcpp_xobs": lambda: O.mrtsrcpp(Xu, xd[:1, :1], 5),
        "predict0_xobs": lambda: O.mrtsrcpp_predict0(Xu, np.eye(3), xn, 5),
        "predict0_xnew": lambda: O.mrtsrcpp_predict0(Xu, xd, rng.random((7, 3)), 5),
        "predict_xnew": lambda: O.mrtsrcpp_predict(Xu, xd, rng.random((7, 1)), fit["BBBH"], fit["UZ"], fit["nconst"], 8),
        "predict_BBBH": lambda: O.mrtsrcpp_predict(Xu, xd, xn, fit["BBBH"][:, :19], fit["UZ"], fit["nconst"], 8),
        "predict_UZ": lambda: O.mrtsrcpp_predict(Xu, xd, xn, fit["BBBH"], fit["UZ"], fit["nconst"], 9),
        "predict_nconst": lambda: O.mrtsrcpp_predict(Xu, xd, xn, fit["BBBH"], fit["UZ"], fit["nconst"][:1], 8),
    }
    for name, f in calls.items():
        with pytest.raises(ValueError) as ei:
            f()
        assert str(ei.value) == msg[name], name


@needs_ref
def test_ref_and_oracle_agree_on_fresh_inputs():
    """Beyond the committed vectors: seeded random problems through both, larger than the fixtures."""
    rng = np.random.default_rng(99)
    from oracle.mrts_oracle import xobs_diag_of

    for d, m, ks, n2 in ((2, 200, 60, 500), (3, 150, 50, 300)):
        Xu, xn = rng.random((m, d)), rng.random((n2, d))
        xd = xobs_diag_of(Xu, m)
        a, b = R.mrtsrcpp(Xu, xd, ks), O.mrtsrcpp(Xu, xd, ks)
        assert colrel(sign_align(b["UZ"][:m, :ks], a["UZ"][:m, :ks]), a["UZ"][:m, :ks]) < 1e-10   # observed 1e-12
        assert ulps(R.tpm_predict(xn, Xu, d), O.tpm_predict(xn, Xu, d)) <= 8
        K = ks + d + 1
        pa = R.mrtsrcpp_predict(Xu, xd, xn, a["BBBH"], a["UZ"], a["nconst"], K)["X1"]
        pb = O.mrtsrcpp_predict(Xu, xd, xn, a["BBBH"], a["UZ"], a["nconst"], K)["X1"]
        assert colrel(pb[:, :ks], pa[:, :ks]) < 1e-12                                             # observed 3e-14
