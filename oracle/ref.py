"""ctypes binding of oracle/_ref/libautofrk_ref.so -- the reference's own src/autoFRK.cpp, compiled unmodified
behind the stand-in headers of oracle/stub/ (see oracle/stub/frk_stub_eigen.h for what that pins).

TEST INFRASTRUCTURE, NOT PRODUCT CODE: only tests/, tests/golden/make_golden_ref.py and bench.py's CPU legs
load it.  The library is built by oracle/build_oracle.py (or `make -C oracle ref`) where /root/reference exists;
the built file travels to the GPU box, the reference sources do not.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

REF_LIB = Path(__file__).resolve().parent / "_ref" / "libautofrk_ref.so"

_lib = None


class RefError(RuntimeError):
    """The message the reference handed to Rcpp::stop (an R error in the real package)."""


def available() -> bool:
    return REF_LIB.exists()


def lib():
    global _lib
    if _lib is None:
        if not REF_LIB.exists():
            raise FileNotFoundError(f"{REF_LIB} is not built (oracle/build_oracle.py needs /root/reference)")
        L = C.CDLL(str(REF_LIB))
        L.ref_last_error.restype = C.c_char_p
        L.ref_source_path.restype = C.c_char_p
        L.ref_list_name.restype = C.c_char_p
        L.ref_list_size.restype = C.c_int64
        for f in ("ref_getASCeigens", "ref_mrtsrcpp", "ref_mrtsrcpp_predict0", "ref_mrtsrcpp_predict"):
            getattr(L, f).restype = C.c_void_p
        L.ref_list_free.argtypes = [C.c_void_p]
        L.ref_list_size.argtypes = [C.c_void_p]
        L.ref_list_name.argtypes = [C.c_void_p, C.c_int64]
        L.ref_list_dims.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.ref_list_copy.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
        _lib = L
    return _lib


def _f(a):
    return np.asfortranarray(np.asarray(a, dtype=np.float64))


def _mat(a):
    a = _f(a)
    if a.ndim == 1:
        a = a.reshape(-1, 1, order="F")
    return a


def _p(a):
    return C.c_void_p(a.ctypes.data)


def _i64(v):
    return C.c_int64(int(v))


def _take(h):
    L = lib()
    if not h:
        raise RefError(L.ref_last_error().decode())
    try:
        out = {}
        for i in range(L.ref_list_size(h)):
            r, c = C.c_int64(), C.c_int64()
            L.ref_list_dims(h, i, C.byref(r), C.byref(c))
            m = np.empty((r.value, c.value), order="F")
            L.ref_list_copy(h, i, _p(m))
            name = L.ref_list_name(h, i).decode()
            out[name] = m[:, 0].copy() if name in ("nconst", "value") else m
        return out
    finally:
        L.ref_list_free(h)


def _check(rc):
    if rc:
        raise RefError(lib().ref_last_error().decode())


def tpm2(P, d):
    """src/autoFRK.cpp:60-103 on a zero matrix: strict upper triangle only (mrts() symmetrises at :189)."""
    P = _mat(P)
    out = np.empty((P.shape[0], P.shape[0]), order="F")
    _check(lib().ref_tpm2(_p(P), _i64(P.shape[0]), C.c_int(P.shape[1]), C.c_int(d), _p(out)))
    return out


def tpm_predict(P_new, P, d):
    """src/autoFRK.cpp:109-153 on `Hnew = Zero(n2, n)` (:311)."""
    P_new, P = _mat(P_new), _mat(P)
    if P_new.shape[1] != P.shape[1]:
        raise ValueError("use tpm_predict_shaped to exercise the reference's own shape checks")
    out = np.empty((P_new.shape[0], P.shape[0]), order="F")
    _check(lib().ref_tpm_predict(_p(P_new), _i64(P_new.shape[0]), _p(P), _i64(P.shape[0]), C.c_int(P.shape[1]),
                                 C.c_int(d), _p(out)))
    return out


def tpm_predict_shaped(P_new, P, L, d):
    P_new, P, L = _mat(P_new), _mat(P), _mat(L)
    _check(lib().ref_tpm_predict_shaped(_p(P_new), _i64(P_new.shape[0]), _i64(P_new.shape[1]), _p(P), _i64(P.shape[0]),
                                        _i64(P.shape[1]), C.c_int(d), _p(L), _i64(L.shape[0]), _i64(L.shape[1])))
    return L


def tpm2_shaped(P, L, d):
    P, L = _mat(P), _mat(L)
    _check(lib().ref_tpm2_shaped(_p(P), _i64(P.shape[0]), _i64(P.shape[1]), C.c_int(d), _p(L), _i64(L.shape[0]),
                                 _i64(L.shape[1])))
    return L


def getASCeigens(A):
    A = _mat(A)
    return _take(lib().ref_getASCeigens(_p(A), _i64(A.shape[0])))


def mrtsrcpp(Xu, xobs_diag, k):
    Xu, xd = _mat(Xu), _mat(xobs_diag)
    return _take(lib().ref_mrtsrcpp(_p(Xu), _i64(Xu.shape[0]), _i64(Xu.shape[1]), _p(xd), _i64(xd.shape[0]),
                                    _i64(xd.shape[1]), C.c_int(k)))


def mrtsrcpp_predict0(Xu, xobs_diag, xnew, k):
    Xu, xd, xn = _mat(Xu), _mat(xobs_diag), _mat(xnew)
    return _take(lib().ref_mrtsrcpp_predict0(_p(Xu), _i64(Xu.shape[0]), _i64(Xu.shape[1]), _p(xd), _i64(xd.shape[0]),
                                             _i64(xd.shape[1]), _p(xn), _i64(xn.shape[0]), _i64(xn.shape[1]), C.c_int(k)))


def mrtsrcpp_predict(Xu, xobs_diag, xnew, BBBH, UZ, nconst, k):
    Xu, xd, xn, B, U = _mat(Xu), _mat(xobs_diag), _mat(xnew), _mat(BBBH), _mat(UZ)
    nc = _f(nconst).reshape(-1)
    return _take(lib().ref_mrtsrcpp_predict(
        _p(Xu), _i64(Xu.shape[0]), _i64(Xu.shape[1]), _p(xd), _i64(xd.shape[0]), _i64(xd.shape[1]),
        _p(xn), _i64(xn.shape[0]), _i64(xn.shape[1]), _p(B), _i64(B.shape[0]), _i64(B.shape[1]),
        _p(U), _i64(U.shape[0]), _i64(U.shape[1]), _p(nc), _i64(nc.shape[0]), C.c_int(k)))
