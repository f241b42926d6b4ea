"""Host-side mirror of autoFRK's mrts interface on top of the C ABI.

Same names, argument meaning and error behaviour as the reference:
  mrtsrcpp_predict(...)     R/RcppExports.R:16-18  -> src/autoFRK.cpp:287-326
  predict_mrts(obj, newx)   R/autoFRK.R:1439-1470  (predict.mrts)
  MrtsBasis                 the R `mrts` object: basis matrix + attributes UZ, Xu, nconst, BBBH
                            (R/autoFRK.R:1115-1120)
  MrtsPlan                  device-resident state for repeated / sharded evaluation (new; the
                            reference re-uploads nothing because it has no device)
All matrices are float64; numpy arrays are converted to column-major (R layout) before the call.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import FRK_OUT_BASIS, FRK_OUT_X1, check, fmat, lib, ptr


class MrtsBasis:
    """R `mrts` object (R/autoFRK.R:1115-1120): `X` is the n x K basis matrix."""

    def __init__(self, X, UZ, Xu, nconst, BBBH):
        self.X = np.asarray(X, dtype=np.float64)
        self.UZ = None if UZ is None else fmat(UZ, "UZ")
        self.Xu = fmat(Xu, "Xu")
        self.nconst = np.asarray(nconst, dtype=np.float64).ravel()
        self.BBBH = None if BBBH is None else fmat(BBBH, "BBBH")
        self._plans = {}

    @property
    def shape(self):
        return self.X.shape

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.X, dtype=dtype)

    def leading(self, K):
        """`Fk[, 1:K]` with the attributes re-attached (R/autoFRK.R:324-328): UZ stays full width."""
        out = MrtsBasis(self.X[:, :K], self.UZ, self.Xu, self.nconst, self.BBBH)
        return out

    def plan(self) -> "MrtsPlan":
        """Device plan for K = NCOL(object) columns, cached on the object."""
        K = self.X.shape[1]
        if K not in self._plans:
            self._plans[K] = MrtsPlan(self.Xu, self.BBBH, self.UZ, self.nconst, K)
        return self._plans[K]


class MrtsPlan:
    """Knots, eigenvector block and polynomial correction resident on the current CUDA device."""

    def __init__(self, Xu, BBBH, UZ, nconst, k):
        Xu, BBBH, UZ = fmat(Xu, "Xu"), fmat(BBBH, "BBBH"), fmat(UZ, "UZ")
        n, d = Xu.shape
        if BBBH.shape != (d + 1, n):
            raise _lib.FrkError(_lib.FRK_EINVAL, "BBBH must be %d x %d" % (d + 1, n))
        self._h = C.c_void_p()
        nc = None if nconst is None else np.ascontiguousarray(nconst, dtype=np.float64).ravel()
        if nc is not None and nc.size != d:
            raise _lib.FrkError(_lib.FRK_EINVAL, "nconst must have length %d" % d)
        check(lib.frk_plan_create(C.byref(self._h), ptr(Xu), n, d, ptr(BBBH), ptr(UZ), UZ.shape[0], UZ.shape[1],
                                  None if nc is None else ptr(nc), int(k), 0))
        self.d, self.k = d, int(k)
        info = (C.c_int64 * 8)()
        check(lib.frk_plan_info(self._h, info))
        self.info = dict(zip(("d", "knots", "k", "k_compute", "tile_rows", "slice_cols", "slices", "chunks"),
                             list(info)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib.frk_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launches(self) -> int:
        return int(lib.frk_plan_launches(self._h))

    def predict(self, xnew, mode=FRK_OUT_X1, out=None) -> np.ndarray:
        """Host arrays in, host array out (n2 x k, column-major)."""
        x = fmat(xnew, "xnew")
        if x.shape[1] != self.d:
            raise _lib.FrkError(_lib.FRK_EINVAL, "xnew must have %d columns" % self.d)
        n2 = x.shape[0]
        if out is None:
            out = np.empty((n2, self.k), dtype=np.float64, order="F")
        assert out.flags.f_contiguous and out.shape == (n2, self.k) and out.dtype == np.float64
        check(lib.frk_plan_predict(self._h, ptr(x), n2, ptr(out), mode))
        return out

    def predict_dev(self, x_t, out_t=None, mode=FRK_OUT_X1):
        """torch CUDA tensors: x_t is (n2, d) column-major (e.g. `coords_dxn.T`), returns (n2, k) column-major.
        Asynchronous on torch's current stream."""
        import torch

        n2, d = x_t.shape
        if d != self.d:
            raise _lib.FrkError(_lib.FRK_EINVAL, "xnew must have %d columns" % self.d)
        ldx = _col_major_ld(x_t)
        if out_t is None:
            out_t = torch.empty((self.k, n2), dtype=torch.float64, device=x_t.device).T
        ldo = _col_major_ld(out_t)
        stream = torch.cuda.current_stream(x_t.device).cuda_stream
        check(lib.frk_plan_predict_dev(self._h, C.c_void_p(x_t.data_ptr()), n2, ldx, C.c_void_p(out_t.data_ptr()),
                                       ldo, mode, C.c_void_p(stream)))
        return out_t


def _col_major_ld(t) -> int:
    import torch

    assert t.dtype == torch.float64 and t.is_cuda and t.dim() == 2
    n, c = t.shape
    if n > 1 and t.stride(0) != 1:
        raise ValueError("tensor must be column-major (stride(0) == 1); pass `row_major.T`-style views")
    ld = t.stride(1) if c > 1 else max(n, 1)
    if ld < n:
        raise ValueError("bad leading dimension")
    return int(ld)


def mrtsrcpp_predict(Xu, xobs_diag, xnew, BBBH, UZ, nconst, k):
    """Drop-in for the `.Call` wrapper at R/RcppExports.R:16-18 (src/autoFRK.cpp:287-326).

    Returns the same named list: X (= Xu, :320), UZ, BBBH, nconst (pass-through) and X1."""
    Xu, xnew, BBBH, UZ = fmat(Xu, "Xu"), fmat(xnew, "xnew"), fmat(BBBH, "BBBH"), fmat(UZ, "UZ")
    nconst = np.ascontiguousarray(nconst, dtype=np.float64).ravel()
    xd = None if xobs_diag is None else fmat(xobs_diag, "xobs_diag")
    n, d = Xu.shape
    n2 = xnew.shape[0]
    X1 = np.empty((n2, int(k)), dtype=np.float64, order="F")
    check(lib.frk_mrtsrcpp_predict(ptr(Xu), n, d, None if xd is None else ptr(xd), ptr(xnew), n2, xnew.shape[1],
                                   ptr(BBBH), BBBH.shape[0], BBBH.shape[1], ptr(UZ), UZ.shape[0], UZ.shape[1],
                                   ptr(nconst), nconst.size, int(k), ptr(X1)))
    return {"X": Xu, "UZ": UZ, "BBBH": BBBH, "nconst": nconst, "X1": X1}


def xobs_diag_of(xobs, n):
    """diag(sqrt(n/(n-1)) / apply(xobs, 2, sd), ndims)   (R/autoFRK.R:1080, :1445)."""
    xobs = fmat(xobs)
    return np.diag(np.sqrt(n / (n - 1.0)) / xobs.std(axis=0, ddof=1))


def predict_mrts(obj: MrtsBasis, newx=None, use_plan=True):
    """predict.mrts (R/autoFRK.R:1439-1470).

    `use_plan=False` follows the R code call for call (mrtsrcpp_predict with k = NCOL(object), then
    slicing and cbind on the host); the default evaluates the same n2 x K matrix in one fused
    device pass through the object's cached plan."""
    if newx is None:
        return obj.X                                            # :1440-1442
    Xu = obj.Xu
    n, ndims = Xu.shape
    k = obj.X.shape[1]                                          # :1447
    x0 = np.asarray(newx, dtype=np.float64)
    x0 = fmat(x0.reshape(-1, ndims, order="F") if x0.ndim != 2 else x0)   # :1448
    kstar = k - ndims - 1                                       # :1449
    if kstar <= 0:                                              # :1450-1451, :1467-1468
        shift = Xu.mean(axis=0)
        return np.column_stack([np.ones(x0.shape[0]), (x0 - shift) / obj.nconst])
    if use_plan:
        return obj.plan().predict(x0, mode=FRK_OUT_BASIS)
    xobs_diag = xobs_diag_of(Xu, n)                             # :1445
    X1 = mrtsrcpp_predict(Xu, xobs_diag, x0, obj.BBBH, obj.UZ, obj.nconst, k)["X1"]  # :1453-1459
    X1 = X1[:, :kstar]                                          # :1460
    shift = Xu.mean(axis=0)                                     # :1462
    X2 = np.column_stack([np.ones(x0.shape[0]), (x0 - shift) / obj.nconst])  # :1463-1464
    return np.column_stack([X2, X1])                            # :1466


def _check_xobs_diag(xobs_diag, d):
    xd = fmat(xobs_diag, "xobs_diag")
    return xd


def mrtsrcpp(Xu, xobs_diag, k):
    """Drop-in for R/RcppExports.R:8-10 (src/autoFRK.cpp:224-242): list(X, UZ, BBBH, nconst)."""
    Xu = fmat(Xu, "Xu")
    xd = _check_xobs_diag(xobs_diag, Xu.shape[1])
    n, d = Xu.shape
    k = int(k)
    kk = max(k, 0)
    X = np.empty((n, kk + d + 1), order="F")
    UZ = np.empty((n + d + 1, kk + d + 1), order="F")
    BBBH = np.empty((d + 1, n), order="F")
    nconst = np.empty(d)
    check(lib.frk_mrtsrcpp(ptr(Xu), n, d, ptr(xd), xd.shape[0], xd.shape[1], k, ptr(X), ptr(UZ), ptr(BBBH),
                           ptr(nconst)))
    return {"X": X, "UZ": UZ, "BBBH": BBBH, "nconst": nconst}


def mrtsrcpp_predict0(Xu, xobs_diag, xnew, k):
    """Drop-in for R/RcppExports.R:12-14 (src/autoFRK.cpp:248-281): fit + evaluate, list(X, UZ, BBBH, nconst, X1)."""
    Xu, xnew = fmat(Xu, "Xu"), fmat(xnew, "xnew")
    xd = _check_xobs_diag(xobs_diag, Xu.shape[1])
    n, d = Xu.shape
    n2 = xnew.shape[0]
    k = int(k)
    kk = max(k, 0)
    X = np.empty((n, kk + d + 1), order="F")
    UZ = np.empty((n + d + 1, kk + d + 1), order="F")
    BBBH = np.empty((d + 1, n), order="F")
    nconst = np.empty(d)
    X1 = np.empty((n2, kk), order="F")
    check(lib.frk_mrtsrcpp_predict0(ptr(Xu), n, d, ptr(xd), xd.shape[0], xd.shape[1], ptr(xnew), n2, xnew.shape[1], k,
                                    ptr(X), ptr(UZ), ptr(BBBH), ptr(nconst), ptr(X1)))
    return {"X": X, "UZ": UZ, "BBBH": BBBH, "nconst": nconst, "X1": X1}


def getASCeigens(A):
    """Drop-in for R/RcppExports.R:4-6 (src/autoFRK.cpp:24-32): list(value ascending, vector)."""
    A = fmat(A, "A")
    n = A.shape[0]
    if A.shape[1] != n:
        raise _lib.FrkError(_lib.FRK_EINVAL, "A must be square")
    value = np.empty(n)
    vector = np.empty((n, n), order="F")
    check(lib.frk_getASCeigens(ptr(A), n, ptr(value), ptr(vector)))
    return {"value": value, "vector": vector}


def unique_rows(x):
    """R `unique` on a matrix (R/helper.R:245-247): first occurrences in original order."""
    x = fmat(x)
    _, idx = np.unique(x, axis=0, return_index=True)
    return np.asfortranarray(x[np.sort(idx)])


def mrts(knot, k, x=None, maxknot=5000):
    """mrts() front-end, R/autoFRK.R:1050-1140.  Returns an MrtsBasis (the R `mrts` object)."""
    xobs = fmat(knot, "knot")                                   # :1055-1059
    Xu = unique_rows(xobs)                                      # :1060
    if x is None and Xu.size != xobs.size:                      # :1061-1063
        x = xobs
    n, ndims = Xu.shape
    if k < ndims + 1:
        raise ValueError("k-1 can not be smaller than the number of dimensions!")   # :1067-1069
    if maxknot < n:
        # :1070-1079 calls subKnot (R/helper.R:93-161), which draws with R's RNG (set.seed + sample);
        # knot thinning stays in R -- knots are an input of this library
        raise NotImplementedError("more than maxknot unique knots: thin them with autoFRK:::subKnot in R first")
    xobs_diag = xobs_diag_of(xobs, n)                           # :1080
    kstar = int(k) - ndims - 1
    res = {}
    if x is not None:
        x = fmat(x, "x")                                        # :1082-1086
        if kstar > 0:
            res = mrtsrcpp_predict0(Xu, xobs_diag, x, kstar)    # :1088
        else:                                                   # :1091-1096
            X2 = Xu - Xu.mean(axis=0)
            shift = Xu.mean(axis=0)
            nconst = np.sqrt(np.diag(X2.T @ X2))
            X2 = np.column_stack([np.ones(x.shape[0]), (x - shift) / nconst * np.sqrt(n)])
            res = {"X": X2[:, :k]}
            x = None
    else:
        if kstar > 0:
            res = mrtsrcpp(Xu, xobs_diag, kstar)                # :1101
        else:                                                   # :1103-1107
            X2 = Xu - Xu.mean(axis=0)
            shift = Xu.mean(axis=0)
            nconst = np.sqrt(np.diag(X2.T @ X2))
            X2 = np.column_stack([np.ones(n), (Xu - shift) / nconst * np.sqrt(n)])
            res = {"X": X2[:, :k]}
    nconst = res.get("nconst")
    if nconst is None:                                          # :1111-1114 (undivided norm: quirk 5 of SURVEY 9)
        X2 = Xu - Xu.mean(axis=0)
        nconst = np.sqrt(np.diag(X2.T @ X2))
    obj = MrtsBasis(res["X"], res.get("UZ"), Xu, nconst, res.get("BBBH"))   # :1115-1120
    if x is None:
        return obj                                              # :1121-1122
    shift = Xu.mean(axis=0)                                     # :1124
    X2 = np.column_stack([np.ones(x.shape[0]), (x - shift) / obj.nconst])   # :1125-1126
    obj0 = np.column_stack([X2, res["X1"]]) if kstar > 0 else X2            # :1127-1131
    return MrtsBasis(obj0, obj.UZ, Xu, obj.nconst, obj.BBBH)
