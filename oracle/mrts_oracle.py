"""numpy restatement of autoFRK's mrts path (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Pinned to the reference's own compiled source (oracle/_ref, tests/test_oracle_ref.py) for the functions of
src/autoFRK.cpp; the two R front-ends at the end have no executable reference here (no R interpreter).

Follows, line by line:
  src/autoFRK.cpp:60-103    tpm2
  src/autoFRK.cpp:109-153   tpm_predict
  src/autoFRK.cpp:160-217   mrts (C++ core)
  src/autoFRK.cpp:39-54     mrtseigencpp  (Spectra Lanczos, tol 1e-10  ->  LAPACK dsyevr on the same
                                           matrix, k largest, descending; eigenvector signs are
                                           arbitrary in both)
  src/autoFRK.cpp:224-326   mrtsrcpp / mrtsrcpp_predict0 / mrtsrcpp_predict
  R/autoFRK.R:1050-1140     mrts()      (R front-end)
  R/autoFRK.R:1439-1470     predict.mrts()
All matrices are float64; layouts are irrelevant to numpy (the C ABI is column-major like R).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import scipy.linalg as sla

_ROW_CHUNK = 1 << 14


def _as2d(a):
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 1:
        a = a.reshape(-1, 1)
    return a


def _phi(r, d):
    """TPS radial kernel of a distance array r. src/autoFRK.cpp:127-128 (d=1), :135-138 (d=2), :145-148 (d=3)."""
    if d == 1:
        return r ** 3 / 12.0                                  # pow(r,3)/12
    if d == 2:
        out = np.zeros_like(r)                                # L stays 0 where r == 0 (:137)
        nz = r != 0
        rn = r[nz]
        out[nz] = rn * rn * np.log(rn) / (8.0 * math.pi)      # r*r*log(r)/(8.0*M_PI)
        return out
    if d == 3:
        return -r / 8.0
    raise ValueError("Unsupported dimension d: %d" % d)        # :114-116


def _dist(P_new, P, d):
    """r_ij exactly as the reference forms it: |dx| (d=1), sqrt(dx^2+dy^2[+dz^2])."""
    if d == 1:
        return np.abs(P_new[:, None, 0] - P[None, :, 0])
    acc = (P_new[:, None, 0] - P[None, :, 0]) ** 2
    for c in range(1, d):
        acc = acc + (P_new[:, None, c] - P[None, :, c]) ** 2
    return np.sqrt(acc)


def tpm_predict(P_new, P, d=None):
    """Phi_new (n2 x m): src/autoFRK.cpp:109-153."""
    P_new, P = _as2d(P_new), _as2d(P)
    d = P.shape[1] if d is None else d
    if d < 1 or d > 3:
        raise ValueError("Unsupported dimension d: %d" % d)
    if P.shape[1] < d or P_new.shape[1] < d:
        raise ValueError("Inputs must have at least %d columns" % d)
    out = np.empty((P_new.shape[0], P.shape[0]))
    for s in range(0, P_new.shape[0], _ROW_CHUNK):
        e = min(s + _ROW_CHUNK, P_new.shape[0])
        out[s:e] = _phi(_dist(P_new[s:e], P, d), d)
    return out


def tpm2(P, d=None):
    """Symmetrised knot kernel matrix H (m x m, zero diagonal): src/autoFRK.cpp:60-103 + :189.

    The reference fills the strict upper triangle and then adds the transpose; the kernel is
    symmetric in (i, j) bit for bit (dx -> -dx only flips a sign that is squared / abs'ed), so
    evaluating the full rectangle and zeroing the diagonal is the same matrix.
    """
    P = _as2d(P)
    d = P.shape[1] if d is None else d
    H = tpm_predict(P, P, d)
    np.fill_diagonal(H, 0.0)
    H = np.triu(H, 1)
    return H + H.T


def fix_signs(V):
    """Canonical eigenvector signs (largest-|.| entry of every column positive).  The reference's
    signs are whatever Spectra/LAPACK return -- arbitrary -- so tests compare up to this."""
    V = np.array(V, copy=True)
    idx = np.argmax(np.abs(V), axis=0)
    s = np.sign(V[idx, np.arange(V.shape[1])])
    s[s == 0] = 1.0
    return V * s


def mrtseigen(M, k):
    """k largest eigenpairs, descending (src/autoFRK.cpp:39-54)."""
    n = M.shape[0]
    w, V = sla.eigh(M, subset_by_index=[n - k, n - 1], driver="evr")
    return w[::-1].copy(), fix_signs(V[:, ::-1])


def mrts_core(Xu, xobs_diag, k):
    """C++ mrts(): src/autoFRK.cpp:160-217.  Returns H, X, UZ, BBB, nconst."""
    Xu = _as2d(Xu)
    n, d = Xu.shape
    if k <= 0:
        raise ValueError("k must be positive")
    if n <= 1:
        raise ValueError("n must be greater than 1")
    if k >= n:
        raise ValueError("k must be less than n (got k=%d, n=%d)" % (k, n))
    root = math.sqrt(n)
    H = tpm2(Xu, d)                                            # :183-189
    B = np.ones((n, d + 1))
    B[:, 1:] = Xu                                              # :190
    Bt = B.T
    BtB = Bt @ B                                               # :192-193
    BBB = sla.cho_solve(sla.cho_factor(BtB, lower=True), Bt)   # :195-196  (d+1) x n
    AH = H - (H @ B) @ BBB                                     # :197
    AHA = AH - BBB.T @ (Bt @ AH)                               # :198
    rho, gamma = mrtseigen(AHA, k)                             # :200-202
    gammas = (gamma - B @ (BBB @ gamma)) / rho[None, :] * root  # :205
    X_temp = Xu - Xu.mean(axis=0)                              # :206
    nconst = np.sqrt((X_temp ** 2).sum(axis=0))                # :207
    X = np.ones((n, k + d + 1))
    X[:, 1:d + 1] = X_temp * (root / nconst)[None, :]          # :209
    X[:, d + 1:] = gamma * root                                # :210
    UZ = np.zeros((n + d + 1, k + d + 1))                      # :212
    UZ[:n, :k] = gammas                                        # :213
    UZ[n, k] = 1.0                                             # :214
    UZ[n + 1:, k + 1:] = np.asarray(xobs_diag, dtype=np.float64).reshape(d, d)  # :215
    nconst = nconst / root                                     # :216
    return H, X, UZ, BBB, nconst


def _check_diag(xobs_diag, d):
    xd = np.asarray(xobs_diag, dtype=np.float64)
    if xd.ndim != 2 or xd.shape != (d, d):
        raise ValueError("xobs_diag must be %d x %d" % (d, d))  # :231-233


def mrtsrcpp(Xu, xobs_diag, k):
    """src/autoFRK.cpp:224-242."""
    Xu = _as2d(Xu)
    _check_diag(xobs_diag, Xu.shape[1])
    H, X, UZ, BBB, nconst = mrts_core(Xu, xobs_diag, k)
    return {"X": X, "UZ": UZ, "BBBH": BBB @ H, "nconst": nconst}


def _project(xnew, Xu, BBBH, UZ, k):
    """X1 = Phi_new * UZ[0:n,0:k] - [1,xnew] * (BBBH * UZ[0:n,0:k]): src/autoFRK.cpp:311-324 / :266-279."""
    n, d = Xu.shape
    n2 = xnew.shape[0]
    W = UZ[:n, :k]
    Cpoly = BBBH @ W
    X1 = np.empty((n2, k))
    for s in range(0, n2, _ROW_CHUNK):
        e = min(s + _ROW_CHUNK, n2)
        Bn = np.ones((e - s, d + 1))
        Bn[:, 1:] = xnew[s:e]
        X1[s:e] = tpm_predict(xnew[s:e], Xu, d) @ W - Bn @ Cpoly
    return X1


def mrtsrcpp_predict0(Xu, xobs_diag, xnew, k):
    """src/autoFRK.cpp:248-281."""
    Xu, xnew = _as2d(Xu), _as2d(xnew)
    d = Xu.shape[1]
    _check_diag(xobs_diag, d)
    if xnew.shape[1] != d:
        raise ValueError("xnew must have %d columns" % d)      # :260-262
    H, X, UZ, BBB, nconst = mrts_core(Xu, xobs_diag, k)
    BBBH = BBB @ H
    return {"X": X, "UZ": UZ, "BBBH": BBBH, "nconst": nconst,
            "X1": _project(xnew, Xu, BBBH, UZ, k)}


def mrtsrcpp_predict(Xu, xobs_diag, xnew, BBBH, UZ, nconst, k):
    """src/autoFRK.cpp:287-326.  NB returns X = Xu (the knots), :320."""
    Xu, xnew = _as2d(Xu), _as2d(xnew)
    BBBH, UZ = np.asarray(BBBH, dtype=np.float64), np.asarray(UZ, dtype=np.float64)
    nconst = np.asarray(nconst, dtype=np.float64).ravel()
    n, d = Xu.shape
    if xnew.shape[1] != d:
        raise ValueError("xnew must have %d columns" % d)      # :298-300
    if BBBH.shape != (d + 1, n):
        raise ValueError("BBBH must be %d x %d" % (d + 1, n))  # :301-303
    if UZ.shape[0] < n or UZ.shape[1] < k:
        raise ValueError("UZ dimensions are inconsistent with n and k")  # :304-306
    if nconst.size != d:
        raise ValueError("nconst must have length %d" % d)     # :307-309
    return {"X": Xu, "UZ": UZ, "BBBH": BBBH, "nconst": nconst,
            "X1": _project(xnew, Xu, BBBH, UZ, k)}


# --------------------------------------------------------------------------------------------
# R front-ends
# --------------------------------------------------------------------------------------------
@dataclass
class MrtsObject:
    """An R 'mrts' object: the basis matrix plus the attributes set at R/autoFRK.R:1115-1120."""
    X: np.ndarray
    UZ: np.ndarray | None
    Xu: np.ndarray
    nconst: np.ndarray
    BBBH: np.ndarray | None
    extra: dict = field(default_factory=dict)

    @property
    def shape(self):
        return self.X.shape

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.X, dtype=dtype)

    def leading(self, K):
        """Fk[, 1:K] with the mrts attributes re-attached (R/autoFRK.R:324-328)."""
        return MrtsObject(self.X[:, :K], self.UZ, self.Xu, self.nconst, self.BBBH, dict(self.extra))


def unique_rows(x):
    """R `unique` on a matrix: first occurrences, original order (R/helper.R:245-247)."""
    x = _as2d(x)
    _, idx = np.unique(x, axis=0, return_index=True)
    return x[np.sort(idx)]


def xobs_diag_of(xobs, n):
    """diag(sqrt(n/(n-1)) / apply(xobs, 2, sd), ndims): R/autoFRK.R:1080, :1445."""
    xobs = _as2d(xobs)
    sd = xobs.std(axis=0, ddof=1)
    return np.diag(math.sqrt(n / (n - 1.0)) / sd)


def mrts(knot, k, x=None, maxknot=5000):
    """R/autoFRK.R:1050-1140."""
    xobs = _as2d(knot)                                         # :1055-1059
    Xu = unique_rows(xobs)                                     # :1060
    if x is None and Xu.size != xobs.size:                     # :1061-1063
        x = xobs
    n, ndims = Xu.shape
    if k < ndims + 1:
        raise ValueError("k-1 can not be smaller than the number of dimensions!")  # :1067-1069
    if maxknot < n:
        # :1070-1079 -> subKnot (R/helper.R:93-161) draws with R's RNG; out of scope
        raise NotImplementedError("subKnot thinning depends on R's RNG; pass <= maxknot knots")
    xobs_diag = xobs_diag_of(xobs, n)                          # :1080
    res = {}
    if x is not None:
        x = _as2d(x)                                           # :1082-1086
        if k - ndims - 1 > 0:
            res = mrtsrcpp_predict0(Xu, xobs_diag, x, k - ndims - 1)  # :1088
        else:
            X2 = Xu - Xu.mean(axis=0)                          # :1091-1096
            shift = Xu.mean(axis=0)
            nconst = np.sqrt(np.diag(X2.T @ X2))
            X2 = np.column_stack([np.ones(x.shape[0]), (x - shift) / nconst * math.sqrt(n)])
            res = {"X": X2[:, :k]}
            x = None
    else:
        if k - ndims - 1 > 0:
            res = mrtsrcpp(Xu, xobs_diag, k - ndims - 1)       # :1101
        else:
            X2 = Xu - Xu.mean(axis=0)                          # :1103-1107
            shift = Xu.mean(axis=0)
            nconst = np.sqrt(np.diag(X2.T @ X2))
            X2 = np.column_stack([np.ones(n), (Xu - shift) / nconst * math.sqrt(n)])
            res = {"X": X2[:, :k]}
    nconst = res.get("nconst")
    if nconst is None:                                         # :1111-1114
        X2 = Xu - Xu.mean(axis=0)
        nconst = np.sqrt(np.diag(X2.T @ X2))
    obj = MrtsObject(res["X"], res.get("UZ"), Xu, np.asarray(nconst), res.get("BBBH"))
    if x is None:
        return obj                                             # :1121-1122
    shift = Xu.mean(axis=0)                                    # :1124
    X2 = np.column_stack([np.ones(x.shape[0]), (x - shift) / obj.nconst])  # :1125-1126
    obj0 = np.column_stack([X2, res["X1"]]) if k - ndims - 1 > 0 else X2   # :1127-1131
    return MrtsObject(obj0, obj.UZ, Xu, obj.nconst, obj.BBBH)


def predict_mrts(obj: MrtsObject, newx=None):
    """R/autoFRK.R:1439-1470."""
    if newx is None:
        return obj.X
    Xu = obj.Xu
    n, ndims = Xu.shape
    xobs_diag = xobs_diag_of(Xu, n)                            # :1445
    k = obj.X.shape[1]                                         # :1447
    x0 = np.asarray(newx, dtype=np.float64).reshape(-1, ndims, order="F") \
        if np.ndim(newx) != 2 else _as2d(newx)                 # :1448
    kstar = k - ndims - 1
    X1 = None
    if kstar > 0:
        X1 = mrtsrcpp_predict(Xu, xobs_diag, x0, obj.BBBH, obj.UZ, obj.nconst, k)["X1"]  # :1453-1459 (k = K !)
        X1 = X1[:, :kstar]                                     # :1460
    shift = Xu.mean(axis=0)                                    # :1462
    X2 = np.column_stack([np.ones(x0.shape[0]), (x0 - shift) / obj.nconst])  # :1463-1464
    return np.column_stack([X2, X1]) if kstar > 0 else X2
