"""Host-side mirror of autoFRK's fixed-rank-kriging fit / predict interface on top of the C ABI.

Same names and argument meaning as the reference's R functions; every n-sized product runs on the
device (K4 Gram reductions, K1 basis evaluation) and every K x K eigen-stage in K5:
  gram_stats        new: G = F'F, C = F'Z, zz  (R/helper.R:20, R/autoFRK.R:431-435)
  getASCeigens      src/autoFRK.cpp:24-32        getEigen / getHalf   R/helper.R:12-26
  cMLE              R/autoFRK.R:333-397          cMLEimat             R/autoFRK.R:399-480
  selectBasis       R/autoFRK.R:199-331          indeMLE              R/autoFRK.R:739-843
  autoFRK           R/autoFRK.R:121-197          predict_FRK          R/autoFRK.R:1168-1423
The LatticeKrig fine-scale branch (cMLElk, SURVEY.md section 2 #20: out of scope) raises NotImplementedError.  Data with
NAs run through the masked statistics + EM0miss / kNN imputation entry points built in round 1;
for such a model `predict_FRK(se_report=True)` returns the per-replicate standard errors of R/autoFRK.R:1224-1245,
computed from per-replicate masked Grams (_missing_se_covariances).
Nothing falls back to host arithmetic.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _lib
from ._lib import check, fmat, lib, ptr
from .mrts import MrtsBasis, getASCeigens, mrts, predict_mrts


def gram_stats(Fk, Data=None, weights=None):
    """(G, C, zz) = (F'WF, F'WZ, sum(w Z^2)) in one device pass over the rows (W = diag(weights), default I)."""
    F = fmat(Fk, "Fk")
    n, K = F.shape
    w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64).ravel()
    if w is not None and w.size != n:
        raise _lib.FrkError(_lib.FRK_EINVAL, "weights must have one entry per row of Fk")
    Z = None if Data is None else fmat(Data, "Data")
    T = 0 if Z is None else Z.shape[1]
    if Z is not None and Z.shape[0] != n:
        raise _lib.FrkError(_lib.FRK_EINVAL, "Data must have as many rows as Fk")
    G = np.empty((K, K), order="F")
    Cm = np.empty((K, max(T, 1)), order="F")
    zz = C.c_double(0.0)
    check(lib.frk_gram_stats(ptr(F), n, K, None if Z is None else ptr(Z), T, None if w is None else ptr(w), ptr(G),
                             ptr(Cm), C.byref(zz)))
    return G, (Cm if T else None), zz.value


def getEigen(A):
    """R/helper.R:12-17."""
    obj = getASCeigens(A)
    return {"value": obj["value"][::-1].copy(), "vector": np.asfortranarray(obj["vector"][:, ::-1])}


def getHalf(Fk, iDFk):
    """R/helper.R:19-26: (t(Fk) %*% iDFk)^(-1/2)."""
    F = fmat(Fk, "Fk")
    if iDFk is Fk or (np.shape(iDFk) == F.shape and np.shares_memory(np.asarray(iDFk), np.asarray(Fk))):
        G, _, _ = gram_stats(F)
    else:
        _, G, _ = gram_stats(F, iDFk)       # the K x K cross-product t(Fk) %*% iDFk   (R/helper.R:20)
    return half_from_gram(G)


def half_from_gram(G):
    G = fmat(G, "G")
    K = G.shape[0]
    half = np.empty((K, K), order="F")
    check(lib.frk_getHalf(ptr(G), K, K, ptr(half)))
    return half


def cmle_from_stats(G, Cm, zz, n, TT, s=0.0, wSave=False, onlylogLike=None, K=None, on_device=False):
    """cMLEimat (R/autoFRK.R:399-480) from the sufficient statistics; K selects the leading K x K / K x T blocks."""
    if onlylogLike is None:
        onlylogLike = not wSave
    if on_device:   # torch CUDA tensors holding column-major K x K / K x T blocks (flat or 2-D views)
        Kfull = int(round(math.sqrt(G.numel())))
        K = Kfull if K is None else int(K)
        pG, pC, ldg, ldc = C.c_void_p(G.data_ptr()), C.c_void_p(Cm.data_ptr()), Kfull, Kfull
    else:
        G, Cm = fmat(G, "G"), fmat(Cm, "C")
        Kfull = G.shape[0]
        K = Kfull if K is None else int(K)
        pG, pC, ldg, ldc = ptr(G), ptr(Cm), Kfull, Kfull
    v, nll = C.c_double(0.0), C.c_double(0.0)
    M = np.empty((K, K), order="F") if not onlylogLike else None
    w = np.empty((K, TT), order="F") if (wSave and not onlylogLike) else None
    V = np.empty((K, K), order="F") if (wSave and not onlylogLike) else None
    check(lib.frk_cMLEimat_stats(pG, ldg, pC, ldc, float(zz), K, int(TT), int(n), float(s), int(bool(wSave)),
                                 int(bool(onlylogLike)), int(bool(on_device)), C.byref(v), C.byref(nll),
                                 None if M is None else ptr(M), None if w is None else ptr(w),
                                 None if V is None else ptr(V)))
    if onlylogLike:
        return {"negloglik": nll.value}
    out = {"v": v.value, "M": M, "s": s, "negloglik": nll.value}
    if wSave:
        out["w"], out["V"] = w, V
    return out


def cmle_sweep_from_stats(G, Cm, zz, n, TT, Ks, s=0.0):
    """negloglik of cMLEimat(Fk[, 1:K], Data, s)$negloglik for every K in Ks (the loop at R/autoFRK.R:305-317) from the
    max-K statistics: one Cholesky factor of G serves all leading blocks (frk_cMLEimat_sweep); a K it cannot serve
    (G not safely positive definite) goes through the eigen path of cmle_from_stats."""
    G, Cm = fmat(G, "G"), fmat(Cm, "C")
    ks = np.ascontiguousarray(np.asarray(Ks, dtype=np.int32))
    out = np.empty(ks.size, dtype=np.float64)
    check(lib.frk_cMLEimat_sweep(ptr(G), G.shape[0], ptr(Cm), Cm.shape[0], float(zz), int(TT), int(n), float(s),
                                 ks.ctypes.data_as(C.c_void_p), int(ks.size), out.ctypes.data_as(C.c_void_p)))
    for i in np.nonzero(np.isnan(out))[0]:
        out[i] = cmle_from_stats(G, Cm, zz, n, TT, s=s, wSave=False, onlylogLike=True, K=int(ks[i]))["negloglik"]
    return out


def cMLEimat(Fk, Data, s, wSave=False, S=None, onlylogLike=None):
    """R/autoFRK.R:399-480."""
    if S is not None:
        raise NotImplementedError("cMLEimat(S=...) (precomputed covariance) has no caller in autoFRK")
    F = fmat(Fk, "Fk")
    Z = fmat(Data, "Data")
    G, Cm, zz = gram_stats(F, Z)
    return cmle_from_stats(G, Cm, zz, F.shape[0], Z.shape[1], s=s, wSave=wSave, onlylogLike=onlylogLike)


def cMLEsp(Fk, Data, Depsilon, wSave=False):
    """R/autoFRK.R:528-558 for a DIAGONAL Depsilon (vector of its diagonal, or a diagonal matrix).

    iD enters only through the weighted statistics G = F' iD F, C = F' iD Z, zz = sum(iD Z^2) (one weighted
    device pass); with half G half = I the wSave products collapse to the same closed forms as cMLEimat."""
    De = np.asarray(Depsilon, dtype=np.float64)
    if De.ndim == 2:
        if np.count_nonzero(De - np.diag(np.diag(De))):
            raise NotImplementedError("non-diagonal Depsilon needs spam's sparse solve, out of scope (SURVEY.md 2 #19)")
        De = np.diag(De)
    F, Z = fmat(Fk, "Fk"), fmat(Data, "Data")
    if De.size != F.shape[0]:
        raise _lib.FrkError(_lib.FRK_EINVAL, "Depsilon must be n x n")
    iD = 1.0 / De                                               # :530
    ldetD = float(np.sum(np.log(De)))                           # :531
    G, Cm, zz = gram_stats(F, Z, weights=iD)                    # :532-538 in one weighted pass
    TT = Z.shape[1]
    out = cmle_from_stats(G, Cm, zz, F.shape[0], TT, s=0.0, wSave=wSave)   # :539-552
    out["negloglik"] = out["negloglik"] + ldetD * TT            # ldet * TT, R/autoFRK.R:379-380,393-396
    if "v" in out:
        out["s"] = out.pop("v")                                 # :554-555
    return out


def cMLE(Fk, TT, trS, half, JSJ=None, s=0.0, ldet=None, wSave=False, onlylogLike=None, vfixed=None):
    """R/autoFRK.R:333-397."""
    if onlylogLike is None:
        onlylogLike = not wSave
    F = fmat(Fk, "Fk")
    n, K = F.shape
    half, JSJ = fmat(half, "half"), fmat(JSJ, "JSJ")
    v, nll, nL = C.c_double(0.0), C.c_double(0.0), C.c_int32(0)
    M = np.empty((K, K), order="F")
    L = np.empty((n, K), order="F") if (wSave and not onlylogLike) else None
    check(lib.frk_cMLE(ptr(F), n, K, int(TT), float(trS), ptr(half), ptr(JSJ), float(s),
                       0.0 if ldet is None else float(ldet), int(bool(wSave)), int(bool(onlylogLike)),
                       int(vfixed is not None), 0.0 if vfixed is None else float(vfixed), C.byref(v), C.byref(nll),
                       ptr(M), None if L is None else ptr(L), C.byref(nL)))
    if onlylogLike:
        return {"negloglik": nll.value}
    return {"v": v.value, "M": M, "s": s, "negloglik": nll.value, "L": None if L is None else L[:, :nL.value]}


def _diag_of(R, n, what):
    """diagonal of a diagonal matrix given as scalar, vector or (diagonal) matrix."""
    R = np.asarray(R, dtype=np.float64)
    if R.ndim == 0:
        return np.full(n, float(R))
    if R.ndim == 2:
        if np.count_nonzero(R - np.diag(np.diag(R))):
            raise NotImplementedError(f"non-diagonal {what} needs spam's sparse solve, out of scope (SURVEY.md 2 #19)")
        R = np.diag(R)
    if R.size != n:
        raise _lib.FrkError(_lib.FRK_EINVAL, f"{what} must be n x n")
    return np.ascontiguousarray(R)


def invCz(R, L, z):
    """R/autoFRK.R:845-852: (R + L L')^{-1} z by Woodbury, R diagonal.  The two Grams t(L) iR L and t(L) iR z are one
    row-weighted device pass; the result is produced chunk by chunk on the device."""
    Lm = fmat(L, "L")
    zm = fmat(z, "z")
    n, K = Lm.shape
    if zm.shape[0] != n:
        raise _lib.FrkError(_lib.FRK_EINVAL, "z must have as many rows as L")
    Rd = _diag_of(R, n, "R")
    out = np.empty((n, zm.shape[1]), order="F")
    check(lib.frk_invCz(ptr(Rd), ptr(Lm), ptr(zm), n, K, zm.shape[1], ptr(out)))
    return out


def ZinvC(R, L, z):
    """R/helper.R:249-256 (dead code upstream): z (R + L L')^{-1} for a row-matrix z; the transpose of invCz."""
    zm = np.atleast_2d(np.asarray(z, dtype=np.float64))
    return np.ascontiguousarray(invCz(R, L, zm.T).T)


def getLikelihood(Data, Fk, M, s, Depsilon):
    """R/helper.R:28-58, diagonal Depsilon: -2 log-likelihood with per-replicate NA masks.  Every replicate's
    masked t(L_t) R_t^-1 L_t and t(L_t) R_t^-1 z_t come from one pass over the rows of Fk with 0/1-over-R weights."""
    Z = fmat(Data, "Data")
    F = fmat(Fk.X if isinstance(Fk, MrtsBasis) else Fk, "Fk")
    n, K = F.shape
    if Z.shape[0] != n:
        raise _lib.FrkError(_lib.FRK_EINVAL, "Data must have as many rows as Fk")
    O = np.asfortranarray(~np.isnan(Z), dtype=np.uint8)                       # :36
    Z0 = np.asfortranarray(np.where(np.isnan(Z), 0.0, Z))
    Dd = _diag_of(Depsilon, n, "Depsilon")
    Mm = fmat(M, "M")
    out = C.c_double(0.0)
    check(lib.frk_getLikelihood(ptr(Z0), O.ctypes.data_as(C.c_void_p), ptr(F), ptr(Mm), float(s), ptr(Dd), n, K, Z.shape[1],
                                C.byref(out)))
    return out.value


def mkpd(M):
    """R/helper.R:278-288 (the eigenvalues come from the device eigensolver)."""
    M = fmat(M, "M")
    v = float(getASCeigens((M + M.T) / 2 if not np.array_equal(M, M.T) else M)["value"][0])
    if v <= 0:
        M = M + np.eye(M.shape[0]) * (max(0.0, -v) + 0.1 ** 7.5)
    return M


def EM0miss(Fk, Data, Depsilon, maxit=50, avgtol=1e-6, wSave=False, external=False, DfromLK=None, num_report=True,
            vfixed=None):
    """R/autoFRK.R:560-737 for diagonal Depsilon and DfromLK = NULL.

    The per-replicate masked products of :587-616 come from one row-weighted pass over Fk; the EM iteration
    (:652-687) is K x K work on the device; the final likelihood is getLikelihood (:690)."""
    if DfromLK is not None:
        raise NotImplementedError("LatticeKrig fine scale is out of scope (SURVEY.md 2 #20)")
    if external:
        raise NotImplementedError("external=TRUE calls the undefined dumpObjects upstream (R/autoFRK.R:619): dead code")
    F = fmat(Fk.X if isinstance(Fk, MrtsBasis) else Fk, "Fk")
    Z = fmat(Data, "Data")
    n, K = F.shape
    TT = Z.shape[1]
    Dd = _diag_of(Depsilon, n, "Depsilon")
    O = np.asfortranarray(~np.isnan(Z), dtype=np.uint8)          # :564
    Z0 = np.asfortranarray(np.where(np.isnan(Z), 0.0, Z))        # :643-644
    G_all = np.empty((K, K, TT), order="F")
    c_all = np.empty((K, TT), order="F")
    zz_all = np.empty(TT)
    nobs = np.empty(TT, dtype=np.int64)
    check(lib.frk_masked_stats(ptr(Z0), O.ctypes.data_as(C.c_void_p), ptr(F), ptr(Dd), n, K, TT, ptr(G_all), ptr(c_all),
                               ptr(zz_all), nobs.ctypes.data_as(C.c_void_p)))
    old = cMLEimat(F, Z0, s=0.0, wSave=True)                     # :645
    s0 = old["v"] if vfixed is None else float(vfixed)           # :646
    M = np.empty((K, K), order="F")
    etatt = np.empty((K, TT), order="F")
    s_out, niter = C.c_double(0.0), C.c_int32(0)
    check(lib.frk_EM0miss_stats(ptr(G_all), ptr(c_all), ptr(zz_all), int(nobs.sum()), K, TT, ptr(fmat(old["M"])), float(s0),
                                int(maxit), float(avgtol), int(vfixed is not None), 0.0 if vfixed is None else float(vfixed),
                                ptr(M), C.byref(s_out), ptr(etatt), C.byref(niter)))
    if num_report:
        print("Number of iteration: ", niter.value)              # :688
    n2loglik = getLikelihood(Z, F, M, s_out.value, Dd)           # :690
    out = {"M": M, "s": s_out.value, "negloglik": n2loglik}
    if wSave:                                                    # :726-735
        out["w"] = etatt
        out["V"] = M - etatt @ etatt.T / TT
        out["missing"] = {"miss": 1 - O, "maxit": maxit, "avgtol": avgtol}
    out["niter"] = niter.value
    return out


def knn_impute(Data, loc, n_neighbor=3):
    """R/autoFRK.R:135-148 (= :273-286): NA -> mean of the n.neighbor nearest observed values of the same replicate."""
    Z = np.array(fmat(Data, "Data"), order="F", copy=True)
    if not np.isnan(Z).any():
        return Z
    L = fmat(loc, "loc")
    if L.shape[0] != Z.shape[0]:
        raise _lib.FrkError(_lib.FRK_EINVAL, "loc must have one row per row of Data")
    check(lib.frk_knn_impute(ptr(L), L.shape[0], L.shape[1], ptr(Z), Z.shape[1], int(n_neighbor)))
    return Z


def kgrid(d, maxK):
    """Default K sequence, R/autoFRK.R:254-257."""
    K = np.unique(np.round(np.arange(d + 1, maxK + 1e-9, maxK ** (1.0 / 3) * d))).astype(int)
    if K.size > 30:
        K = np.unique(np.round(np.linspace(d + 1, maxK, 30))).astype(int)
    return K


def _krige_new_observations(obj, obsData, obsloc, mu_obs, se_report):
    """predict.FRK with obsData (R/autoFRK.R:1197-1202, 1248-1274): replaces object$w (and V) by the kriging
    coefficients for the new observations.  Everything n-sized is a row-weighted Gram pass over the basis at the
    observation sites (fused with its evaluation when obsloc is given); the rest is K x K."""
    G = obj["G"]
    Gmat = G.X if isinstance(G, MrtsBasis) else fmat(G, "G")
    nobs = Gmat.shape[0] if obsloc is None else fmat(obsloc, "obsloc").shape[0]
    z = np.asarray(obsData, dtype=np.float64).reshape(-1, 1) - mu_obs           # :1198
    if z.size != nobs:
        raise ValueError("Dimensions of obsloc and obsData are not compatible!")   # :1199-1201
    pick = np.nonzero(~np.isnan(z.ravel()))[0]                                  # :1249
    M = fmat(obj["M"], "M")
    K = M.shape[0]
    s = float(obj["s"])
    zp = np.asfortranarray(z[pick])
    Gg = np.empty((K, K), order="F")
    c = np.empty((K, 1), order="F")
    if obsloc is None:
        De = np.ones(Gmat.shape[0]) if "pinfo" not in obj or "D" not in obj["pinfo"] else np.asarray(obj["pinfo"]["D"])
        w = 1.0 / (s * De[pick])                                                # object$s * De   :1251,1266
        Gg, c, _ = gram_stats(np.asfortranarray(Gmat[pick]), zp, weights=w)     # G <- object$G[pick, ]  :1252
    else:
        if not isinstance(G, MrtsBasis):
            raise ValueError("Basis matrix of new locations should be given (unless the model was fitted with mrts bases)!")
        xo = np.asfortranarray(fmat(obsloc, "obsloc")[pick])                    # De <- identity  :1255-1256
        w = np.full(pick.size, 1.0 / s)
        zz = C.c_double(0.0)
        check(lib.frk_plan_predict_gram(G.plan()._h, ptr(xo), xo.shape[0], ptr(zp), 1, ptr(w), ptr(Gg), ptr(c), C.byref(zz)))
    coef = np.empty((K, 1), order="F")
    V = np.empty((K, K), order="F") if se_report else None
    check(lib.frk_krige_from_stats(ptr(Gg), ptr(c), 1, ptr(M), K, int(bool(se_report)), ptr(coef),
                                   None if V is None else ptr(V)))
    new = dict(obj)
    new["w"] = coef
    if se_report:
        new["V"] = V
    return new


def selectBasis(Data, loc, D=None, maxit=50, avgtol=1e-6, maxK=None, Kseq=None, method="fast", n_neighbor=3,
                maxknot=5000, DfromLK=None, Fk=None):
    """R/autoFRK.R:199-331 (identity D, complete data, method = "fast").

    The AIC sweep uses ONE max-K pass over the rows: every candidate K works on the leading K columns of
    Fk, i.e. on the leading principal blocks of G = F'F and C = F'Z (R recomputes both for each K)."""
    if DfromLK is not None:
        raise NotImplementedError("LatticeKrig fine scale (cMLElk) is out of scope (SURVEY.md 2 #20)")
    iD = None
    if D is not None:                                           # diagonal D: iD = solve(D) as row weights (:289-291)
        Dd = np.asarray(D, dtype=np.float64)
        if Dd.ndim == 2:
            if np.count_nonzero(Dd - np.diag(np.diag(Dd))):
                raise NotImplementedError("non-diagonal D needs spam's sparse solve, out of scope (SURVEY.md 2 #19)")
            Dd = np.diag(Dd)
        iD = 1.0 / Dd
    if method not in ("fast", "EM"):
        raise ValueError("'arg' should be one of 'fast', 'EM'")   # match.arg(method)
    Z = fmat(Data, "Data")
    loc = fmat(loc, "loc")
    keep_t = ~np.all(np.isnan(Z), axis=0)                       # empty replicates dropped (:203-206)
    Z = Z[:, keep_t]
    pick = np.nonzero(~np.all(np.isnan(Z), axis=1))[0]          # rows never observed dropped (:213-220)
    Z = np.asfortranarray(Z[pick])
    withNA = bool(np.isnan(Z).any())
    d = loc.shape[1]
    N = Z.shape[0]
    klim = int(min(N, np.round(10 * math.sqrt(N))))             # :222
    if N < maxknot:
        knot = loc[pick]                                        # :223-224
    else:
        raise NotImplementedError("N >= maxknot needs subKnot (R RNG), out of scope: pass G = mrts(knot, K, loc)")
    if maxK is not None:
        maxK = int(np.round(maxK))
    elif Kseq is not None:
        maxK = int(np.round(max(Kseq)))
    else:
        maxK = klim                                             # :228-236
    if Kseq is not None:
        K = np.array(sorted(set(int(np.round(k)) for k in Kseq)))
        if K.max() > maxK:
            raise ValueError("maximum of Kseq is larger than maxK!")
        K = K[K > d]
        if K.size == 0:
            raise ValueError("Not valid Kseq!")
    else:
        K = kgrid(d, maxK)                                      # :254-257
    if Fk is None:
        Fk = mrts(knot, int(K.max()), loc, maxknot)             # :260
    if method == "EM":                                          # :264-271: one indeMLE (EM0miss when there are NAs) per K
        Dp = None if D is None else (1.0 / iD)[pick]
        AIC = np.full(K.size, np.inf)
        for i, k in enumerate(K):
            AIC[i] = indeMLE(Z, np.asfortranarray(Fk.X[pick][:, :int(k)]), Dp, maxit, avgtol, wSave=False)["negloglik"]
        TT = Z.shape[1]
        df = np.where(K <= TT, K * (K + 1) / 2 + 1, K * TT + 1 - TT * (TT - 1) / 2)   # :321
        AIC = AIC + 2 * df
        out = Fk.leading(int(K[np.argmin(AIC)]))
        out.AIC, out.Kgrid = AIC, K
        return out
    if withNA:
        Z = knn_impute(Z, loc[pick], n_neighbor)                # :273-286
    TT = Z.shape[1]
    if iD is not None:
        iD = iD[pick]
    G, Cm, zz = gram_stats(np.asfortranarray(Fk.X[pick]), Z, weights=iD)   # one pass; trS = zz/TT (:304)
    AIC = cmle_sweep_from_stats(G, Cm, zz, N, TT, K, s=0.0)    # :305-317, all K from one Cholesky factor
    df = np.where(K <= TT, K * (K + 1) / 2 + 1, K * TT + 1 - TT * (TT - 1) / 2)   # :321
    AIC = AIC + 2 * df                                          # :322
    Kopt = int(K[np.argmin(AIC)])                               # :323
    out = Fk.leading(Kopt)                                      # :324-328
    out.AIC, out.Kgrid = AIC, K
    return out


def indeMLE(Data, Fk, D=None, maxit=50, avgtol=1e-6, wSave=False, DfromLK=None, vfixed=None, sigma_FRK=None):
    """R/autoFRK.R:739-843, complete data: isimat -> cMLEimat (:778-803), other diagonal D -> cMLEsp (:805-813)."""
    if DfromLK is not None:
        raise NotImplementedError("LatticeKrig fine scale (cMLElk) is out of scope (SURVEY.md 2 #20)")
    Z = fmat(Data, "Data")
    F = Fk.X if isinstance(Fk, MrtsBasis) else fmat(Fk, "Fk")
    if np.isnan(Z).any():                                       # withNA -> EM0miss (:829-841)
        keep_t = ~np.all(np.isnan(Z), axis=0)                   # empty replicates are dropped (:746-750) ...
        keep_r = ~np.all(np.isnan(Z[:, keep_t]), axis=1)        # ... and so are rows never observed (:754-770)
        Dd = np.ones(Z.shape[0]) if D is None else _diag_of(D, Z.shape[0], "D")
        out = EM0miss(F[keep_r], Z[np.ix_(keep_r, keep_t)], Dd[keep_r], maxit, avgtol, wSave, False, None, False, vfixed)
        if wSave:
            w = np.zeros((F.shape[1], Z.shape[1]))
            w[:, keep_t] = out["w"]                             # :834-836
            out["w"] = w
            out["pinfo"] = {"D": Dd, "pick": np.nonzero(keep_r)[0]}   # :838
        return out
    if D is not None:
        Dd = np.asarray(D, dtype=np.float64)
        if Dd.ndim == 2:
            if np.count_nonzero(Dd - np.diag(np.diag(Dd))):
                raise NotImplementedError("non-diagonal D needs spam's sparse solve, out of scope (SURVEY.md 2 #19)")
            Dd = np.diag(Dd)
        # isimat: checkDiag(D) and a constant diagonal (R/autoFRK.R:775-776) -- the constant itself is ignored there
        if np.sum(np.abs(np.mean(Dd) - Dd)) >= np.finfo(float).eps:
            out = cMLEsp(F, Z, Dd, wSave)                       # :806
            if wSave:
                out["pinfo"] = {"D": Dd, "pick": np.arange(Z.shape[0])}   # :811
            return out
    sigma = 0.0 if sigma_FRK is None else float(sigma_FRK)      # options(sigma_FRK) :779-783
    out = cMLEimat(F, Z, s=sigma, wSave=wSave)                  # :787
    if "v" in out:                                              # :788-795
        out["s"] = out["v"] if sigma == 0 else sigma
        del out["v"]
    if wSave:
        out["pinfo"] = {"pick": np.arange(Z.shape[0])}          # :800 (D0 = identity)
    return out


def autoFRK(Data, loc, mu=0.0, D=None, G=None, finescale=False, maxit=50, tolerance=1e-6, maxK=None, Kseq=None,
            method="fast", n_neighbor=3, maxknot=5000):
    """R/autoFRK.R:121-197 (finescale = FALSE)."""
    if finescale:
        raise NotImplementedError("finescale=TRUE (LatticeKrig) is out of scope (SURVEY.md 2 #20)")
    Z = fmat(Data, "Data") - mu                                 # :125
    if G is not None:
        Fk = G                                                  # :126-127
    else:
        Fk = selectBasis(Z, loc, D, maxit, tolerance, maxK, Kseq, method, n_neighbor, maxknot, None)  # :129-132
    if method == "fast":
        Z = knn_impute(Z, loc, n_neighbor)                      # :135-148
    obj = indeMLE(Z, Fk, D, maxit, avgtol=tolerance, wSave=True)  # :151
    obj["G"] = Fk                                               # :185
    obj["LKobj"] = None
    return obj


def _missing_se_covariances(obj):
    """R/autoFRK.R:1225-1242: a model fitted through EM0miss gets one V per replicate,
    V_t = M - t(G_t M) invCz(s De_t, L_t, G_t M) over the rows observed in replicate t (pinfo$pick, pinfo$D).
    With R = s De_t diagonal this is K x K algebra on t(G_t) R^-1 G_t: the per-replicate NA-masked weighted Grams
    come from ONE pass over the rows of the basis (frk_masked_stats), each V_t from frk_krige_from_stats."""
    G = obj["G"]
    pinfo = obj.get("pinfo") or {}
    Gall = G.X if isinstance(G, MrtsBasis) else fmat(G, "G")
    pick = np.asarray(pinfo.get("pick", np.arange(Gall.shape[0])))
    D0 = np.ones(Gall.shape[0]) if "D" not in pinfo else np.asarray(pinfo["D"], dtype=np.float64)
    D0 = np.ascontiguousarray(D0[pick])                                          # D0 <- pinfo$D[pick, pick]  :1227
    miss = np.asarray(obj["missing"]["miss"]) == 1                               # :1228
    Fk = np.asfortranarray(Gall[pick])                                           # Fk <- object$G[pick, ]     :1229
    M = fmat(obj["M"], "M")
    K = M.shape[0]
    TT = fmat(obj["w"], "w").shape[1]
    if miss.shape[1] < TT:
        # empty replicates were dropped before EM0miss (:746-750) but object$w keeps their columns (:834-836):
        # miss[, tt] fails in R for the trailing replicates
        raise IndexError("subscript out of bounds (the model was fitted with empty replicates: R/autoFRK.R:1233)")
    if miss.shape[0] != Fk.shape[0] or Fk.shape[1] != K:
        raise ValueError("Dimensions of the missing-data pattern, object$G and object$M are not compatible!")
    s = float(obj["s"])
    if not s > 0:
        raise ValueError("object$s * De is singular (s = 0): solve() fails in invCz (R/autoFRK.R:847)")
    O = np.asfortranarray(~miss[:, :TT], dtype=np.uint8)
    Z0 = np.zeros((Fk.shape[0], TT), order="F")
    G_all = np.empty((K, K, TT), order="F")
    c_all = np.empty((K, TT), order="F")
    zz_all = np.empty(TT)
    nobs = np.empty(TT, dtype=np.int64)
    check(lib.frk_masked_stats(ptr(Z0), O.ctypes.data_as(C.c_void_p), ptr(Fk), ptr(D0), Fk.shape[0], K, TT, ptr(G_all),
                               ptr(c_all), ptr(zz_all), nobs.ctypes.data_as(C.c_void_p)))
    Vs = []
    c0 = np.zeros((K, 1), order="F")
    coef = np.empty((K, 1), order="F")
    for tt in range(TT):                                                         # :1232-1241
        Gg = np.asfortranarray(G_all[:, :, tt] / s)                              # t(G) solve(s De) G
        V = np.empty((K, K), order="F")
        check(lib.frk_krige_from_stats(ptr(Gg), ptr(c0), 1, ptr(M), K, 1, ptr(coef), ptr(V)))
        Vs.append(V)
    return Vs


def _predict_with_missing_se(obj, newloc, basis, mu_new):
    """predict.FRK(se.report = TRUE) for a model with a "missing" attribute (R/autoFRK.R:1216, 1224-1245):
    yhat = basis %*% w as always, se[, t] from the replicate's own V_t.  One fused pass per replicate
    (columns [w_t | factor of V_t] behind the basis evaluation) when the basis is evaluated here."""
    G = obj["G"]
    w = fmat(obj["w"], "w")
    K, TT = w.shape
    Vs = _missing_se_covariances(obj)
    fused = basis is None and newloc is not None
    if fused:
        x = fmat(newloc, "newloc")
        if G.X.shape[1] != K:
            raise ValueError("Dimensions of basis and object$w are not compatible!")
        if G.X.shape[1] - G.Xu.shape[1] - 1 <= 0:
            basis, fused = predict_mrts(G, x), False                              # polynomial-only basis (k* = 0)
    if not fused:
        B = fmat(G.X if isinstance(G, MrtsBasis) else G, "G") if basis is None else fmat(basis, "basis")
        if newloc is not None and B.shape[0] != fmat(newloc).shape[0]:
            raise ValueError("Dimensions of newloc and basis are not compatible!")
        if B.shape[1] != K:
            raise ValueError("Dimensions of basis and object$w are not compatible!")
    n2 = x.shape[0] if fused else B.shape[0]
    yhat = np.empty((n2, TT), order="F")
    se = np.empty((n2, TT), order="F")
    yt, st = np.empty((n2, 1), order="F"), np.empty(n2)
    for tt in range(TT):
        wt = np.asfortranarray(w[:, tt:tt + 1])
        if fused:
            check(lib.frk_plan_predict_frk(G.plan()._h, ptr(x), n2, ptr(wt), 1, ptr(Vs[tt]), ptr(yt), ptr(st)))
        else:
            check(lib.frk_basis_predict_frk(ptr(B), n2, K, ptr(wt), 1, ptr(Vs[tt]), ptr(yt), ptr(st)))
        yhat[:, tt] = yt[:, 0]
        se[:, tt] = st
    return {"pred.value": yhat + mu_new, "se": se}


def predict_FRK(obj, obsData=None, obsloc=None, mu_obs=0.0, newloc=None, basis=None, mu_new=0.0, se_report=False):
    """R/autoFRK.R:1168-1423, LKobj = NULL and obsData = NULL: yhat = basis %*% w (:1216),
    se = sqrt(pmax(0, rowSums((basis %*% V) * basis))) (:1220), both fused behind the basis evaluation."""
    if "w" not in obj:
        raise ValueError('input model (object) should use the option "wsave=TRUE"!')   # :1171-1173
    if obsloc is not None and obsData is None:
        # the reference computes yhat only when both are NULL (:1214) or obsData is given (:1248): this call dies there
        # with "object 'yhat' not found"
        raise ValueError("obsloc was given without obsData: the reference leaves yhat undefined here (R/autoFRK.R:1214,1248)")
    if newloc is None:
        newloc = obsloc
    G = obj["G"]
    if obsData is not None:
        obj = _krige_new_observations(obj, obsData, obsloc, mu_obs, se_report)   # :1248-1274 -> same products below
    if basis is None and newloc is not None and not isinstance(G, MrtsBasis):
        raise ValueError("Basis matrix of new locations should be given (unless the model was fitted with mrts bases)!")
    w = fmat(obj["w"], "w")
    K, TT = w.shape
    if se_report and obsData is None and obj.get("missing") is not None:
        return _predict_with_missing_se(obj, newloc, basis, mu_new)               # :1224-1245
    V = fmat(obj["V"], "V") if se_report else None
    if V is not None and V.shape != (K, K):
        raise ValueError("Dimensions of object$V and object$w are not compatible!")
    se = None
    if basis is None and newloc is not None:
        # basis <- predict.mrts(object$G, newx = newloc) (:1184), fused with the two products on the device
        x = fmat(newloc, "newloc")
        kstar = G.X.shape[1] - G.Xu.shape[1] - 1
        if G.X.shape[1] != K:                                   # the fused call reads K x T of w and K x K of V
            raise ValueError("Dimensions of basis and object$w are not compatible!")
        if kstar > 0:
            n2 = x.shape[0]
            yhat = np.empty((n2, TT), order="F")
            se = np.empty(n2) if se_report else None
            check(lib.frk_plan_predict_frk(G.plan()._h, ptr(x), n2, ptr(w), TT, None if V is None else ptr(V), ptr(yhat),
                                           None if se is None else ptr(se)))
            return {"pred.value": yhat + mu_new, "se": None if se is None else np.repeat(se[:, None], TT, axis=1)}
        basis = predict_mrts(G, x)                              # polynomial-only basis (k* = 0)
    B = fmat(G.X if isinstance(G, MrtsBasis) else G, "G") if basis is None else fmat(basis, "basis")   # :1176,1182
    if newloc is not None and B.shape[0] != fmat(newloc).shape[0]:
        raise ValueError("Dimensions of newloc and basis are not compatible!")          # :1203-1207
    if B.shape[1] != K:
        raise ValueError("Dimensions of basis and object$w are not compatible!")
    n2 = B.shape[0]
    yhat = np.empty((n2, TT), order="F")
    se = np.empty(n2) if se_report else None
    check(lib.frk_basis_predict_frk(ptr(B), n2, K, ptr(w), TT, None if V is None else ptr(V), ptr(yhat),
                                    None if se is None else ptr(se)))
    return {"pred.value": yhat + mu_new, "se": None if se is None else np.repeat(se[:, None], TT, axis=1)}
