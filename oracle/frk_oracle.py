"""numpy restatement of autoFRK's fixed-rank-kriging fit path (TEST INFRASTRUCTURE).

PARITY UNPINNED (see oracle/__init__.py).  Follows, line by line:
  src/autoFRK.cpp:24-32     getASCeigens       (Eigen SelfAdjointEigenSolver -> LAPACK dsyevd, ascending)
  R/helper.R:12-26          getEigen, getHalf
  R/autoFRK.R:333-397       cMLE   (incl. sol.v :335-351, sol.eta :352, neg2llik :353-362)
  R/autoFRK.R:399-480       cMLEimat
  R/autoFRK.R:845-852       invCz ;  R/helper.R:249-256 ZinvC (dead code upstream, restated for completeness)
  R/helper.R:28-58          getLikelihood
  R/autoFRK.R:199-331       selectBasis  (identity-D, no-NA 'fast' branch)
  R/autoFRK.R:739-843       indeMLE      (isimat / no-NA branch -> cMLEimat)
  R/autoFRK.R:121-197       autoFRK      (finescale = FALSE)
  R/autoFRK.R:1168-1423     predict.FRK  (LKobj = NULL branches)
`gram_stats` / `cmle_from_stats` are NOT in the reference: they restate cMLEimat as a function of
(G, C, zz) only -- the algebra the CUDA path uses -- and are themselves checked against the
literal restatement in tests/test_oracle.py.
"""
from __future__ import annotations

import math

import numpy as np
import scipy.linalg as sla

from .mrts_oracle import MrtsObject, mrts, predict_mrts, _as2d


def getASCeigens(A):
    """src/autoFRK.cpp:24-32: full symmetric eigendecomposition, ascending."""
    w, V = sla.eigh(np.asarray(A, dtype=np.float64), driver="evd")
    return {"value": w, "vector": V}


def getEigen(A):
    """R/helper.R:12-17: reverse to descending."""
    obj = getASCeigens(A)
    return {"value": obj["value"][::-1].copy(), "vector": obj["vector"][:, ::-1].copy()}


def getHalf(Fk, iDFk):
    """R/helper.R:19-26: (Fk' iDFk)^(-1/2), zero eigenvalues -> 0."""
    dec = getEigen(np.asarray(Fk).T @ np.asarray(iDFk))       # :20  (full product, not syrk)
    sroot = np.sqrt(np.maximum(dec["value"], 0.0))             # :22
    with np.errstate(divide="ignore"):
        sroot = np.where(sroot == 0, np.inf, sroot)            # :23
        sroot = 1.0 / sroot                                    # :24
    return dec["vector"] @ (sroot[:, None] * dec["vector"].T)  # :25


def sol_v(d, s, trS, n):
    """R/autoFRK.R:335-351 (identical copy at :400-416)."""
    d = np.asarray(d, dtype=np.float64)
    k = d.size
    if d.max() < max(trS / n, s):
        return max(trS / n - s, 0.0)
    cumd = np.cumsum(d)
    ks = np.arange(1, k + 1, dtype=np.float64)
    if k == n:
        ks[n - 1] = n - 1
    pick = d > ((trS - cumd) / (n - ks))
    L = int(np.max(np.nonzero(pick)[0])) + 1                   # 1-based max(which(pick))
    if L == n:
        L = n - 1
    return max((trS - cumd[L - 1]) / (n - L) - s, 0.0)


def sol_eta(d, s, v):
    """R/autoFRK.R:352."""
    return np.maximum(np.asarray(d) - s - v, 0.0)


def neg2llik(d, s, v, trS, n):
    """R/autoFRK.R:353-362."""
    d = np.asarray(d, dtype=np.float64)
    k = d.size
    eta = np.maximum(d - s - v, 0.0)
    with np.errstate(divide="ignore", invalid="ignore"):
        if np.max(eta / (s + v)) > 1e20:
            return math.inf
    return (n * math.log(2 * math.pi) + np.sum(np.log(eta + s + v)) + math.log(s + v) * (n - k)
            + 1.0 / (s + v) * trS - 1.0 / (s + v) * np.sum(d * eta / (eta + s + v)))


def _eigen_desc(A, k):
    """R `eigen(A)` on a symmetric matrix: descending values, first k."""
    w, V = sla.eigh(A)
    return w[::-1][:k].copy(), V[:, ::-1][:, :k].copy()


def cMLE(Fk, TT, trS, half, JSJ, s=0.0, ldet=None, wSave=False, onlylogLike=None, vfixed=None):
    """R/autoFRK.R:333-397."""
    if onlylogLike is None:
        onlylogLike = not wSave
    if ldet is None:
        ldet = 0.0
    Fk = np.asarray(Fk)
    n, k = Fk.shape
    d, P = _eigen_desc(JSJ, k)                                 # :368-370
    v = sol_v(d, s, trS, n) if vfixed is None else vfixed      # :371-375
    dii = np.maximum(d, 0.0)
    dhat = sol_eta(dii, s, v)
    if onlylogLike:
        return {"negloglik": neg2llik(dii, s, v, trS, n) * TT + ldet * TT}  # :378-381
    M = half @ P @ (dhat[:, None] * P.T) @ half                # :382
    L = None
    if wSave:
        L = Fk @ ((np.sqrt(dhat)[:, None] * P.T) @ half).T     # :387
        if np.all(dhat == 0):
            dhat = dhat.copy()
            dhat[0] = 0.1 ** 10                                # :388-390
        L = L[:, dhat > 0]                                     # :391
    return {"v": v, "M": M, "s": s, "negloglik": neg2llik(dii, s, v, trS, n) * TT + ldet * TT, "L": L}


def invCz(R, L, z):
    """R/autoFRK.R:845-852: (R + L L')^{-1} z by Woodbury.  R is a dense/diagonal n x n matrix or a
    1-D vector holding its diagonal (the only forms the in-scope callers produce)."""
    L = np.asarray(L, dtype=np.float64)
    z = _as2d(z)
    K = L.shape[1]
    R = np.asarray(R, dtype=np.float64)
    if R.ndim == 1:
        iRZ = z / R[:, None]
        iRL = L / R[:, None]
        inner = np.linalg.solve(np.eye(K) + L.T @ iRL, L.T @ iRZ)
        return iRZ - (L @ inner) / R[:, None]
    iR = np.linalg.inv(R)                                      # :847
    iRZ = iR @ z                                               # :848
    right = L @ np.linalg.solve(np.eye(K) + L.T @ iR @ L, L.T @ iRZ)  # :849
    return iRZ - iR @ right                                    # :851


def ZinvC(R, L, z):
    """R/helper.R:249-256 (never called upstream)."""
    L = np.asarray(L, dtype=np.float64)
    z = np.atleast_2d(np.asarray(z, dtype=np.float64))
    K = L.shape[1]
    iR = np.linalg.inv(np.asarray(R, dtype=np.float64))
    ZiR = z @ iR
    left = ZiR @ L @ np.linalg.inv(np.eye(K) + L.T @ iR @ L) @ L.T
    return ZiR - left @ iR


def cMLEimat(Fk, Data, s, wSave=False, S=None, onlylogLike=None):
    """R/autoFRK.R:399-480 -- the literal sequence of n-sized products."""
    if onlylogLike is None:
        onlylogLike = not wSave
    Fk = np.asarray(Fk, dtype=np.float64)
    Data = _as2d(Data)
    n, k = Fk.shape
    TT = Data.shape[1]
    trS = np.sum(np.sum(Data ** 2, axis=1)) / TT              # :431
    half = getHalf(Fk, Fk)                                     # :432
    ihF = half @ Fk.T                                          # :433
    if S is None:
        t = ihF @ Data
        JSJ = (t @ t.T) / TT                                   # :435
    else:
        JSJ = (ihF @ S) @ ihF.T                                # :438
    JSJ = (JSJ + JSJ.T) / 2                                    # :440
    d, P = _eigen_desc(JSJ, k)                                 # :441-443
    v = sol_v(d, s, trS, n)                                    # :444
    dii = np.maximum(d, 0.0)
    if onlylogLike:
        return {"negloglik": neg2llik(dii, s, v, trS, n) * TT}
    dhat = sol_eta(dii, s, v)                                  # :450
    M = half @ P @ (dhat[:, None] * P.T) @ half                # :451
    nll = neg2llik(dii, s, v, trS, n) * TT
    if not wSave:
        return {"v": v, "M": M, "s": s, "negloglik": nll}
    L = Fk @ ((np.sqrt(dhat)[:, None] * P.T) @ half).T         # :459
    if np.all(dhat == 0):
        dhat = dhat.copy()
        dhat[0] = 0.1 ** 10                                    # :460-462
    L = L[:, dhat > 0]                                         # :463
    invD = np.ones(n) / (s + v)                                # :464
    iDZ = invD[:, None] * Data                                 # :465
    KL = L.shape[1]
    right = L @ (np.linalg.solve(np.eye(KL) + L.T @ (invD[:, None] * L), L.T @ iDZ))  # :466-467
    INVtZ = iDZ - invD[:, None] * right                        # :468
    etatt = M @ Fk.T @ INVtZ                                   # :469
    GM = Fk @ M                                                # :470
    V = M - GM.T @ invCz(np.full(n, s + v), L, GM)             # :471-474
    return {"v": v, "M": M, "s": s, "negloglik": nll, "w": etatt, "V": V}


def cMLEsp(Fk, Data, Ddiag, wSave=False):
    """R/autoFRK.R:528-558 for a DIAGONAL Depsilon given by its diagonal `Ddiag` (the literal n-sized products)."""
    Fk = np.asarray(Fk, dtype=np.float64)
    Data = _as2d(Data)
    Dd = np.asarray(Ddiag, dtype=np.float64).ravel()
    iD = 1.0 / Dd                                              # :530
    ldetD = float(np.sum(np.log(Dd)))                          # :531
    iDFk = iD[:, None] * Fk                                    # :532
    half = getHalf(Fk, iDFk)                                   # :533
    ihFiD = half @ iDFk.T                                      # :534
    TT = Data.shape[1]
    t = ihFiD @ Data
    JSJ = (t @ t.T) / TT                                       # :536
    JSJ = (JSJ + JSJ.T) / 2                                    # :537
    trS = np.sum(np.sum((iD[:, None] * Data) * Data, axis=1)) / TT   # :538
    out = cMLE(Fk, TT, trS, half, JSJ, s=0.0, ldet=ldetD, wSave=wSave)  # :539
    if wSave:
        L = out["L"]                                           # :541
        invD = iD / (out["s"] + out["v"])                      # :542
        iDZ = invD[:, None] * Data                             # :543
        KL = L.shape[1]
        right0 = L @ np.linalg.inv(np.eye(KL) + L.T @ (invD[:, None] * L))      # :544-545
        INVtZ = iDZ - invD[:, None] * (right0 @ (L.T @ iDZ))   # :546
        out["w"] = out["M"] @ Fk.T @ INVtZ                     # :547-548
        GM = Fk @ out["M"]                                     # :549
        iDGM = invD[:, None] * GM                              # :550
        out["V"] = out["M"] - GM.T @ (iDGM - invD[:, None] * (right0 @ (L.T @ iDGM)))  # :551-552
    out["s"] = out.get("v")                                    # :554
    out.pop("v", None)                                         # :555
    out.pop("L", None)                                         # :556
    return out


# --------------------------------------------------------------------------------------------
# (G, C, zz) restatement -- what the CUDA path evaluates
# --------------------------------------------------------------------------------------------
def gram_stats(Fk, Data):
    """G = F'F (R/helper.R:20), C = F'Z, zz = sum(Z^2) (R/autoFRK.R:431)."""
    Fk = np.asarray(Fk, dtype=np.float64)
    Data = _as2d(Data)
    return Fk.T @ Fk, Fk.T @ Data, float(np.sum(Data ** 2))


def half_from_gram(G):
    dec = getEigen(G)
    sroot = np.sqrt(np.maximum(dec["value"], 0.0))
    with np.errstate(divide="ignore"):
        sroot = 1.0 / np.where(sroot == 0, np.inf, sroot)
    return dec["vector"] @ (sroot[:, None] * dec["vector"].T)


def cmle_from_stats(G, C, zz, n, TT, s=0.0, wSave=True):
    """cMLEimat as a function of the sufficient statistics only.

    ihF*Z = half*(F'Z) = half*C, so JSJ = (half C)(half C)'/T (R/autoFRK.R:433-435).  With
    L'L = D^{1/2} P' half G half P D^{1/2} = diag(dhat) (half G half = I on the range of G):
      w = half P diag(dhat/(s+v+dhat)) P' half C        (closed form of :464-469)
      V = half P diag(dhat (s+v)/(s+v+dhat)) P' half    (closed form of :470-474)
    """
    k = G.shape[0]
    trS = zz / TT
    half = half_from_gram(G)
    t = half @ C
    JSJ = (t @ t.T) / TT
    JSJ = (JSJ + JSJ.T) / 2
    d, P = _eigen_desc(JSJ, k)
    v = sol_v(d, s, trS, n)
    dii = np.maximum(d, 0.0)
    dhat = sol_eta(dii, s, v)
    hP = half @ P
    M = hP @ (dhat[:, None] * hP.T)
    out = {"v": v, "M": M, "s": s, "negloglik": neg2llik(dii, s, v, trS, n) * TT}
    if wSave:
        sv = s + v
        out["w"] = hP @ ((dhat / (sv + dhat))[:, None] * (hP.T @ C))
        out["V"] = hP @ ((dhat * sv / (sv + dhat))[:, None] * hP.T)
    return out


def getLikelihood(Data, Fk, M, s, Depsilon_diag):
    """R/helper.R:28-58 with a diagonal Depsilon (given as its diagonal)."""
    Data = _as2d(Data)
    Fk = np.asarray(Fk, dtype=np.float64)
    O = ~np.isnan(Data)
    TT = Data.shape[1]
    n2loglik = O.sum() * math.log(2 * math.pi)                 # :38
    Rdiag = s * np.asarray(Depsilon_diag, dtype=np.float64)    # :39
    w, U = sla.eigh(M)                                         # :40
    L = Fk @ U @ np.diag(np.sqrt(np.maximum(w, 0.0))) @ U.T    # :41-42
    K = Fk.shape[1]
    for tt in range(TT):
        o = O[:, tt]
        zt = Data[o, tt]
        Rt = Rdiag[o]
        Lt = L[o, :]
        if Lt.shape[0] == 1:                                   # :48-51 (NCOL(Lt)==1 after drop)
            n2loglik += math.log(Rt[0] + float((Lt @ Lt.T)[0, 0]))
        else:
            sign, ld = np.linalg.slogdet(np.eye(K) + Lt.T @ (Lt / Rt[:, None]))
            n2loglik += ld + np.sum(np.log(Rt)) + float(np.sum(zt * invCz(Rt, Lt, zt).ravel()))  # :53-54
    return n2loglik


def mkpd(M):
    """R/helper.R:278-288."""
    M = np.asarray(M, dtype=np.float64)
    v = np.min(sla.eigvalsh((M + M.T) / 2 if not np.allclose(M, M.T, rtol=0, atol=0) else M))
    if v <= 0:
        M = M + np.eye(M.shape[0]) * (max(0.0, -v) + 0.1 ** 7.5)
    return M


def ginv(X, tol=math.sqrt(np.finfo(float).eps)):
    """MASS::ginv: SVD pseudo-inverse, singular values <= tol * max dropped."""
    U, d, Vt = np.linalg.svd(np.asarray(X, dtype=np.float64))
    pos = d > max(tol * d[0], 0.0)
    return (Vt[pos].T * (1.0 / d[pos])) @ U[:, pos].T


def EM0miss(Fk, Data, Ddiag, maxit=50, avgtol=1e-6, wSave=False, vfixed=None):
    """R/autoFRK.R:560-737, DfromLK = NULL, external = FALSE, diagonal Depsilon (its diagonal `Ddiag`)."""
    Fk = np.asarray(Fk, dtype=np.float64)
    Data = _as2d(Data)
    O = ~np.isnan(Data)                                        # :564
    TT, K = Data.shape[1], Fk.shape[1]
    iD = 1.0 / np.asarray(Ddiag, dtype=np.float64).ravel()     # :575
    ziDz = np.zeros(TT)
    ziDB = np.zeros((TT, K))
    db = []
    for tt in range(TT):                                       # :587-636
        o = O[:, tt]
        iDt = iD[o]                                            # :607
        Bt = Fk[o, :]                                          # :609
        iDBt = iDt[:, None] * Bt                               # :611
        zt = Data[o, tt]                                       # :612
        ziDz[tt] = np.sum(zt * (iDt * zt))                     # :613
        ziDB[tt] = zt @ iDBt                                   # :614
        db.append((iDBt, zt, Bt.T @ iDBt))                     # :615, :628-634
    Z0 = np.where(np.isnan(Data), 0.0, Data)                   # :643-644
    old = cMLEimat(Fk, Z0, s=0.0, wSave=True)                  # :645
    old["s"] = old["v"] if vfixed is None else vfixed          # :646
    old["M"] = mkpd(old["M"])                                  # :647
    Ptt1 = old["M"]
    dif, cnt = math.inf, 0
    new, etatt = old, np.zeros((K, TT))
    while dif > avgtol * (100 * K ** 2) and cnt < maxit:       # :652
        etatt = np.zeros((K, TT))
        sumPtt = np.zeros((K, K))
        s1 = np.zeros(TT)
        for tt in range(TT):                                   # :657-669
            iDBt, zt, BiDBt = db[tt]
            iP = mkpd(ginv(mkpd(Ptt1)) + BiDBt / old["s"])     # :659
            Ptt = np.linalg.inv(iP)                            # :660
            Gt = Ptt @ iDBt.T / old["s"]                       # :661
            eta = Gt @ zt                                      # :662
            s1kk = np.diag(BiDBt @ (np.outer(eta, eta) + Ptt))  # :663
            sumPtt += Ptt                                      # :666
            etatt[:, tt] = eta                                 # :667
            s1[tt] = np.sum(s1kk)                              # :668
        M = (etatt @ etatt.T + sumPtt) / TT                    # :672
        if vfixed is None:
            sn = max((np.sum(ziDz) - 2 * np.sum(ziDB * etatt.T) + np.sum(s1)) / np.sum(O), 0.1 ** 8)  # :673
        else:
            sn = vfixed                                        # :678
        M = (M + M.T) / 2                                      # :681
        dif = np.sum(np.abs(M - old["M"])) + abs(sn - old["s"])  # :682
        cnt += 1
        old = new = {"M": M, "s": sn}
        Ptt1 = M                                               # :685
    n2loglik = getLikelihood(Data, Fk, new["M"], new["s"], Ddiag)   # :690
    out = {"M": new["M"], "s": new["s"], "negloglik": n2loglik, "niter": cnt}
    if wSave:
        out["w"] = etatt                                       # :729
        out["V"] = new["M"] - etatt @ etatt.T / TT
    return out


# --------------------------------------------------------------------------------------------
# model-fit orchestration (identity D, no missing values)
# --------------------------------------------------------------------------------------------
def _r_round(x):
    """R's round(): IEC 60559 half-to-even, same as numpy."""
    return np.round(x)


def kgrid(d, maxK):
    """Default K sequence, R/autoFRK.R:254-257."""
    K = np.unique(_r_round(np.arange(d + 1, maxK + 1e-9, maxK ** (1.0 / 3) * d))).astype(int)
    if K.size > 30:
        K = np.unique(_r_round(np.linspace(d + 1, maxK, 30))).astype(int)
    return K


def knn_impute(Data, loc, n_neighbor=3):
    """R/autoFRK.R:135-148 / :273-286: FNN::get.knnx by brute force (ties -> lower index)."""
    Data = np.array(_as2d(Data), copy=True)
    loc = _as2d(loc)
    for tt in range(Data.shape[1]):
        where = np.isnan(Data[:, tt])
        if not where.any():
            continue
        cidx = np.nonzero(~where)[0]
        q, r = loc[where], loc[cidx]
        d2 = ((q[:, None, :] - r[None, :, :]) ** 2).sum(axis=2)
        nn = np.argsort(d2, axis=1, kind="stable")[:, :n_neighbor]
        Data[where, tt] = Data[cidx[nn], tt].mean(axis=1)
    return Data


def selectBasis(Data, loc, maxK=None, Kseq=None, maxknot=5000, Fk=None, n_neighbor=3):
    """R/autoFRK.R:199-331, D = I, method='fast'.

    Replicates without data and rows never observed are dropped first (:203-220), so the knots are the observed
    locations; remaining NAs are imputed by kNN means (:273-286).  NB upstream indexes the UNPRUNED `loc` with
    pruned row numbers at :281 (misaligned when rows were dropped); the restatement uses the aligned locations."""
    Data = _as2d(Data)
    loc = _as2d(loc)
    Data = Data[:, ~np.all(np.isnan(Data), axis=0)]            # :203-206
    pick = np.nonzero(~np.all(np.isnan(Data), axis=1))[0]      # :213-220
    Data = Data[pick]
    d = loc.shape[1]
    N = pick.size                                              # :221
    klim = int(min(N, _r_round(10 * math.sqrt(N))))            # :222
    if N < maxknot:
        knot = loc[pick]                                       # :223-224
    else:
        raise NotImplementedError("subKnot (R RNG) is out of scope")
    if maxK is not None:
        maxK = int(_r_round(maxK))
    elif Kseq is not None:
        maxK = int(_r_round(max(Kseq)))
    else:
        maxK = klim                                            # :228-236
    if Kseq is not None:
        K = np.array(sorted(set(int(_r_round(k)) for k in Kseq)))  # unique(round(Kseq))
        if K.max() > maxK:
            raise ValueError("maximum of Kseq is larger than maxK!")
        K = K[K > d]
        if K.size == 0:
            raise ValueError("Not valid Kseq!")
    else:
        K = kgrid(d, maxK)                                     # :254-257
    if Fk is None:
        Fk = mrts(knot, int(K.max()), loc, maxknot)            # :260
    if np.isnan(Data).any():
        Data = knn_impute(Data, loc[pick], n_neighbor)         # :273-286
    F = Fk.X[pick]
    TT = Data.shape[1]
    iDFk = F                                                   # iD = I (:289-291)
    trS = np.sum(np.sum(Data * Data, axis=1)) / TT             # :304
    AIClist = np.full(K.size, np.inf)
    for i, k in enumerate(K):
        half = getHalf(F[:, :k], iDFk[:, :k])                  # :306
        ihFiD = half @ iDFk[:, :k].T                           # :307
        t = ihFiD @ Data
        JSJ = (t @ t.T) / TT                                   # :308
        JSJ = (JSJ + JSJ.T) / 2                                # :309
        AIClist[i] = cMLE(F[:, :k], TT, trS, half, JSJ)["negloglik"]  # :310-316
    df = np.where(K <= TT, K * (K + 1) / 2 + 1, K * TT + 1 - TT * (TT - 1) / 2)  # :321
    AIClist = AIClist + 2 * df                                 # :322
    Kopt = int(K[np.argmin(AIClist)])                          # :323
    out = Fk.leading(Kopt)                                     # :324-328
    out.extra["AIC"] = AIClist
    out.extra["Kgrid"] = K
    return out


def indeMLE(Data, Fk, wSave=False, sigma=0.0, Ddiag=None, maxit=50, avgtol=1e-6):
    """R/autoFRK.R:739-843: isimat & no-NA branch (:778-803); with NAs (diagonal D) the EM0miss branch (:829-841)
    after dropping empty replicates (:746-750) and never-observed rows (:754-771)."""
    Data = _as2d(Data)
    Fk = np.asarray(Fk, dtype=np.float64)
    if np.isnan(Data).any():
        TT, K = Data.shape[1], Fk.shape[1]
        notempty = np.nonzero((~np.isnan(Data)).sum(axis=0) > 0)[0]            # :746-747
        Data = Data[:, notempty]                                               # :749
        keep = (~np.isnan(Data)).sum(axis=1) > 0                               # del <- which(rowSums(...) == 0)  :754
        pick = np.nonzero(keep)[0]                                             # :755, :762
        D0 = np.ones(Fk.shape[0]) if Ddiag is None else np.asarray(Ddiag, dtype=np.float64).ravel()   # :759
        out = EM0miss(Fk[keep], Data[keep], D0[keep], maxit, avgtol, wSave)     # :830
        if wSave:
            w = np.zeros((K, TT))                                              # :834-836
            w[:, notempty] = out["w"]
            out["w"] = w
            out["pinfo"] = {"D": D0, "pick": pick}                             # :838
            out["missing"] = {"miss": np.isnan(Data[keep]).astype(float), "maxit": maxit, "avgtol": avgtol}  # :731-734
        return out
    out = cMLEimat(Fk, Data, s=sigma, wSave=wSave)             # :787
    if "v" in out:
        out["s"] = out["v"] if sigma == 0 else sigma           # :788-795
        del out["v"]
    return out


def autoFRK(Data, loc, mu=0.0, G=None, maxK=None, Kseq=None, maxknot=5000, method="fast"):
    """R/autoFRK.R:121-197, finescale=FALSE, D = I; method = "EM" needs G (selectBasis with NAs is not restated)."""
    Data = _as2d(Data) - mu                                    # :125
    Fk = G if G is not None else selectBasis(Data, loc, maxK=maxK, Kseq=Kseq, maxknot=maxknot)  # :126-133
    if method == "fast" and np.isnan(Data).any():
        Data = knn_impute(Data, loc)                           # :135-148 (method = "fast")
    F = Fk.X if isinstance(Fk, MrtsObject) else np.asarray(Fk)
    obj = indeMLE(Data, F, wSave=True)                         # :151
    obj["G"] = Fk                                              # :185
    obj["LKobj"] = None
    return obj


def predict_FRK(obj, obsData=None, obsloc=None, mu_obs=0.0, newloc=None, basis=None, mu_new=0.0,
                se_report=False):
    """R/autoFRK.R:1168-1423, LKobj = NULL (with or without the "missing" attribute of an EM0miss fit)."""
    if "w" not in obj:
        raise ValueError('input model (object) should use the option "wsave=TRUE"!')
    if newloc is None:
        newloc = obsloc                                        # default newloc = obsloc
    G = obj["G"]
    Gmat = G.X if isinstance(G, MrtsObject) else np.asarray(G)
    if basis is None:                                          # :1174-1188
        if newloc is None and obsloc is None:
            basis = Gmat
        else:
            if not isinstance(G, MrtsObject):
                raise ValueError("Basis matrix of new locations should be given (unless the model was fitted with mrts bases)!")
            basis = Gmat if newloc is None else predict_mrts(G, newloc)
    basis = np.atleast_2d(np.asarray(basis, dtype=np.float64))
    nobs = Gmat.shape[0] if obsloc is None else _as2d(obsloc).shape[0]
    if obsData is not None:
        obsData = np.asarray(obsData, dtype=np.float64).reshape(-1, 1) - mu_obs  # :1198
        if obsData.size != nobs:
            raise ValueError("Dimensions of obsloc and obsData are not compatible!")
    if newloc is not None:
        if basis.shape[0] != _as2d(newloc).shape[0]:
            raise ValueError("Dimensions of newloc and basis are not compatible!")
    elif basis.shape[0] != Gmat.shape[0]:
        raise ValueError("Dimensions of obsloc and basis are not compatible!")
    se = None
    if obsloc is None and obsData is None:                     # :1214-1247
        yhat = basis @ obj["w"]                                # :1216
        if se_report:
            TT = obj["w"].shape[1]
            if obj.get("missing") is None:
                se = np.sqrt(np.maximum(0.0, np.sum((basis @ obj["V"]) * basis, axis=1)))  # :1220
                se = np.repeat(se[:, None], TT, axis=1)        # :1222
            else:                                              # :1224-1245, literal
                se = np.full((basis.shape[0], TT), np.nan)
                pick = obj["pinfo"]["pick"]                    # :1226
                D0 = np.asarray(obj["pinfo"]["D"], dtype=np.float64)[pick]   # :1227 (diagonal)
                miss = np.asarray(obj["missing"]["miss"]) == 1  # :1228
                Fk = Gmat[pick, :]                             # :1229
                M = obj["M"]
                lam, U = sla.eigh(M)
                lam, U = lam[::-1], U[:, ::-1]                 # dec <- eigen(M)  :1231
                for tt in range(TT):                           # :1232 (miss[, tt] past its last column: R error)
                    o = ~miss[:, tt]
                    Gt = Fk[o, :]                              # :1233
                    GM = Gt @ M                                # :1234
                    L = Gt @ U @ np.diag(np.sqrt(np.maximum(lam, 0.0)))   # :1236-1239
                    V = M - GM.T @ invCz(obj["s"] * D0[o], L, GM)          # :1240
                    se[:, tt] = np.sqrt(np.maximum(0.0, np.sum((basis @ V) * basis, axis=1)))   # :1241-1242
    if obsData is not None:                                    # :1248-1274
        pick = np.nonzero(~np.isnan(obsData.ravel()))[0]
        if obsloc is None:
            Gp = Gmat[pick, :]
        else:
            Gp = predict_mrts(G, _as2d(obsloc)[pick, :])       # :1256
        M = obj["M"]
        GM = Gp @ M
        w, U = sla.eigh(M)
        w, U = w[::-1], U[:, ::-1]                             # R eigen(): descending
        L = Gp @ U @ np.diag(np.sqrt(np.maximum(w, 0.0)))      # :1261-1264
        Rd = np.full(pick.size, obj["s"])                      # s * De, De = I
        yhat = basis @ GM.T @ invCz(Rd, L, obsData[pick])      # :1265-1268
        if se_report:
            V = M - GM.T @ invCz(Rd, L, GM)                    # :1270-1271
            se = np.sqrt(np.maximum(0.0, np.sum((basis @ V) * basis, axis=1)))[:, None]  # :1272
    return {"pred.value": yhat + mu_new, "se": se}
