# R side of the frk_b200 overlay for autoFRK 1.4.4: source this file AFTER the package's own R/*.R (or append it to
# R/autoFRK.R) in a package built with r/src/frk_b200_shim.c + r/src/Makevars.  Each function below replaces the
# reference function of the same name; everything not listed here (mrts(), autoFRK(), indeMLE(), cMLE(), ...) is
# the package's own code and keeps working because the four registered routines keep their names and contracts
# (R/RcppExports.R:4-18 stays byte for byte).
#
# NOT RUN in the build container (no R interpreter there).  The C routines these wrappers call are executed in the
# test-suite against a mock of R's C API (tests/test_r_shim_gpu.py); the R code itself is a reviewed transliteration.

## predict.mrts (R/autoFRK.R:1439-1470) -----------------------------------------------------------------------
## one fused kernel pass: no n2 x m kernel matrix, no n2 x K intermediate copies (the reference makes three)
predict.mrts <- function(object, newx, ...) {
  if (missing(newx)) {
    return(object)
  }
  Xu <- attr(object, "Xu")
  ndims <- NCOL(Xu)
  k <- NCOL(object)
  x0 <- matrix(as.double(as.matrix(newx)), ncol = ndims)
  kstar <- (k - ndims - 1)
  if (kstar <= 0) {                          # polynomial-only object: the reference's own R arithmetic (:1462-1464)
    shift <- colMeans(Xu)
    X2 <- sweep(cbind(x0), 2, shift, "-")
    return(as.matrix(cbind(1, sweep(X2, 2, attr(object, "nconst"), "/"))))
  }
  .Call(`_autoFRK_predict_mrts_basis`, Xu, attr(object, "BBBH"), attr(object, "UZ"),
        as.double(attr(object, "nconst")), as.integer(k), x0)
}

## getHalf (R/helper.R:19-26) -----------------------------------------------------------------------------------
## t(Fk) %*% iDFk is the n K^2 product of the fit; iDFk = Fk / D for the diagonal-D callers of this function
getHalf <- function(Fk, iDFk) {
  G <- if (identical(Fk, iDFk)) {
    .Call(`_autoFRK_gram_stats`, as.matrix(Fk), matrix(0, NROW(Fk), 1))$G
  } else {
    t(Fk) %*% iDFk                           # general (non-diagonal) D: the reference's product
  }
  .Call(`_autoFRK_getHalf_stats`, G)
}

## cMLEimat (R/autoFRK.R:399-480) -------------------------------------------------------------------------------
## one pass over the rows for (G, C, zz); v, M, negloglik, w, V follow from those K x K / K x T statistics
cMLEimat <- function(Fk, Data, s, wSave = FALSE, S = NULL, onlylogLike = !wSave) {
  Data <- as.matrix(Data)
  st <- .Call(`_autoFRK_gram_stats`, as.matrix(Fk), Data)
  out <- .Call(`_autoFRK_cMLEimat_stats`, st$G, st$C, st$zz, as.double(NROW(Fk)), as.double(s), as.logical(wSave))
  if (onlylogLike) {
    return(list(negloglik = out$negloglik))
  }
  out$v <- as.numeric(out$v)
  out$negloglik <- as.numeric(out$negloglik)
  out
}

## the AIC loop of selectBasis (R/autoFRK.R:305-317), identity D, no missing values ---------------------------
## every candidate K works on leading blocks of ONE (G, C): one Cholesky factor instead of 2 x length(K) eigen-solves
selectBasis_AIClist <- function(Fk, Data, K) {
  Data <- as.matrix(Data)
  st <- .Call(`_autoFRK_gram_stats`, as.matrix(Fk[, 1:max(K)]), Data)
  nll <- .Call(`_autoFRK_cMLEimat_sweep`, st$G, st$C, st$zz, as.integer(NCOL(Data)), as.double(NROW(Fk)), 0, as.integer(K))
  for (k in which(is.nan(nll))) {            # a leading block that is not safely positive definite: the eigen path
    nll[k] <- cMLEimat(Fk[, 1:K[k]], Data, s = 0, wSave = FALSE)$negloglik
  }
  nll
}

## predict.FRK at new locations of an mrts model, no LatticeKrig part, no new observations
## (R/autoFRK.R:1184,1216,1220): yhat and se from ONE kernel pass with T + rank(V) columns; the n2 x K basis and the
## n2 x K x K product basis %*% V never exist.  Everything else falls through to the package's predict.FRK.
predict.FRK <- local({
  reference_predict.FRK <- get0("predict.FRK", mode = "function")
  function(object, obsData = NULL, obsloc = NULL, mu.obs = 0, newloc = obsloc, basis = NULL, mu.new = 0,
           se.report = FALSE, ...) {
    fused <- is.null(object$LKobj) && is.null(obsData) && is.null(obsloc) && is.null(basis) && !is.null(newloc) &&
      inherits(object$G, "mrts") && !is.null(object$w) &&
      NCOL(object$G) > NCOL(attr(object$G, "Xu")) + 1
    if (!fused) {
      return(reference_predict.FRK(object, obsData = obsData, obsloc = obsloc, mu.obs = mu.obs, newloc = newloc,
                                   basis = basis, mu.new = mu.new, se.report = se.report, ...))
    }
    G <- object$G
    x0 <- matrix(as.double(as.matrix(newloc)), ncol = NCOL(attr(G, "Xu")))
    miss <- attr(object, "missing")
    if (!is.null(miss) && se.report) {
      ## a model fitted through EM0miss: one V per replicate from the rows observed in it (R/autoFRK.R:1224-1245).
      ## All masked Grams t(G_t) solve(De_t) G_t come from one pass over the basis rows, each V_t is K x K algebra,
      ## and yhat[, tt], se[, tt] come out of one fused kernel pass per replicate.
      pinfo <- attr(object, "pinfo")
      pick <- pinfo$pick
      D0 <- as.double(diag(pinfo$D))[pick]
      observed <- 1 - as.matrix(miss$miss)
      TT <- NCOL(object$w)
      if (NCOL(observed) < TT) stop("subscript out of bounds")   # miss[, tt] at :1233 when the fit dropped empty replicates
      Fk <- matrix(as.double(unclass(G)[pick, ]), nrow = length(pick))
      K <- NCOL(Fk)
      GG <- .Call(`_autoFRK_masked_grams`, Fk, matrix(as.double(observed), nrow = NROW(observed)), D0)
      yhat <- se <- matrix(0, NROW(x0), TT)
      for (tt in 1:TT) {
        Vt <- .Call(`_autoFRK_krige_V`, matrix(GG[, tt], K, K) / object$s, as.matrix(object$M))
        res <- .Call(`_autoFRK_predict_frk`, attr(G, "Xu"), attr(G, "BBBH"), attr(G, "UZ"), as.double(attr(G, "nconst")),
                     as.integer(NCOL(G)), x0, as.matrix(object$w[, tt, drop = FALSE]), Vt)
        yhat[, tt] <- res$yhat
        se[, tt] <- res$se
      }
      return(list(pred.value = yhat + mu.new, se = se))
    }
    res <- .Call(`_autoFRK_predict_frk`, attr(G, "Xu"), attr(G, "BBBH"), attr(G, "UZ"), as.double(attr(G, "nconst")),
                 as.integer(NCOL(G)), x0, as.matrix(object$w), if (se.report) as.matrix(object$V) else NULL)
    TT <- NCOL(object$w)
    list(pred.value = res$yhat + mu.new,
         se = if (se.report) matrix(res$se, length(res$se), TT) else NULL)
  }
})

## The mrts attribute round trip (R/autoFRK.R:1115-1120): mrts() attaches UZ, Xu, nconst, BBBH and class "mrts" to the
## basis matrix; `[` drops them (SURVEY.md 9.11) and selectBasis re-attaches them (:324-328).  The routines above read
## exactly these four attributes, as double vectors / matrices with their R dims -- nothing else travels.
