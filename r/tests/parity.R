## Parity of the overlay against stock autoFRK, for a maintainer with R and a B200 (never parsed by R here: R is not
## installed in the build container).  Two installations of the package cannot live in one session, so:
##   Rscript parity.R dump stock.rds      # in a library that holds egpivo/autoFRK 1.4.4
##   Rscript parity.R dump overlay.rds    # in a library that holds the package with r/ applied (FRK_B200_HOME set)
##   Rscript parity.R compare stock.rds overlay.rds
## The cases follow the shape of the reference's documented examples (a 30 x 30 grid, 150 sites, 20 replicates; knots
## in d = 1, 2, 3), generated here; tolerances as in tests/: 1e-10 column-max-relative, eigen-columns up to sign.
args <- commandArgs(trailingOnly = TRUE)

colrel <- function(a, b) {            # max over columns of max|a - b| / max|b|, columns of `a` sign-aligned to `b`
  a <- as.matrix(a); b <- as.matrix(b)
  stopifnot(all(dim(a) == dim(b)))
  worst <- 0
  for (j in seq_len(ncol(b))) {
    s <- if (sum(a[, j] * b[, j]) < 0) -1 else 1
    worst <- max(worst, max(abs(s * a[, j] - b[, j])) / max(abs(b[, j]), .Machine$double.xmin))
  }
  worst
}

dump <- function(path) {
  library(autoFRK)
  out <- list()
  set.seed(1)
  for (d in 1:3) {                                   # mrts() / predict.mrts(): src/autoFRK.cpp:60-153, 160-324
    knot <- matrix(runif(60 * d), 60, d)
    newx <- matrix(runif(500 * d), 500, d)
    B <- mrts(knot, 20)
    out[[paste0("mrts_d", d)]] <- list(X = unclass(B)[, ], UZ = attr(B, "UZ"), BBBH = attr(B, "BBBH"),
                                       nconst = attr(B, "nconst"), pred = unclass(predict(B, newx))[, ])
  }
  ## a smooth two-component field on the unit square, 20 replicates with random amplitudes, 150 noisy sites
  side <- seq(0, 1, length.out = 30)
  grids <- as.matrix(expand.grid(side, side))
  bump <- function(cx, cy, freq) cos(freq * pi * sqrt((grids[, 1] - cx)^2 + (grids[, 2] - cy)^2))
  field <- cbind(bump(0, 1, 1), bump(0.75, 0.25, 2))
  amp <- rbind(rnorm(20, sd = 5), rnorm(20, sd = 3))
  sites <- sample(nrow(grids), 150)
  zt <- (field %*% amp)[sites, ] + matrix(rnorm(150 * 20, sd = sqrt(5)), 150, 20)
  X <- grids[sites, ]
  one <- autoFRK(Data = zt[, 1], loc = X, maxK = 15)                       # R/autoFRK.R:121-197, T = 1
  out$one <- list(K = NCOL(one$G), M = one$M, s = one$s, w = one$w, V = one$V, negloglik = one$negloglik,
                  pred = predict(one, newloc = grids, se.report = TRUE))
  G <- mrts(X, 15)
  multi <- autoFRK(Data = zt, loc = X, G = G)                             # 20 replicates, given basis
  out$multi <- list(M = multi$M, s = multi$s, w = multi$w, negloglik = multi$negloglik,
                    pred = predict(multi, newloc = grids, se.report = TRUE))
  ztm <- zt; ztm[sample(length(ztm), 300)] <- NA                           # EM0miss and its per-replicate se
  em <- autoFRK(Data = ztm, loc = X, G = G, method = "EM")
  out$em <- list(M = em$M, s = em$s, w = em$w, pred = predict(em, newloc = grids, se.report = TRUE))
  saveRDS(out, path)
}

compare <- function(pa, pb) {
  a <- readRDS(pa); b <- readRDS(pb)
  report <- function(what, err, tol) cat(sprintf("%-28s %.3e  (tol %.0e) %s\n", what, err, tol, if (err <= tol) "ok" else "FAIL"))
  for (d in 1:3) {
    k <- paste0("mrts_d", d)
    report(paste(k, "X"), colrel(a[[k]]$X, b[[k]]$X), 1e-8)               # per-eigenvector quantities: 1e-8 (SURVEY 8c)
    report(paste(k, "predict"), colrel(a[[k]]$pred, b[[k]]$pred), 1e-8)
    report(paste(k, "nconst"), max(abs(a[[k]]$nconst - b[[k]]$nconst) / abs(b[[k]]$nconst)), 1e-10)
  }
  rel <- function(x, y) max(abs(x - y)) / max(abs(y))
  stopifnot(a$one$K == b$one$K)
  for (k in c("one", "multi", "em")) {
    report(paste(k, "M"), rel(a[[k]]$M, b[[k]]$M), 1e-9)
    report(paste(k, "s"), abs(a[[k]]$s - b[[k]]$s) / abs(b[[k]]$s), 1e-9)
    report(paste(k, "pred.value"), rel(a[[k]]$pred$pred.value, b[[k]]$pred$pred.value), 1e-8)
    report(paste(k, "se"), rel(a[[k]]$pred$se, b[[k]]$pred$se), 1e-5)     # V = M - ... cancels in the reference
  }
}

if (length(args) == 2 && args[1] == "dump") dump(args[2]) else
  if (length(args) == 3 && args[1] == "compare") compare(args[2], args[3]) else
    stop("usage: Rscript parity.R dump out.rds | compare a.rds b.rds")
