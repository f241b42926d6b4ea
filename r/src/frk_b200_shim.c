/* .Call shim binding autoFRK's four registered native routines (src/RcppExports.cpp:71-77), plus the new
 * routines the R overlay (r/R/overrides.R) calls, to libfrk_b200.so.  Written against R's C API only
 * (Rinternals.h).  R is not installed in the build container: there the file is compiled against declarations of
 * that API (tests/r_stub) and EXECUTED against a small mock of it (tests/r_stub/mock_r.c, tests/test_r_shim_gpu.py),
 * which checks argument marshalling, the returned list layouts and the error path; linking against a real R is the
 * one step left to a maintainer (r/README.md).
 *
 * Build inside the package: r/src/Makevars; drop src/autoFRK.cpp + src/RcppExports.cpp (this file replaces both;
 * LinkingTo: Rcpp, RcppEigen, RSpectra are no longer needed).
 *
 * Conventions honoured: arguments are REALSXP matrices read in place (column-major,
 * dims from the dim attribute; integer input is an error exactly as with Eigen::Map<MatrixXd>),
 * results are freshly allocated named lists, errors leave through Rf_error() after every native
 * resource has been released (the library never throws across the C ABI).
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <stdint.h>
#include "frk_b200.h"

static void need_real_matrix(SEXP x, const char *what) {
    if (TYPEOF(x) != REALSXP) Rf_error("Wrong R type for mapped matrix");   /* RcppEigen's message */
    if (!Rf_isMatrix(x) && XLENGTH(x) > 0 && what[0] != 'n') (void)what;
}
static R_xlen_t nrow_of(SEXP x) { return Rf_isMatrix(x) ? Rf_nrows(x) : XLENGTH(x); }
static int ncol_of(SEXP x) { return Rf_isMatrix(x) ? Rf_ncols(x) : 1; }
static void check(int32_t status) { if (status != FRK_OK) Rf_error("%s", frk_last_error()); }

static SEXP named_list(int n, const char **names) {
    SEXP out = PROTECT(Rf_allocVector(VECSXP, n));
    SEXP nm = PROTECT(Rf_allocVector(STRSXP, n));
    for (int i = 0; i < n; i++) SET_STRING_ELT(nm, i, Rf_mkChar(names[i]));
    Rf_setAttrib(out, R_NamesSymbol, nm);
    UNPROTECT(2);
    return out;
}

/* getASCeigens(A) -> list(value, vector)                       src/autoFRK.cpp:24-32 */
SEXP _autoFRK_getASCeigens(SEXP A) {
    need_real_matrix(A, "A");
    int n = Rf_nrows(A);
    const char *names[] = {"value", "vector"};
    SEXP out = PROTECT(named_list(2, names));
    SEXP value = PROTECT(Rf_allocVector(REALSXP, n));
    SEXP vector = PROTECT(Rf_allocMatrix(REALSXP, n, n));
    check(frk_getASCeigens(REAL(A), n, REAL(value), REAL(vector)));
    SET_VECTOR_ELT(out, 0, value);
    SET_VECTOR_ELT(out, 1, vector);
    UNPROTECT(3);
    return out;
}

/* mrtsrcpp(Xu, xobs_diag, k) -> list(X, UZ, BBBH, nconst)       src/autoFRK.cpp:224-242 */
SEXP _autoFRK_mrtsrcpp(SEXP Xu, SEXP xobs_diag, SEXP k_) {
    need_real_matrix(Xu, "Xu"); need_real_matrix(xobs_diag, "xobs_diag");
    int64_t n = nrow_of(Xu); int d = ncol_of(Xu); int k = Rf_asInteger(k_);
    int kk = k > 0 ? k : 0;
    const char *names[] = {"X", "UZ", "BBBH", "nconst"};
    SEXP out = PROTECT(named_list(4, names));
    SEXP X = PROTECT(Rf_allocMatrix(REALSXP, (int)n, kk + d + 1));
    SEXP UZ = PROTECT(Rf_allocMatrix(REALSXP, (int)n + d + 1, kk + d + 1));
    SEXP BBBH = PROTECT(Rf_allocMatrix(REALSXP, d + 1, (int)n));
    SEXP nconst = PROTECT(Rf_allocVector(REALSXP, d));
    check(frk_mrtsrcpp(REAL(Xu), n, d, REAL(xobs_diag), (int)nrow_of(xobs_diag), ncol_of(xobs_diag), k,
                       REAL(X), REAL(UZ), REAL(BBBH), REAL(nconst)));
    SET_VECTOR_ELT(out, 0, X); SET_VECTOR_ELT(out, 1, UZ); SET_VECTOR_ELT(out, 2, BBBH); SET_VECTOR_ELT(out, 3, nconst);
    UNPROTECT(5);
    return out;
}

/* mrtsrcpp_predict0(Xu, xobs_diag, xnew, k) -> list(X, UZ, BBBH, nconst, X1)   src/autoFRK.cpp:248-281 */
SEXP _autoFRK_mrtsrcpp_predict0(SEXP Xu, SEXP xobs_diag, SEXP xnew, SEXP k_) {
    need_real_matrix(Xu, "Xu"); need_real_matrix(xobs_diag, "xobs_diag"); need_real_matrix(xnew, "xnew");
    int64_t n = nrow_of(Xu), n2 = nrow_of(xnew); int d = ncol_of(Xu); int k = Rf_asInteger(k_);
    int kk = k > 0 ? k : 0;
    const char *names[] = {"X", "UZ", "BBBH", "nconst", "X1"};
    SEXP out = PROTECT(named_list(5, names));
    SEXP X = PROTECT(Rf_allocMatrix(REALSXP, (int)n, kk + d + 1));
    SEXP UZ = PROTECT(Rf_allocMatrix(REALSXP, (int)n + d + 1, kk + d + 1));
    SEXP BBBH = PROTECT(Rf_allocMatrix(REALSXP, d + 1, (int)n));
    SEXP nconst = PROTECT(Rf_allocVector(REALSXP, d));
    SEXP X1 = PROTECT(Rf_allocMatrix(REALSXP, (int)n2, kk));
    check(frk_mrtsrcpp_predict0(REAL(Xu), n, d, REAL(xobs_diag), (int)nrow_of(xobs_diag), ncol_of(xobs_diag),
                                REAL(xnew), n2, ncol_of(xnew), k, REAL(X), REAL(UZ), REAL(BBBH), REAL(nconst), REAL(X1)));
    SET_VECTOR_ELT(out, 0, X); SET_VECTOR_ELT(out, 1, UZ); SET_VECTOR_ELT(out, 2, BBBH);
    SET_VECTOR_ELT(out, 3, nconst); SET_VECTOR_ELT(out, 4, X1);
    UNPROTECT(6);
    return out;
}

/* mrtsrcpp_predict(Xu, xobs_diag, xnew, BBBH, UZ, nconst, k) -> list(X = Xu, UZ, BBBH, nconst, X1)
 * src/autoFRK.cpp:287-326; the first four members are the caller's own objects (:320-323). */
SEXP _autoFRK_mrtsrcpp_predict(SEXP Xu, SEXP xobs_diag, SEXP xnew, SEXP BBBH, SEXP UZ, SEXP nconst, SEXP k_) {
    need_real_matrix(Xu, "Xu"); need_real_matrix(xnew, "xnew"); need_real_matrix(BBBH, "BBBH");
    need_real_matrix(UZ, "UZ"); need_real_matrix(nconst, "nconst");
    int64_t n = nrow_of(Xu), n2 = nrow_of(xnew); int d = ncol_of(Xu); int k = Rf_asInteger(k_);
    const char *names[] = {"X", "UZ", "BBBH", "nconst", "X1"};
    SEXP out = PROTECT(named_list(5, names));
    SEXP X1 = PROTECT(Rf_allocMatrix(REALSXP, (int)n2, k > 0 ? k : 0));
    check(frk_mrtsrcpp_predict(REAL(Xu), n, d, TYPEOF(xobs_diag) == REALSXP ? REAL(xobs_diag) : NULL, REAL(xnew), n2,
                               ncol_of(xnew), REAL(BBBH), nrow_of(BBBH), ncol_of(BBBH), REAL(UZ), nrow_of(UZ),
                               ncol_of(UZ), REAL(nconst), XLENGTH(nconst), k, REAL(X1)));
    SET_VECTOR_ELT(out, 0, Xu); SET_VECTOR_ELT(out, 1, UZ); SET_VECTOR_ELT(out, 2, BBBH);
    SET_VECTOR_ELT(out, 3, nconst); SET_VECTOR_ELT(out, 4, X1);
    UNPROTECT(2);
    return out;
}

/* new: frk_gram_stats(Fk, Data) -> list(G, C, zz); frk_cMLEimat_stats(G, C, zz, n, TT, s, wSave) -> list(v, M, s, negloglik[, w, V]) */
SEXP _autoFRK_gram_stats(SEXP Fk, SEXP Data) {
    need_real_matrix(Fk, "Fk"); need_real_matrix(Data, "Data");
    int64_t n = nrow_of(Fk); int K = ncol_of(Fk), T = ncol_of(Data);
    const char *names[] = {"G", "C", "zz"};
    SEXP out = PROTECT(named_list(3, names));
    SEXP G = PROTECT(Rf_allocMatrix(REALSXP, K, K));
    SEXP C = PROTECT(Rf_allocMatrix(REALSXP, K, T));
    SEXP zz = PROTECT(Rf_allocVector(REALSXP, 1));
    check(frk_gram_stats(REAL(Fk), n, K, REAL(Data), T, NULL, REAL(G), REAL(C), REAL(zz)));
    SET_VECTOR_ELT(out, 0, G); SET_VECTOR_ELT(out, 1, C); SET_VECTOR_ELT(out, 2, zz);
    UNPROTECT(4);
    return out;
}

/* getHalf from the K x K cross-product t(Fk) %*% iDFk   (R/helper.R:19-26) */
SEXP _autoFRK_getHalf_stats(SEXP G) {
    need_real_matrix(G, "G");
    int K = Rf_nrows(G);
    SEXP half = PROTECT(Rf_allocMatrix(REALSXP, K, K));
    check(frk_getHalf(REAL(G), K, K, REAL(half)));
    UNPROTECT(1);
    return half;
}

SEXP _autoFRK_cMLEimat_stats(SEXP G, SEXP C, SEXP zz, SEXP n_, SEXP s_, SEXP wSave_) {
    int K = Rf_nrows(G), T = Rf_ncols(C), wsave = Rf_asLogical(wSave_);
    const char *names[] = {"v", "M", "s", "negloglik", "w", "V"};
    SEXP out = PROTECT(named_list(wsave ? 6 : 4, names));
    SEXP v = PROTECT(Rf_allocVector(REALSXP, 1)), nll = PROTECT(Rf_allocVector(REALSXP, 1));
    SEXP M = PROTECT(Rf_allocMatrix(REALSXP, K, K));
    SEXP w = PROTECT(Rf_allocMatrix(REALSXP, K, T)), V = PROTECT(Rf_allocMatrix(REALSXP, K, K));
    check(frk_cMLEimat_stats(REAL(G), K, REAL(C), K, Rf_asReal(zz), K, T, (int64_t)Rf_asReal(n_), Rf_asReal(s_), wsave, 0, 0,
                             REAL(v), REAL(nll), REAL(M), wsave ? REAL(w) : NULL, wsave ? REAL(V) : NULL));
    SET_VECTOR_ELT(out, 0, v); SET_VECTOR_ELT(out, 1, M); SET_VECTOR_ELT(out, 2, s_); SET_VECTOR_ELT(out, 3, nll);
    if (wsave) { SET_VECTOR_ELT(out, 4, w); SET_VECTOR_ELT(out, 5, V); }
    UNPROTECT(6);
    return out;
}

/* selectBasis' K loop (R/autoFRK.R:305-317): negloglik for every K in `Ks` (integer vector) from the max-K statistics */
SEXP _autoFRK_cMLEimat_sweep(SEXP G, SEXP C, SEXP zz, SEXP TT_, SEXP n_, SEXP s_, SEXP Ks) {
    need_real_matrix(G, "G");
    need_real_matrix(C, "C");
    if (TYPEOF(Ks) != INTSXP) Rf_error("Ks must be an integer vector");
    int nK = (int)XLENGTH(Ks);
    SEXP nll = PROTECT(Rf_allocVector(REALSXP, nK));
    check(frk_cMLEimat_sweep(REAL(G), Rf_nrows(G), REAL(C), Rf_nrows(C), Rf_asReal(zz), Rf_asInteger(TT_),
                             (int64_t)Rf_asReal(n_), Rf_asReal(s_), INTEGER(Ks), nK, REAL(nll)));
    UNPROTECT(1);
    return nll;   /* NaN entries: the caller uses _autoFRK_cMLEimat_stats on the leading blocks for that K */
}

/* ---- routines behind r/R/overrides.R ------------------------------------------------------------------------ */
static frk_plan *plan_of(SEXP Xu, SEXP BBBH, SEXP UZ, SEXP nconst, SEXP K_) {
    need_real_matrix(Xu, "Xu"); need_real_matrix(BBBH, "BBBH"); need_real_matrix(UZ, "UZ"); need_real_matrix(nconst, "nconst");
    frk_plan *plan = NULL;
    /* predict.mrts passes k = NCOL(object) (R/autoFRK.R:1458): the trailing d+1 zero columns are detected and skipped */
    check(frk_plan_create(&plan, REAL(Xu), nrow_of(Xu), ncol_of(Xu), REAL(BBBH), REAL(UZ), nrow_of(UZ), ncol_of(UZ),
                          REAL(nconst), Rf_asInteger(K_), 0));
    return plan;
}
/* a failing call must release the plan before Rf_error() longjmps away */
static void check_plan(frk_plan *plan, int32_t status) {
    if (status != FRK_OK) {
        frk_plan_destroy(plan);
        Rf_error("%s", frk_last_error());
    }
}

/* predict.mrts body (R/autoFRK.R:1453-1466): cbind(1, (x0 - shift) / nconst, X1[, 1:kstar]) in one kernel pass */
SEXP _autoFRK_predict_mrts_basis(SEXP Xu, SEXP BBBH, SEXP UZ, SEXP nconst, SEXP K_, SEXP xnew) {
    need_real_matrix(xnew, "xnew");
    if (ncol_of(xnew) != ncol_of(Xu)) Rf_error("xnew must have %d columns", ncol_of(Xu));   /* src/autoFRK.cpp:299 */
    int64_t n2 = nrow_of(xnew); int K = Rf_asInteger(K_);
    frk_plan *plan = plan_of(Xu, BBBH, UZ, nconst, K_);
    SEXP out = PROTECT(Rf_allocMatrix(REALSXP, (int)n2, K));
    check_plan(plan, frk_plan_predict(plan, REAL(xnew), n2, REAL(out), FRK_OUT_BASIS));
    frk_plan_destroy(plan);
    UNPROTECT(1);
    return out;
}

/* the sufficient statistics of cMLEimat on the mrts basis at `xnew` without forming it: list(G, C, zz)
 * (R/helper.R:20, R/autoFRK.R:431-435 fused behind R/autoFRK.R:1453-1466) */
SEXP _autoFRK_predict_gram(SEXP Xu, SEXP BBBH, SEXP UZ, SEXP nconst, SEXP K_, SEXP xnew, SEXP Data) {
    need_real_matrix(xnew, "xnew"); need_real_matrix(Data, "Data");
    if (ncol_of(xnew) != ncol_of(Xu)) Rf_error("xnew must have %d columns", ncol_of(Xu));
    if (nrow_of(Data) != nrow_of(xnew)) Rf_error("Data must have one row per location");
    int64_t n2 = nrow_of(xnew); int K = Rf_asInteger(K_), T = ncol_of(Data);
    frk_plan *plan = plan_of(Xu, BBBH, UZ, nconst, K_);
    const char *names[] = {"G", "C", "zz"};
    SEXP out = PROTECT(named_list(3, names));
    SEXP G = PROTECT(Rf_allocMatrix(REALSXP, K, K));
    SEXP C = PROTECT(Rf_allocMatrix(REALSXP, K, T));
    SEXP zz = PROTECT(Rf_allocVector(REALSXP, 1));
    check_plan(plan, frk_plan_predict_gram(plan, REAL(xnew), n2, REAL(Data), T, NULL, REAL(G), REAL(C), REAL(zz)));
    frk_plan_destroy(plan);
    SET_VECTOR_ELT(out, 0, G); SET_VECTOR_ELT(out, 1, C); SET_VECTOR_ELT(out, 2, zz);
    UNPROTECT(4);
    return out;
}

/* predict.FRK at new locations of an mrts model (R/autoFRK.R:1184,1216,1220): list(yhat n2 x T, se n2 or NULL);
 * V = R_NilValue skips the standard errors */
SEXP _autoFRK_predict_frk(SEXP Xu, SEXP BBBH, SEXP UZ, SEXP nconst, SEXP K_, SEXP xnew, SEXP w, SEXP V) {
    need_real_matrix(xnew, "xnew"); need_real_matrix(w, "w");
    int K = Rf_asInteger(K_), T = ncol_of(w);
    int64_t n2 = nrow_of(xnew);
    const int want_se = TYPEOF(V) == REALSXP;
    if (ncol_of(xnew) != ncol_of(Xu)) Rf_error("xnew must have %d columns", ncol_of(Xu));
    if (nrow_of(w) != K || (want_se && (Rf_nrows(V) != K || Rf_ncols(V) != K)))
        Rf_error("Dimensions of basis and object$w are not compatible!");
    frk_plan *plan = plan_of(Xu, BBBH, UZ, nconst, K_);
    const char *names[] = {"yhat", "se"};
    SEXP out = PROTECT(named_list(2, names));
    SEXP yhat = PROTECT(Rf_allocMatrix(REALSXP, (int)n2, T));
    SEXP se = PROTECT(want_se ? Rf_allocVector(REALSXP, n2) : R_NilValue);
    check_plan(plan, frk_plan_predict_frk(plan, REAL(xnew), n2, REAL(w), T, want_se ? REAL(V) : NULL, REAL(yhat),
                                          want_se ? REAL(se) : NULL));
    frk_plan_destroy(plan);
    SET_VECTOR_ELT(out, 0, yhat); SET_VECTOR_ELT(out, 1, se);
    UNPROTECT(3);
    return out;
}

/* Per-replicate NA-masked weighted Grams of the basis rows (R/autoFRK.R:1233-1235: G <- Fk[!miss[, tt], ],
 * De <- D0[!miss[, tt], !miss[, tt]]; the same products as EM0miss' BiDBt, :609-615): `observed` is an n x TT matrix of
 * 0/1 doubles, `Ddiag` = diag(D0).  Returns a (K*K) x TT matrix: column tt holds t(G) solve(De) G, column-major.
 * One pass over the rows of Fk for all replicates. */
SEXP _autoFRK_masked_grams(SEXP Fk, SEXP observed, SEXP Ddiag) {
    need_real_matrix(Fk, "Fk"); need_real_matrix(observed, "observed");
    R_xlen_t n = nrow_of(Fk); int K = ncol_of(Fk), T = ncol_of(observed);
    if (TYPEOF(Ddiag) != REALSXP || XLENGTH(Ddiag) != n || nrow_of(observed) != n)
        Rf_error("observed and Ddiag must have one row per row of Fk");
    SEXP out = PROTECT(Rf_allocMatrix(REALSXP, K * K, T));
    SEXP mask = PROTECT(Rf_allocVector(RAWSXP, n * T));
    SEXP zero = PROTECT(Rf_allocMatrix(REALSXP, (int)n, T));
    SEXP rest = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)K * T + 2 * (R_xlen_t)T));   /* c_all | zz_all | nobs_all (int64) */
    const double *o = REAL(observed);
    unsigned char *mk = RAW(mask);
    double *z = REAL(zero);
    for (R_xlen_t i = 0; i < n * T; ++i) { mk[i] = (o[i] != 0.0); z[i] = 0.0; }
    double *c_all = REAL(rest), *zz_all = c_all + (R_xlen_t)K * T;
    check(frk_masked_stats(z, mk, REAL(Fk), REAL(Ddiag), n, K, T, REAL(out), c_all, zz_all, (int64_t *)(zz_all + T)));
    UNPROTECT(4);
    return out;
}

/* V = M - t(G M) invCz(R, L, G M), L = G eigen(M)$vectors diag(sqrt(pmax(eigen(M)$values, 0)))  (R/autoFRK.R:1236-1240,
 * 1260-1271), as K x K algebra on Gg = t(G) solve(R) G. */
SEXP _autoFRK_krige_V(SEXP Gg, SEXP M) {
    need_real_matrix(Gg, "Gg"); need_real_matrix(M, "M");
    int K = Rf_nrows(M);
    if (Rf_ncols(M) != K || Rf_nrows(Gg) != K || Rf_ncols(Gg) != K) Rf_error("Gg and M must be K x K");
    SEXP V = PROTECT(Rf_allocMatrix(REALSXP, K, K));
    SEXP c = PROTECT(Rf_allocVector(REALSXP, 2 * (R_xlen_t)K));     /* a zero data column and its (unused) coefficients */
    for (int i = 0; i < 2 * K; ++i) REAL(c)[i] = 0.0;
    check(frk_krige_from_stats(REAL(Gg), REAL(c), 1, REAL(M), K, 1, REAL(c) + K, REAL(V)));
    UNPROTECT(2);
    return V;
}

static const R_CallMethodDef CallEntries[] = {
    {"_autoFRK_getASCeigens", (DL_FUNC)&_autoFRK_getASCeigens, 1},
    {"_autoFRK_mrtsrcpp", (DL_FUNC)&_autoFRK_mrtsrcpp, 3},
    {"_autoFRK_mrtsrcpp_predict0", (DL_FUNC)&_autoFRK_mrtsrcpp_predict0, 4},
    {"_autoFRK_mrtsrcpp_predict", (DL_FUNC)&_autoFRK_mrtsrcpp_predict, 7},
    {"_autoFRK_gram_stats", (DL_FUNC)&_autoFRK_gram_stats, 2},
    {"_autoFRK_getHalf_stats", (DL_FUNC)&_autoFRK_getHalf_stats, 1},
    {"_autoFRK_cMLEimat_stats", (DL_FUNC)&_autoFRK_cMLEimat_stats, 6},
    {"_autoFRK_cMLEimat_sweep", (DL_FUNC)&_autoFRK_cMLEimat_sweep, 7},
    {"_autoFRK_predict_mrts_basis", (DL_FUNC)&_autoFRK_predict_mrts_basis, 6},
    {"_autoFRK_predict_gram", (DL_FUNC)&_autoFRK_predict_gram, 7},
    {"_autoFRK_predict_frk", (DL_FUNC)&_autoFRK_predict_frk, 8},
    {"_autoFRK_masked_grams", (DL_FUNC)&_autoFRK_masked_grams, 3},
    {"_autoFRK_krige_V", (DL_FUNC)&_autoFRK_krige_V, 2},
    {NULL, NULL, 0}};

void R_init_autoFRK(DllInfo *dll) {
    R_registerRoutines(dll, NULL, CallEntries, NULL, NULL);
    R_useDynamicSymbols(dll, FALSE);
}
