/* C entry points around the reference's OWN translation unit -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * FRK_REF_SRC is the path of the upstream src/autoFRK.cpp (given by oracle/Makefile, normally
 * /root/reference/src/autoFRK.cpp); it is #included verbatim, never copied into this repository.  Its three
 * third-party includes (<Rcpp.h>, <RcppEigen.h>, <SymEigs.h>) resolve to the stand-ins under oracle/stub/ --
 * read oracle/stub/frk_stub_eigen.h for what that does and does not pin.
 *
 * Every function returns 0 or 1; on 1 the text the reference passed to Rcpp::stop (or the stand-in's
 * std::exception) is available from ref_last_error().  The list-returning .Call targets give back an opaque
 * handle whose items are read with ref_list_*.
 */
#ifndef FRK_REF_SRC
#error "compile with -DFRK_REF_SRC='\"/path/to/reference/src/autoFRK.cpp\"'"
#endif
#include FRK_REF_SRC

#include <cstdint>
#include <cstring>
#include <string>

namespace {
thread_local std::string g_err;
typedef Eigen::Map<Eigen::MatrixXd> MapM;
typedef Eigen::Map<Eigen::VectorXd> MapV;

template <class F> int guarded(F f) {
    try {
        f();
        return 0;
    } catch (const std::exception &e) {
        g_err = e.what();
    } catch (...) {
        g_err = "unknown C++ exception";
    }
    return 1;
}
template <class F> void *guarded_list(F f) {
    Rcpp::List *out = nullptr;
    if (guarded([&] { out = new Rcpp::List(f()); })) return nullptr;
    return out;
}
}  // namespace

extern "C" {

const char *ref_last_error(void) { return g_err.c_str(); }
const char *ref_source_path(void) { return FRK_REF_SRC; }

/* src/autoFRK.cpp:60-103.  L (p x p, column-major) is zero-filled here the way mrts() does at :183 before the call;
 * only the strict upper triangle is written by the reference. */
int ref_tpm2(const double *P, int64_t p, int pcols, int d, double *L) {
    return guarded([&] {
        std::memset(L, 0, sizeof(double) * static_cast<size_t>(p) * static_cast<size_t>(p));
        MapM Pm(P, p, pcols), Lm(L, p, p);
        tpm2(Pm, Lm, d);
    });
}

/* src/autoFRK.cpp:109-153.  L (p1 x p2) is zero-filled here as `Hnew = MatrixXd::Zero(n2, n)` does at :311. */
int ref_tpm_predict(const double *Pnew, int64_t p1, const double *P, int64_t p2, int pcols, int d, double *L) {
    return guarded([&] {
        std::memset(L, 0, sizeof(double) * static_cast<size_t>(p1) * static_cast<size_t>(p2));
        MapM Pn(Pnew, p1, pcols), Pm(P, p2, pcols), Lm(L, p1, p2);
        tpm_predict(Pn, Pm, Lm, d);
    });
}

/* the shape checks of tpm2 / tpm_predict (:64-72, :114-122) need inputs of the caller's shapes */
int ref_tpm_predict_shaped(const double *Pnew, int64_t p1, int64_t newcols, const double *P, int64_t p2, int64_t pcols,
                           int d, double *L, int64_t lrows, int64_t lcols) {
    return guarded([&] {
        MapM Pn(Pnew, p1, newcols), Pm(P, p2, pcols), Lm(L, lrows, lcols);
        tpm_predict(Pn, Pm, Lm, d);
    });
}
int ref_tpm2_shaped(const double *P, int64_t p, int64_t pcols, int d, double *L, int64_t lrows, int64_t lcols) {
    return guarded([&] {
        MapM Pm(P, p, pcols), Lm(L, lrows, lcols);
        tpm2(Pm, Lm, d);
    });
}

/* src/autoFRK.cpp:24-32 */
void *ref_getASCeigens(const double *A, int64_t n) {
    return guarded_list([&] { return getASCeigens(MapM(A, n, n)); });
}

/* src/autoFRK.cpp:224-242 */
void *ref_mrtsrcpp(const double *Xu, int64_t n, int64_t d, const double *xobs_diag, int64_t xr, int64_t xc, int k) {
    return guarded_list([&] { return mrtsrcpp(MapM(Xu, n, d), MapM(xobs_diag, xr, xc), k); });
}

/* src/autoFRK.cpp:248-281 */
void *ref_mrtsrcpp_predict0(const double *Xu, int64_t n, int64_t d, const double *xobs_diag, int64_t xr, int64_t xc,
                            const double *xnew, int64_t n2, int64_t dnew, int k) {
    return guarded_list(
        [&] { return mrtsrcpp_predict0(MapM(Xu, n, d), MapM(xobs_diag, xr, xc), MapM(xnew, n2, dnew), k); });
}

/* src/autoFRK.cpp:287-326 */
void *ref_mrtsrcpp_predict(const double *Xu, int64_t n, int64_t d, const double *xobs_diag, int64_t xr, int64_t xc,
                           const double *xnew, int64_t n2, int64_t dnew, const double *BBBH, int64_t br, int64_t bc,
                           const double *UZ, int64_t ur, int64_t uc, const double *nconst, int64_t nn, int k) {
    return guarded_list([&] {
        return mrtsrcpp_predict(MapM(Xu, n, d), MapM(xobs_diag, xr, xc), MapM(xnew, n2, dnew), MapM(BBBH, br, bc),
                                MapM(UZ, ur, uc), MapV(nconst, nn), k);
    });
}

int64_t ref_list_size(const void *h) { return static_cast<int64_t>(static_cast<const Rcpp::List *>(h)->size()); }
const char *ref_list_name(const void *h, int64_t i) {
    return static_cast<const Rcpp::List *>(h)->at(static_cast<size_t>(i)).name.c_str();
}
void ref_list_dims(const void *h, int64_t i, int64_t *rows, int64_t *cols) {
    const Eigen::MatrixXd &m = static_cast<const Rcpp::List *>(h)->at(static_cast<size_t>(i)).value;
    *rows = m.rows();
    *cols = m.cols();
}
void ref_list_copy(const void *h, int64_t i, double *dst) {
    const Eigen::MatrixXd &m = static_cast<const Rcpp::List *>(h)->at(static_cast<size_t>(i)).value;
    if (m.size() > 0) std::memcpy(dst, m.data(), sizeof(double) * static_cast<size_t>(m.size()));
}
void ref_list_free(void *h) { delete static_cast<Rcpp::List *>(h); }

}  // extern "C"
