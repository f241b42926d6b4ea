/* frk_b200 -- C ABI of the B200-native autoFRK hot path (mrts basis evaluation + FRK reductions).
 *
 * Drop-in boundary: these are the entry points an autoFRK `.Call` shim binds instead of the
 * package's Rcpp/Eigen code (r/frk_b200_shim.c shows the binding; INTEGRATION.md the R side).
 * Conventions, chosen to match what R hands to `.Call`:
 *   - every matrix is FP64, COLUMN-MAJOR, dense (leading dimension = number of rows unless an
 *     explicit `ld*` argument says otherwise);
 *   - inputs are borrowed for the duration of the call and never written;
 *   - outputs are caller-allocated;
 *   - every function returns an int status (FRK_OK == 0); on failure `frk_last_error()` holds the
 *     message (thread-local).  The argument-check messages are the reference's `Rcpp::stop` texts
 *     (src/autoFRK.cpp:49,65,...,308) so the shim can forward them to `Rf_error()` verbatim;
 *   - functions without a `_dev` suffix take HOST pointers and are synchronous (the result is in
 *     host memory on return, as `.Call` requires); `_dev` functions take DEVICE pointers and a
 *     `cudaStream_t` (passed as void*) and are asynchronous on that stream.
 * There is no CPU fallback: without a CUDA device every compute entry point fails with FRK_ECUDA.
 */
#ifndef FRK_B200_H
#define FRK_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FRK_OK 0
#define FRK_EINVAL 1    /* argument check failed (an Rcpp::stop site of the reference) */
#define FRK_EINTERNAL 2
#define FRK_ECUDA 3     /* CUDA runtime / library failure, or no device */
#define FRK_ENOTCONV 4  /* eigensolver did not converge (src/autoFRK.cpp:48-50) */

#define FRK_OUT_X1 0    /* n2 x k  : what mrtsrcpp_predict returns as `X1` (src/autoFRK.cpp:324) */
#define FRK_OUT_BASIS 1 /* n2 x k  : cbind(1, (x - shift)/nconst, X1[, 1:(k-d-1)])  (R/autoFRK.R:1460-1466) */

const char* frk_last_error(void);
int32_t frk_version(void);
int32_t frk_device_count(int32_t* count);
int32_t frk_set_device(int32_t device);

/* ---------------------------------------------------------------------------------------------
 * Reference boundary, one function per registered `.Call` routine (src/RcppExports.cpp:71-77).
 * ------------------------------------------------------------------------------------------- */

/* replaces _autoFRK_mrtsrcpp_predict (src/RcppExports.cpp:55-69 -> src/autoFRK.cpp:287-326).
 *   Xu        n x d knots               xobs_diag  d x d (carried, never read: src/autoFRK.cpp:288)
 *   xnew      n2 x xnew_cols            BBBH       bbbh_rows x bbbh_cols   (must be (d+1) x n)
 *   UZ        uz_rows x uz_cols         nconst     nconst_len              (must be d)
 *   X1 (out)  n2 x k, = Phi(xnew,Xu) * UZ[0:n,0:k] - [1,xnew] * (BBBH * UZ[0:n,0:k])
 * The pass-through list members X, UZ, BBBH, nconst (src/autoFRK.cpp:320-323) are the caller's own
 * inputs and are not copied. */
int32_t frk_mrtsrcpp_predict(const double* Xu, int64_t n, int32_t d, const double* xobs_diag,
                             const double* xnew, int64_t n2, int32_t xnew_cols, const double* BBBH,
                             int64_t bbbh_rows, int64_t bbbh_cols, const double* UZ, int64_t uz_rows,
                             int64_t uz_cols, const double* nconst, int64_t nconst_len, int32_t k,
                             double* X1);

/* ---------------------------------------------------------------------------------------------
 * Plans: keep knots / eigenvector block / polynomial correction resident on the device so that
 * repeated predict.mrts() calls (and the multi-GPU row shards) pay the upload and packing once.
 * ------------------------------------------------------------------------------------------- */
typedef struct frk_plan frk_plan;

/* Xu n x d, BBBH (d+1) x n, UZ uz_rows x uz_cols (uz_rows >= n, uz_cols >= k), nconst d (may be
 * NULL if FRK_OUT_BASIS is never requested).  `on_device` != 0: the four arrays are device
 * pointers on the current device (nconst stays a host pointer). */
int32_t frk_plan_create(frk_plan** plan, const double* Xu, int64_t n, int32_t d, const double* BBBH,
                        const double* UZ, int64_t uz_rows, int64_t uz_cols, const double* nconst,
                        int32_t k, int32_t on_device);
int32_t frk_plan_destroy(frk_plan* plan);
/* info[0..7] = d, knots, k, k_compute (leading non-zero columns), tile rows, slice columns, slices, knot chunks */
int32_t frk_plan_info(const frk_plan* plan, int64_t info[8]);
/* kernels launched through this plan so far */
int64_t frk_plan_launches(const frk_plan* plan);

/* host buffers; H2D of xnew, chunked compute overlapped with D2H of the result; synchronous */
int32_t frk_plan_predict(frk_plan* plan, const double* xnew, int64_t n2, double* out, int32_t out_mode);
/* device buffers: xnew n2 x d (ld ldx), out n2 x k (ld ldo); asynchronous on `stream` */
int32_t frk_plan_predict_dev(frk_plan* plan, const double* d_xnew, int64_t n2, int64_t ldx, double* d_out,
                             int64_t ldo, int32_t out_mode, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FRK_B200_H */
