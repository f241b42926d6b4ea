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
 *
 * Concurrency: the library is safe to call from several host threads at once (per-thread handles, workspaces and
 * default stream; a device-memory cache behind a mutex).  Scratch that consecutive `_dev` calls reuse (the Gram
 * workspace of a host thread, a plan's chunk buffers) is ordered across streams by events: a call issued on another
 * stream than the previous one first waits for that call's completion event, and buffers are only re-allocated after
 * it.  So one thread may alternate streams freely; what it gains from several streams is ordering flexibility, not
 * overlap of two calls that share a workspace -- use two host threads (and two plans) for that.
 * Limits: dimensions are int64_t / int32_t as declared; n2 < 2^31 rows per call for the TMA-staged Gram kernels
 * (larger inputs take the cp.async kernel); knots m <= 65535 (the reference caps them at maxknot = 5000).
 * Device memory a calling thread keeps between calls: the Gram workspace (partial tiles; up to about 1.9 GB once a
 * Gram with K >= ~250 over >= 2^20 rows has run), and per plan the 2 GB basis chunk of the fused predict -> Gram /
 * predict.FRK paths; host memory: two 128 MB pinned staging slots per thread that uploads pageable matrices and two
 * pinned result chunks per plan that returns a basis to pageable memory.
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
/* device temporaries are cached between calls; this returns the cached blocks to the driver */
int32_t frk_release_cached_memory(void);
/* Measurement aid (no reference counterpart): issue rate of the FP64 tensor instruction every kernel here is built
 * from (DMMA.8x8x4), in TFLOP/s on the current device: `iters` x 16 dependent-free DMMAs per warp, 16 warps per SM.
 * bench.py reports it beside the cuBLAS DGEMM figure as the roofline denominator's cross-check. */
int32_t frk_measure_dmma_peak(int32_t iters, double* tflops);

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

/* replaces _autoFRK_mrtsrcpp (src/RcppExports.cpp:28-39 -> src/autoFRK.cpp:224-242): the mrts fit.
 *   Xu n x d knots, xobs_diag xd_rows x xd_cols (must be d x d), k = number of eigen-columns (K - d - 1).
 * Outputs (caller-allocated): X n x (k+d+1), UZ (n+d+1) x (k+d+1), BBBH (d+1) x n, nconst d.
 * Eigenvector signs are arbitrary in the reference (Spectra); here every eigenvector is normalised so
 * that its largest-magnitude entry is positive, which makes the fit reproducible across ranks. */
int32_t frk_mrtsrcpp(const double* Xu, int64_t n, int32_t d, const double* xobs_diag, int32_t xd_rows,
                     int32_t xd_cols, int32_t k, double* X, double* UZ, double* BBBH, double* nconst);

/* replaces _autoFRK_mrtsrcpp_predict0 (src/RcppExports.cpp:41-53 -> src/autoFRK.cpp:248-281): fit, then
 * evaluate at xnew (n2 x xnew_cols, must have d columns).  X1 is n2 x k.  The eigenvector block never
 * leaves the device between the fit and the evaluation. */
int32_t frk_mrtsrcpp_predict0(const double* Xu, int64_t n, int32_t d, const double* xobs_diag, int32_t xd_rows,
                              int32_t xd_cols, const double* xnew, int64_t n2, int32_t xnew_cols, int32_t k,
                              double* X, double* UZ, double* BBBH, double* nconst, double* X1);

/* replaces _autoFRK_getASCeigens (src/RcppExports.cpp:17-26 -> src/autoFRK.cpp:24-32): full symmetric
 * eigendecomposition of the n x n matrix A (lower triangle read), values ascending, vectors n x n. */
int32_t frk_getASCeigens(const double* A, int32_t n, double* value, double* vector);

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

/* ---------------------------------------------------------------------------------------------
 * Fixed-rank-kriging reductions.  These are R code in the reference (base `%*%`, `eigen`); the
 * R-side functions named below become thin wrappers around them.
 * ------------------------------------------------------------------------------------------- */

/* Sufficient statistics of the fit in one pass over the rows:
 *   G = F'F (K x K; replaces `t(Fk) %*% iDFk`, R/helper.R:20), C = F'Z (K x T; all that
 *   R/autoFRK.R:433-435 needs because ihF %*% Data == half %*% (F'Z)), zz = sum(Z^2) (R/autoFRK.R:431).
 * F is n x K, Z is n x T (T may be 0, then Z, C may be NULL), host pointers, column-major.
 * `weights` (n, may be NULL): row weights w_i, giving G = F' diag(w) F, C = F' diag(w) Z, zz = sum(w Z^2) -- the
 * diagonal-D statistics `t(Fk) %*% iD %*% Fk` etc. of cMLEsp (R/autoFRK.R:528-538) and, with 0/1 weights, the
 * per-replicate NA-masked Grams of EM0miss (R/autoFRK.R:609-615). */
int32_t frk_gram_stats(const double* F, int64_t n, int32_t K, const double* Z, int32_t T, const double* weights,
                       double* G, double* C, double* zz);
/* length of the packed device result [G (K*K) | C (K*T) | zz (1)] */
int64_t frk_gram_packed_len(int32_t K, int32_t T);
/* Host-only introspection of how the tensor-core Gram kernel cuts `t(Fk) %*% iDFk` / `t(Fk) %*% Data`
 * (R/helper.R:20, R/autoFRK.R:433-435) into work: the 32-column block pairs (a <= b over F's blocks, then F x Z
 * with Z's blocks numbered after F's) and the job each pair is packed into.  Writes up to `cap` triples
 * (job, block_a, block_b) to `triples` (may be NULL) and returns the number of pairs; *njobs receives the job
 * count.  No device work. */
int64_t frk_gram_job_plan(int32_t K, int32_t T, int32_t* njobs, int32_t* triples, int64_t cap);
/* The same for the register-resident kernel that takes the shapes default autoFRK() selects (K + T <= 256 columns):
 * 8 x 8 output tiles dealt to 16 warps.  Writes up to `cap` quintuples (row group, column group, is_Z, warp,
 * accumulator slot) to `tiles` (may be NULL), the number of row-stage sets (1, 2 or 4) to *nsets and the DMMAs per
 * k4-step that land on each of the four SM sub-partitions to sp_load[4]; returns the tile count, or -1 when the
 * shape does not fit this kernel.  No device work. */
int64_t frk_gram_mid_plan(int32_t K, int32_t T, int32_t* nsets, int32_t* sp_load, int32_t* tiles, int64_t cap);
/* device pointers: F n x K (ld ldf), Z n x T (ld ldz) -> d_packed; asynchronous on `stream`.
 * Multi-GPU: every rank calls this on its row shard, then one allreduce(sum) of d_packed. */
int32_t frk_gram_stats_dev(const double* d_F, int64_t n, int32_t K, int64_t ldf, const double* d_Z, int32_t T,
                           int64_t ldz, const double* d_weights, double* d_packed, void* stream);
/* fused predict.mrts -> Gram: the n2 x K basis is produced chunk by chunk (chunk_rows rows, 0 = default)
 * and reduced at once, so it is never materialised.  d_Z is n2 x T (ld ldz). */
int32_t frk_plan_predict_gram_dev(frk_plan* plan, const double* d_xnew, int64_t n2, int64_t ldx, const double* d_Z,
                                  int32_t T, int64_t ldz, const double* d_weights, double* d_packed,
                                  int64_t chunk_rows, void* stream);

/* Multi-GPU combine step of the fit, as one kernel over NVLink peer memory instead of a library all-reduce:
 * d_ptrs[r] is rank r's packed statistics buffer mapped into THIS device's address space (symmetric memory /
 * CUDA IPC); d_out[i] = sum_r d_ptrs[r][i], added in rank order on every rank (bitwise-identical results).
 * The caller orders it after every rank's frk_gram_stats_dev (a device-side barrier over the same allocation). */
int32_t frk_peer_sum_dev(const void* const* d_ptrs, int32_t nranks, int64_t len, double* d_out, void* stream);

/* getHalf (R/helper.R:19-26) given G = t(Fk) %*% iDFk (K x K, ld ldg, lower triangle read): half = G^(-1/2),
 * zero eigenvalues -> 0. */
int32_t frk_getHalf(const double* G, int64_t ldg, int32_t K, double* half);

/* cMLEimat (R/autoFRK.R:399-480) from the sufficient statistics.  G (ld ldg) and C (ld ldc) may be the
 * leading K x K / K x T blocks of larger matrices (selectBasis' AIC sweep, R/autoFRK.R:305-317).
 * `n` = nrow(Fk).  Outputs: v, negloglik always; M (K x K) unless only_loglik; w (K x T), V (K x K) if wsave.
 * Unused outputs may be NULL.  `on_device` != 0: G and C are device pointers (outputs stay host pointers). */
int32_t frk_cMLEimat_stats(const double* G, int64_t ldg, const double* C, int64_t ldc, double zz, int32_t K, int32_t T,
                           int64_t n, double s, int32_t wsave, int32_t only_loglik, int32_t on_device, double* v,
                           double* negloglik, double* M, double* w, double* V);

/* The K loop of selectBasis (R/autoFRK.R:305-317): negloglik[i] = cMLEimat(Fk[, 1:Ks[i]], Data, s, wSave = FALSE)$negloglik
 * for every candidate basis size, from the max-K statistics G (>= max(Ks) square, ld ldg), C (ld ldc), zz (host
 * pointers).  One Cholesky factor of G serves every leading block (the eigenvalues of JSJ_K are those of the
 * min(K, T)-sized Gram of the leading K rows of L^-1 C), instead of two K x K eigen-decompositions per candidate.
 * Candidates whose leading block of G is not safely positive definite get NaN: call frk_cMLEimat_stats for them
 * (it follows getHalf's zero-eigenvalue rule, R/helper.R:22-24). */
int32_t frk_cMLEimat_sweep(const double* G, int64_t ldg, const double* C, int64_t ldc, double zz, int32_t T, int64_t n,
                           double s, const int32_t* Ks, int32_t nK, double* negloglik);

/* cMLE (R/autoFRK.R:333-397): half, JSJ are K x K host matrices; Fk (n x K, host, may be NULL) is only needed
 * for L (n x nL, written when wsave != 0 and Fk, L are non-NULL; L must hold n x K doubles). */
int32_t frk_cMLE(const double* Fk, int64_t n, int32_t K, int32_t TT, double trS, const double* half, const double* JSJ,
                 double s, double ldet, int32_t wsave, int32_t only_loglik, int32_t has_vfixed, double vfixed,
                 double* v, double* negloglik, double* M, double* L, int32_t* nL);

/* predict.FRK core (R/autoFRK.R:1216,1220) fused into the basis evaluation: yhat = basis %*% w (n2 x T),
 * se = sqrt(pmax(0, rowSums((basis %*% V) * basis))) (n2; V, se may be NULL; V symmetric, lower triangle read).
 * Host pointers.  w and the eigen-factor R of V = R diag(+-1) R' are folded through the projection onto the knot
 * eigenvectors (W' = W [w | R]), so one kernel pass with T + rank(V) columns produces both and the n2 x K basis
 * is never formed; se^2 is accumulated as a signed sum of squares. */
int32_t frk_plan_predict_frk(frk_plan* plan, const double* xnew, int64_t n2, const double* w, int32_t T,
                             const double* V, double* yhat, double* se);

/* invCz (R/autoFRK.R:845-852): out = (R + L L')^{-1} z for a DIAGONAL R given by its diagonal Rdiag (n);
 * L is n x K, z and out are n x T (host, column-major).  ZinvC (R/helper.R:249-256) is its transpose. */
int32_t frk_invCz(const double* Rdiag, const double* L, const double* z, int64_t n, int32_t K, int32_t T, double* out);

/* getLikelihood (R/helper.R:28-58) for a DIAGONAL Depsilon (Ddiag, n): -2 log-likelihood of Data given Fk, M, s.
 * Data0 is Data with its NAs replaced by 0 and `observed` (n x T bytes, column-major) is !is.na(Data). */
int32_t frk_getLikelihood(const double* Data0, const uint8_t* observed, const double* Fk, const double* M, double s,
                          const double* Ddiag, int64_t n, int32_t K, int32_t T, double* n2loglik);

/* predict.FRK with new observations (R/autoFRK.R:1248-1274) as K x K algebra on the statistics of the basis at
 * the observation sites: Gg = t(G) iR G (K x K), c = t(G) iR obsData (K x Tc), R = s * De diagonal.  Returns
 *   coef = t(GM) %*% invCz(R, L, obsData)      (K x Tc;  yhat = basis %*% coef, :1265)
 *   V    = M - t(GM) %*% invCz(R, L, GM)       (K x K, if want_V; :1270)
 * with L = G %*% dec$vector %*% diag(sqrt(pmax(dec$value, 0))), dec = eigen(M) (:1260-1264). */
int32_t frk_krige_from_stats(const double* Gg, const double* c, int32_t Tc, const double* M, int32_t K, int32_t want_V,
                             double* coef, double* V);

/* host-buffer version of frk_plan_predict_gram_dev: statistics of predict.mrts(object, xnew) against Z with
 * optional row weights, without ever materialising the n2 x K basis. */
int32_t frk_plan_predict_gram(frk_plan* plan, const double* xnew, int64_t n2, const double* Z, int32_t T,
                              const double* weights, double* G, double* C, double* zz);

/* EM0miss (R/autoFRK.R:560-737; diagonal Depsilon, DfromLK = NULL) in two steps.
 * (1) per-replicate NA-masked statistics (:587-616) in one pass over the rows of Fk:
 *     G_all[, , t] = t(Bt) iDt Bt (K x K), c_all[, t] = t(iDBt) zt (K), zz_all[t] = t(zt) iDt zt, nobs_all[t] = sum(O[, t]);
 *     Data0 = Data with NA -> 0, observed = !is.na(Data) as bytes, Ddiag = diag(Depsilon). */
int32_t frk_masked_stats(const double* Data0, const uint8_t* observed, const double* Fk, const double* Ddiag, int64_t n,
                         int32_t K, int32_t T, double* G_all, double* c_all, double* zz_all, int64_t* nobs_all);
/* (2) the EM iteration (:652-687) from those statistics, started at M0 = cMLEimat(Fk, Z0, 0, TRUE)$M and s0
 *     (its v, or vfixed): returns M (K x K), s, etatt (K x T) and the iteration count. */
int32_t frk_EM0miss_stats(const double* G_all, const double* c_all, const double* zz_all, int64_t sumO, int32_t K, int32_t T,
                          const double* M0, double s0, int32_t maxit, double avgtol, int32_t has_vfixed, double vfixed,
                          double* M_out, double* s_out, double* etatt_out, int32_t* niter);

/* kNN mean imputation, the method = "fast" pre-step of autoFRK() / selectBasis() (R/autoFRK.R:135-148, :273-286):
 * every NA of Data (n x T, column-major, modified in place) becomes the mean of the n.neighbor = k nearest
 * observed values of its replicate (FNN::get.knnx on loc, n x d). */
int32_t frk_knn_impute(const double* loc, int64_t n, int32_t d, double* Data, int32_t T, int32_t k);

/* the same two products for a basis matrix the caller already holds (predict.FRK with `basis=` or with
 * newloc = NULL, where R uses object$G): F is n x K, host, column-major; streamed through the device. */
int32_t frk_basis_predict_frk(const double* F, int64_t n, int32_t K, const double* w, int32_t T, const double* V,
                              double* yhat, double* se);

#ifdef __cplusplus
}
#endif
#endif /* FRK_B200_H */
