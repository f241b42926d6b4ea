/* Plain-C restatement of the reference's scalar TPS kernel loops -- TEST INFRASTRUCTURE, NOT
 * PRODUCT CODE (see oracle/__init__.py).  Pinned: tests/test_oracle_ref.py checks it bit for bit
 * (0 ulp, r == 0 included) against the reference's own loops compiled unmodified (oracle/_ref) and
 * against the vectors those produced (tests/golden/ref_vectors.npz).
 *
 * Follows (column-major storage, like Eigen::MatrixXd / R):
 *   oracle_tpm_predict : src/autoFRK.cpp:109-153  (same loop order: i outer, j inner)
 *   oracle_tpm2        : src/autoFRK.cpp:60-103 + the symmetrisation at :189
 *   oracle_predict_x1  : src/autoFRK.cpp:311-324  (Hnew * UZ.block(0,0,n,k) - [1,xnew]*(BBBH*UZ.block))
 *                        naive triple loop, for small cross-checks only.
 * The reference's loops are single-threaded and so is this file.  To give the CPU baseline every
 * host core, the Python harness calls oracle_tpm_predict on disjoint row blocks from several
 * threads (ctypes drops the GIL); `ldn` / `ldl` exist so a row block can be addressed in place.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* L is p1 x p2, column-major, ld = ldl; P_new is p1 x d (ld ldn); P is p2 x d (ld p2). */
int oracle_tpm_predict(const double *P_new, int64_t p1, int64_t ldn, const double *P, int64_t p2,
                       int d, double *L, int64_t ldl) {
    if (d < 1 || d > 3) return 1;                       /* :114-116 */
    if (d == 1) {
        for (int64_t i = 0; i < p1; i++)
            for (int64_t j = 0; j < p2; ++j) {
                double r = fabs(P_new[i] - P[j]);       /* :127 */
                L[i + j * ldl] = pow(r, 3) / 12;         /* :128 */
            }
    } else if (d == 2) {
        for (int64_t i = 0; i < p1; i++)
            for (int64_t j = 0; j < p2; ++j) {
                double r = sqrt(pow(P_new[i] - P[j], 2) + pow(P_new[i + ldn] - P[j + p2], 2)); /* :135-136 */
                if (r != 0)
                    L[i + j * ldl] = r * r * log(r) / (8.0 * M_PI);  /* :137-138 */
                else
                    L[i + j * ldl] = 0.0;                /* Hnew = Zero(n2, n) at :311 */
            }
    } else {
        for (int64_t i = 0; i < p1; i++)
            for (int64_t j = 0; j < p2; ++j) {
                double r = sqrt(pow(P_new[i] - P[j], 2) + pow(P_new[i + ldn] - P[j + p2], 2) +
                                pow(P_new[i + 2 * ldn] - P[j + 2 * p2], 2));     /* :145-147 */
                L[i + j * ldl] = -r / 8;                 /* :148 */
            }
    }
    return 0;
}

/* H is p x p column-major: strict upper triangle by the reference loop, then H += H' (:189). */
int oracle_tpm2(const double *P, int64_t p, int d, double *H) {
    if (d < 1 || d > 3) return 1;                       /* :64-66 */
    memset(H, 0, sizeof(double) * (size_t)p * (size_t)p);   /* :183 */
    for (int64_t i = 0; i < p; i++)
        for (int64_t j = i + 1; j < p; ++j) {
            double r, v = 0.0;
            if (d == 1) {
                r = fabs(P[i] - P[j]);                  /* :77 */
                v = pow(r, 3) / 12;                     /* :78 */
            } else if (d == 2) {
                r = sqrt(pow(P[i] - P[j], 2) + pow(P[i + p] - P[j + p], 2));    /* :85-86 */
                if (r != 0) v = r * r * log(r) / (8.0 * M_PI);                  /* :87-88 */
            } else {
                r = sqrt(pow(P[i] - P[j], 2) + pow(P[i + p] - P[j + p], 2) +
                         pow(P[i + 2 * p] - P[j + 2 * p], 2));                  /* :95-97 */
                v = -r / 8;                             /* :98 */
            }
            H[i + j * p] = v;
            H[j + i * p] = v;                           /* :189 */
        }
    return 0;
}

/* X1 (n2 x k, ld n2) = Hnew * UZ[0:n,0:k] - [1,xnew] * (BBBH * UZ[0:n,0:k]).  UZ has ld = ldu. */
int oracle_predict_x1(const double *Xu, int64_t n, int d, const double *xnew, int64_t n2,
                      const double *BBBH, const double *UZ, int64_t ldu, int k, double *X1) {
    double *Hnew = (double *)malloc(sizeof(double) * (size_t)n2 * (size_t)n);
    double *Cp = (double *)calloc((size_t)(d + 1) * (size_t)k, sizeof(double));
    if (!Hnew || !Cp) { free(Hnew); free(Cp); return 2; }
    int rc = oracle_tpm_predict(xnew, n2, n2, Xu, n, d, Hnew, n2);               /* :311-313 */
    if (rc) { free(Hnew); free(Cp); return rc; }
    for (int c = 0; c < k; c++)                                             /* BBBH * UZ.block */
        for (int a = 0; a <= d; a++) {
            double s = 0.0;
            for (int64_t j = 0; j < n; j++) s += BBBH[a + j * (d + 1)] * UZ[j + (int64_t)c * ldu];
            Cp[a + c * (d + 1)] = s;
        }
    for (int c = 0; c < k; c++)
        for (int64_t i = 0; i < n2; i++) {
            double s = 0.0;
            for (int64_t j = 0; j < n; j++) s += Hnew[i + j * n2] * UZ[j + (int64_t)c * ldu];  /* :316 */
            double b = Cp[c * (d + 1)];
            for (int a = 0; a < d; a++) b += xnew[i + a * n2] * Cp[a + 1 + c * (d + 1)];
            X1[i + (int64_t)c * n2] = s - b;                                                 /* :324 */
        }
    free(Hnew); free(Cp);
    return 0;
}
