/* Stand-in for Spectra's <SymEigs.h> (pre-1.0 pointer-op API, as RSpectra ships it) -- TEST INFRASTRUCTURE.
 *
 * Spectra is not vendored in the reference and absent from this image.  The reference asks it for the k
 * algebraically largest eigenpairs of a dense symmetric matrix (src/autoFRK.cpp:39-54: LARGEST_ALGE, ncv, up to
 * 1000 restarts, tolerance 1e-10) and receives them in DESCENDING order.  This stand-in answers the same question
 * with a full Jacobi decomposition (frk_stub_eigen.h) instead of implicitly restarted Lanczos: the converged
 * eigenpairs are the same mathematical objects to the solver's tolerance, eigenvector signs are arbitrary in both.
 * The constructor keeps Spectra's documented argument contract (1 <= nev <= n-1, nev < ncv <= n).
 */
#ifndef FRK_STUB_SYMEIGS_H
#define FRK_STUB_SYMEIGS_H
#include <stdexcept>
#include <vector>

#include "frk_stub_eigen.h"

namespace Spectra {

enum SELECT_EIGENVALUE { LARGEST_MAGN = 0, LARGEST_REAL, LARGEST_IMAG, LARGEST_ALGE, SMALLEST_MAGN, SMALLEST_REAL,
                         SMALLEST_IMAG, SMALLEST_ALGE, BOTH_ENDS };
enum COMPUTATION_INFO { SUCCESSFUL = 0, NOT_COMPUTED, NOT_CONVERGING, NUMERICAL_ISSUE };

template <typename Scalar> class DenseSymMatProd {
    const Eigen::MatrixXd &m_;

  public:
    explicit DenseSymMatProd(const Eigen::MatrixXd &m) : m_(m) {}
    Eigen::Index rows() const { return m_.rows(); }
    Eigen::Index cols() const { return m_.cols(); }
    const Eigen::MatrixXd &matrix() const { return m_; }
};

template <typename Scalar, int SelectionRule, typename OpType> class SymEigsSolver {
    OpType *op_;
    int nev_, ncv_, info_;
    Eigen::VectorXd values_;
    Eigen::MatrixXd vectors_;

  public:
    SymEigsSolver(OpType *op, int nev, int ncv) : op_(op), nev_(nev), ncv_(ncv), info_(NOT_COMPUTED) {
        const int n = static_cast<int>(op->rows());
        if (nev < 1 || nev > n - 1) throw std::invalid_argument("nev must satisfy 1 <= nev <= n - 1, n is the size of matrix");
        if (ncv <= nev || ncv > n) throw std::invalid_argument("ncv must satisfy nev < ncv <= n, n is the size of matrix");
        static_assert(SelectionRule == LARGEST_ALGE, "the stand-in only implements the rule the reference uses");
    }
    void init() {}
    int compute(int /*maxit*/ = 1000, Scalar /*tol*/ = 1e-10) {
        std::vector<double> w;
        Eigen::MatrixXd V;
        Eigen::stubdetail::jacobi_eigh(op_->matrix(), w, V);
        const Eigen::Index n = V.rows();
        values_ = Eigen::VectorXd(nev_);
        vectors_ = Eigen::MatrixXd(n, nev_);
        for (int j = 0; j < nev_; ++j) {                  /* descending: largest first */
            values_(j) = w[static_cast<size_t>(n - 1 - j)];
            for (Eigen::Index i = 0; i < n; ++i) vectors_(i, j) = V(i, n - 1 - j);
        }
        info_ = SUCCESSFUL;
        return nev_;
    }
    int info() const { return info_; }
    Eigen::VectorXd eigenvalues() const { return values_; }
    Eigen::MatrixXd eigenvectors() const { return vectors_; }
};

}  // namespace Spectra
#endif
