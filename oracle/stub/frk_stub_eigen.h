/* Stand-in for the parts of Eigen that /root/reference/src/autoFRK.cpp uses -- TEST INFRASTRUCTURE.
 *
 * Eigen (via RcppEigen) is not vendored in the reference and is absent from this image, so the
 * reference translation unit cannot see the real headers.  This file supplies, under the same names,
 * exactly the interface that translation unit touches, so that the reference source compiles
 * UNMODIFIED from where it lies (oracle/Makefile, target _ref/libautofrk_ref.so).
 *
 * What that pins and what it does not:
 *   - the reference's own text is what runs: loop order, scalar libm expressions, index arithmetic,
 *     block offsets, argument checks and messages, the order of the matrix operations;
 *   - the dense algebra behind those names is THIS file's (eager, column-major, sequential sums in
 *     increasing index order; Cholesky; cyclic Jacobi for the symmetric eigenproblems), not Eigen's
 *     blocked/vectorised kernels, so results can differ from a real build at rounding level
 *     (summation order) and eigenvector signs are whatever Jacobi returns.  The scalar kernel loops
 *     tpm2 / tpm_predict (src/autoFRK.cpp:60-153) use nothing but operator()(i,j), rows(), cols():
 *     for them the compiled code is the reference's arithmetic bit for bit.
 * Semantics follow Eigen's documented behaviour (column-major storage; SelfAdjointEigenSolver returns
 * ascending eigenvalues with orthonormal eigenvectors and reads the lower triangle; LLT is A = LL';
 * vector = row-vector assignment transposes implicitly; array expressions assign to matrices).
 */
#ifndef FRK_STUB_EIGEN_H
#define FRK_STUB_EIGEN_H
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstring>
#include <numeric>
#include <stdexcept>
#include <vector>

namespace Eigen {

typedef std::ptrdiff_t Index;
enum UpLoType { Lower = 1, Upper = 2 };

class MatrixXd;
class ArrayXXd;

namespace stubdetail {
void jacobi_eigh(const MatrixXd &A, std::vector<double> &w, MatrixXd &V);   /* ascending */
}

/* Dense column-major matrix of doubles: owns its storage, or (through Map) views the caller's. */
class MatrixXd {
  protected:
    Index r_, c_;
    double *p_;
    std::vector<double> own_;
    void take(Index r, Index c) {
        r_ = r; c_ = c;
        own_.assign(static_cast<size_t>(r) * static_cast<size_t>(c), 0.0);
        p_ = own_.data();
    }
    void copy_from(const MatrixXd &o) {
        take(o.r_, o.c_);
        if (!own_.empty()) std::memcpy(p_, o.p_, sizeof(double) * own_.size());
    }
    struct ViewTag {};
    MatrixXd(ViewTag, double *p, Index r, Index c) : r_(r), c_(c), p_(p) {}

  public:
    MatrixXd() : r_(0), c_(0), p_(nullptr) {}
    MatrixXd(Index r, Index c) { take(r, c); }
    MatrixXd(const MatrixXd &o) { copy_from(o); }
    MatrixXd &operator=(const MatrixXd &o) {
        if (this != &o) {
            if (own_.empty() && p_ && r_ == o.r_ && c_ == o.c_)        /* a Map keeps viewing its buffer */
                std::memcpy(p_, o.p_, sizeof(double) * static_cast<size_t>(r_) * static_cast<size_t>(c_));
            else
                copy_from(o);
        }
        return *this;
    }
    MatrixXd(const ArrayXXd &a);
    template <int UpLo> class SelfAdjointView;
    template <int UpLo> MatrixXd(const SelfAdjointView<UpLo> &s);

    Index rows() const { return r_; }
    Index cols() const { return c_; }
    Index size() const { return r_ * c_; }
    double *data() { return p_; }
    const double *data() const { return p_; }
    double &operator()(Index i, Index j) { return p_[i + j * r_]; }
    const double &operator()(Index i, Index j) const { return p_[i + j * r_]; }

    static MatrixXd Constant(Index r, Index c, double v) {
        MatrixXd m(r, c);
        std::fill(m.p_, m.p_ + m.size(), v);
        return m;
    }
    static MatrixXd Zero(Index r, Index c) { return MatrixXd(r, c); }
    static MatrixXd Ones(Index r, Index c) { return Constant(r, c, 1.0); }
    MatrixXd &setZero() {
        std::fill(p_, p_ + size(), 0.0);
        return *this;
    }
    MatrixXd &noalias() { return *this; }
    MatrixXd eval() const { return *this; }
    MatrixXd transpose() const {
        MatrixXd t(c_, r_);
        for (Index j = 0; j < c_; ++j)
            for (Index i = 0; i < r_; ++i) t(j, i) = (*this)(i, j);
        return t;
    }
    MatrixXd &operator+=(const MatrixXd &o) {
        if (o.r_ != r_ || o.c_ != c_) throw std::invalid_argument("stub Eigen: += size mismatch");
        for (Index t = 0; t < size(); ++t) p_[t] += o.p_[t];
        return *this;
    }
    MatrixXd &operator/=(double s) {
        for (Index t = 0; t < size(); ++t) p_[t] /= s;
        return *this;
    }

    /* writable window into a matrix (block / rightCols on a non-const object) */
    class BlockRef {
        MatrixXd &m_;
        Index i0_, j0_, p_, q_;

      public:
        BlockRef(MatrixXd &m, Index i0, Index j0, Index p, Index q) : m_(m), i0_(i0), j0_(j0), p_(p), q_(q) {
            if (i0 < 0 || j0 < 0 || p < 0 || q < 0 || i0 + p > m.rows() || j0 + q > m.cols())
                throw std::out_of_range("stub Eigen: block out of range");
        }
        BlockRef &operator=(const MatrixXd &o) {
            if (o.rows() != p_ || o.cols() != q_) throw std::invalid_argument("stub Eigen: block = size mismatch");
            for (Index j = 0; j < q_; ++j)
                for (Index i = 0; i < p_; ++i) m_(i0_ + i, j0_ + j) = o(i, j);
            return *this;
        }
        BlockRef &operator=(const BlockRef &o) { return *this = static_cast<MatrixXd>(o); }
        BlockRef &operator=(const ArrayXXd &a);
        operator MatrixXd() const {
            MatrixXd out(p_, q_);
            for (Index j = 0; j < q_; ++j)
                for (Index i = 0; i < p_; ++i) out(i, j) = m_(i0_ + i, j0_ + j);
            return out;
        }
    };
    BlockRef block(Index i0, Index j0, Index p, Index q) { return BlockRef(*this, i0, j0, p, q); }
    MatrixXd block(Index i0, Index j0, Index p, Index q) const {
        return static_cast<MatrixXd>(BlockRef(const_cast<MatrixXd &>(*this), i0, j0, p, q));
    }
    BlockRef rightCols(Index q) { return BlockRef(*this, 0, c_ - q, r_, q); }
    MatrixXd rightCols(Index q) const { return block(0, c_ - q, r_, q); }

    /* m.selfadjointView<Lower>().rankUpdate(u): lower triangle of m += u u' */
    template <int UpLo> class SelfAdjointView {
        MatrixXd &m_;

      public:
        explicit SelfAdjointView(MatrixXd &m) : m_(m) {
            if (m.rows() != m.cols()) throw std::invalid_argument("stub Eigen: selfadjointView of a non-square matrix");
        }
        SelfAdjointView &rankUpdate(const MatrixXd &u, double alpha = 1.0) {
            if (u.rows() != m_.rows()) throw std::invalid_argument("stub Eigen: rankUpdate size mismatch");
            for (Index j = 0; j < m_.cols(); ++j)
                for (Index i = 0; i < m_.rows(); ++i) {
                    const bool mine = (UpLo == Lower) ? (i >= j) : (i <= j);
                    if (!mine) continue;
                    double s = 0.0;
                    for (Index t = 0; t < u.cols(); ++t) s += u(i, t) * u(j, t);
                    m_(i, j) += alpha * s;
                }
            return *this;
        }
        double coeff(Index i, Index j) const {
            const bool stored = (UpLo == Lower) ? (i >= j) : (i <= j);
            return stored ? m_(i, j) : m_(j, i);
        }
        Index rows() const { return m_.rows(); }
    };
    template <int UpLo> SelfAdjointView<UpLo> selfadjointView() { return SelfAdjointView<UpLo>(*this); }

    ArrayXXd array() const;

    /* m.rowwise() - rowvector ; m.colwise().mean() / .norm() */
    class RowwiseOp {
        const MatrixXd &m_;

      public:
        explicit RowwiseOp(const MatrixXd &m) : m_(m) {}
        MatrixXd operator-(const MatrixXd &row) const {
            if (row.rows() != 1 || row.cols() != m_.cols()) throw std::invalid_argument("stub Eigen: rowwise - needs a row vector");
            MatrixXd out(m_.rows(), m_.cols());
            for (Index j = 0; j < m_.cols(); ++j)
                for (Index i = 0; i < m_.rows(); ++i) out(i, j) = m_(i, j) - row(0, j);
            return out;
        }
    };
    class ColwiseOp {
        const MatrixXd &m_;

      public:
        explicit ColwiseOp(const MatrixXd &m) : m_(m) {}
        MatrixXd mean() const {
            MatrixXd out(1, m_.cols());
            for (Index j = 0; j < m_.cols(); ++j) {
                double s = 0.0;
                for (Index i = 0; i < m_.rows(); ++i) s += m_(i, j);
                out(0, j) = s / static_cast<double>(m_.rows());
            }
            return out;
        }
        MatrixXd norm() const {
            MatrixXd out(1, m_.cols());
            for (Index j = 0; j < m_.cols(); ++j) {
                double s = 0.0;
                for (Index i = 0; i < m_.rows(); ++i) s += m_(i, j) * m_(i, j);
                out(0, j) = std::sqrt(s);
            }
            return out;
        }
    };
    RowwiseOp rowwise() const { return RowwiseOp(*this); }
    ColwiseOp colwise() const { return ColwiseOp(*this); }
};

template <int UpLo> MatrixXd::MatrixXd(const SelfAdjointView<UpLo> &s) {
    take(s.rows(), s.rows());
    for (Index j = 0; j < c_; ++j)
        for (Index i = 0; i < r_; ++i) (*this)(i, j) = s.coeff(i, j);
}

/* Column vector.  Assignment from a 1 x n matrix transposes implicitly, as Eigen does for vectors. */
class VectorXd : public MatrixXd {
  protected:
    VectorXd(ViewTag t, double *p, Index n) : MatrixXd(t, p, n, 1) {}

  public:
    VectorXd() : MatrixXd() {}
    explicit VectorXd(Index n) : MatrixXd(n, 1) {}
    VectorXd(const VectorXd &o) : MatrixXd(o) {}
    VectorXd(const MatrixXd &o) : MatrixXd() { *this = o; }
    VectorXd &operator=(const VectorXd &o) {
        MatrixXd::operator=(o);
        return *this;
    }
    VectorXd &operator=(const MatrixXd &o) {
        if (o.cols() != 1 && o.rows() != 1) throw std::invalid_argument("stub Eigen: vector = matrix");
        take(o.size(), 1);
        for (Index t = 0; t < size(); ++t) p_[t] = o.data()[t];
        return *this;
    }
    Index size() const { return r_; }
    double &operator()(Index i) { return p_[i]; }
    const double &operator()(Index i) const { return p_[i]; }
    double &operator[](Index i) { return p_[i]; }
    const double &operator[](Index i) const { return p_[i]; }
    using MatrixXd::operator();
};

/* Zero-copy view of caller memory (Eigen::Map): what RcppEigen hands the reference for an R matrix. */
template <class T> class Map;
template <> class Map<MatrixXd> : public MatrixXd {
  public:
    Map(double *p, Index r, Index c) : MatrixXd(ViewTag(), p, r, c) {}
    Map(const double *p, Index r, Index c) : MatrixXd(ViewTag(), const_cast<double *>(p), r, c) {}
    Map(const Map &o) : MatrixXd(ViewTag(), o.p_, o.r_, o.c_) {}
};
template <> class Map<VectorXd> : public VectorXd {
  public:
    Map(double *p, Index n) : VectorXd(ViewTag(), p, n) {}
    Map(const double *p, Index n) : VectorXd(ViewTag(), const_cast<double *>(p), n) {}
    Map(const Map &o) : VectorXd(ViewTag(), o.p_, o.r_) {}
};

/* coefficient-wise world: .array(), rowwise() / and *, scalar / array, array * scalar */
class ArrayXXd {
    MatrixXd m_;

  public:
    explicit ArrayXXd(const MatrixXd &m) : m_(m) {}
    const MatrixXd &matrix() const { return m_; }
    Index rows() const { return m_.rows(); }
    Index cols() const { return m_.cols(); }
    double operator()(Index i, Index j) const { return m_(i, j); }
    class RowwiseOp {
        const ArrayXXd &a_;
        template <class F> ArrayXXd apply(const ArrayXXd &row, F f) const {
            if (row.rows() != 1 || row.cols() != a_.cols()) throw std::invalid_argument("stub Eigen: array rowwise op needs a row");
            MatrixXd out(a_.rows(), a_.cols());
            for (Index j = 0; j < a_.cols(); ++j)
                for (Index i = 0; i < a_.rows(); ++i) out(i, j) = f(a_(i, j), row(0, j));
            return ArrayXXd(out);
        }

      public:
        explicit RowwiseOp(const ArrayXXd &a) : a_(a) {}
        ArrayXXd operator/(const ArrayXXd &row) const {
            return apply(row, [](double x, double y) { return x / y; });
        }
        ArrayXXd operator*(const ArrayXXd &row) const {
            return apply(row, [](double x, double y) { return x * y; });
        }
    };
    RowwiseOp rowwise() const { return RowwiseOp(*this); }
    ArrayXXd operator*(double s) const {
        MatrixXd out(m_);
        for (Index t = 0; t < out.size(); ++t) out.data()[t] *= s;
        return ArrayXXd(out);
    }
};
inline ArrayXXd operator/(double s, const ArrayXXd &a) {
    MatrixXd out(a.rows(), a.cols());
    for (Index j = 0; j < a.cols(); ++j)
        for (Index i = 0; i < a.rows(); ++i) out(i, j) = s / a(i, j);
    return ArrayXXd(out);
}
inline ArrayXXd MatrixXd::array() const { return ArrayXXd(*this); }
inline MatrixXd::MatrixXd(const ArrayXXd &a) { copy_from(a.matrix()); }
inline MatrixXd::BlockRef &MatrixXd::BlockRef::operator=(const ArrayXXd &a) { return *this = a.matrix(); }

/* matrix algebra: plain loops, every dot product summed in increasing index order */
inline MatrixXd operator*(const MatrixXd &a, const MatrixXd &b) {
    if (a.cols() != b.rows()) throw std::invalid_argument("stub Eigen: product size mismatch");
    MatrixXd c(a.rows(), b.cols());
    const Index n = a.rows(), kk = a.cols();
    for (Index j = 0; j < b.cols(); ++j) {
        double *cj = c.data() + j * n;
        for (Index t = 0; t < kk; ++t) {
            const double bt = b(t, j);
            const double *at = a.data() + t * n;
            for (Index i = 0; i < n; ++i) cj[i] += at[i] * bt;
        }
    }
    return c;
}
inline MatrixXd operator*(const MatrixXd &a, double s) {
    MatrixXd c(a);
    for (Index t = 0; t < c.size(); ++t) c.data()[t] *= s;
    return c;
}
inline MatrixXd operator-(const MatrixXd &a, const MatrixXd &b) {
    if (a.rows() != b.rows() || a.cols() != b.cols()) throw std::invalid_argument("stub Eigen: difference size mismatch");
    MatrixXd c(a);
    for (Index t = 0; t < c.size(); ++t) c.data()[t] -= b.data()[t];
    return c;
}

/* A = L L' (lower), solve by forward and back substitution */
template <class M> class LLT {
    MatrixXd l_;

  public:
    explicit LLT(const MatrixXd &a) : l_(a.rows(), a.cols()) {
        const Index n = a.rows();
        for (Index j = 0; j < n; ++j) {
            double s = a(j, j);
            for (Index t = 0; t < j; ++t) s -= l_(j, t) * l_(j, t);
            const double ljj = std::sqrt(s);
            l_(j, j) = ljj;
            for (Index i = j + 1; i < n; ++i) {
                double v = a(i, j);
                for (Index t = 0; t < j; ++t) v -= l_(i, t) * l_(j, t);
                l_(i, j) = v / ljj;
            }
        }
    }
    MatrixXd solve(const MatrixXd &b) const {
        const Index n = l_.rows();
        MatrixXd x(b);
        for (Index c = 0; c < x.cols(); ++c) {
            for (Index i = 0; i < n; ++i) {
                double v = x(i, c);
                for (Index t = 0; t < i; ++t) v -= l_(i, t) * x(t, c);
                x(i, c) = v / l_(i, i);
            }
            for (Index i = n - 1; i >= 0; --i) {
                double v = x(i, c);
                for (Index t = i + 1; t < n; ++t) v -= l_(t, i) * x(t, c);
                x(i, c) = v / l_(i, i);
            }
        }
        return x;
    }
};

/* full symmetric eigendecomposition, eigenvalues ascending (reads the lower triangle) */
template <class M> class SelfAdjointEigenSolver {
    VectorXd w_;
    MatrixXd v_;

  public:
    explicit SelfAdjointEigenSolver(const MatrixXd &a) {
        std::vector<double> w;
        stubdetail::jacobi_eigh(a, w, v_);
        w_ = VectorXd(static_cast<Index>(w.size()));
        for (size_t t = 0; t < w.size(); ++t) w_(static_cast<Index>(t)) = w[t];
    }
    const VectorXd &eigenvalues() const { return w_; }
    const MatrixXd &eigenvectors() const { return v_; }
};

namespace stubdetail {
/* Cyclic Jacobi on the symmetric matrix defined by A's lower triangle; ascending eigenvalues,
 * orthonormal eigenvectors in the columns of V.  O(n^3) per sweep: meant for the oracle's sizes. */
inline void jacobi_eigh(const MatrixXd &A, std::vector<double> &w, MatrixXd &V) {
    const Index n = A.rows();
    if (A.cols() != n) throw std::invalid_argument("stub Eigen: eigensolver needs a square matrix");
    MatrixXd S(n, n);
    for (Index j = 0; j < n; ++j)
        for (Index i = j; i < n; ++i) S(i, j) = S(j, i) = A(i, j);
    V = MatrixXd(n, n);
    for (Index i = 0; i < n; ++i) V(i, i) = 1.0;
    double total = 0.0;
    for (Index t = 0; t < n * n; ++t) total += S.data()[t] * S.data()[t];
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0;
        for (Index j = 0; j < n; ++j)
            for (Index i = 0; i < j; ++i) off += S(i, j) * S(i, j);
        if (off <= 1e-36 * total || off == 0.0) break;
        for (Index p = 0; p < n - 1; ++p)
            for (Index q = p + 1; q < n; ++q) {
                const double apq = S(p, q);
                if (apq == 0.0) continue;
                if (std::fabs(apq) < 1e-300) {             /* underflow range: nothing left to rotate */
                    S(p, q) = S(q, p) = 0.0;
                    continue;
                }
                const double app = S(p, p), aqq = S(q, q);
                const double theta = (aqq - app) / (2.0 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                for (Index k = 0; k < n; ++k) {          /* columns p, q of S */
                    const double skp = S(k, p), skq = S(k, q);
                    S(k, p) = c * skp - s * skq;
                    S(k, q) = s * skp + c * skq;
                }
                for (Index k = 0; k < n; ++k) {          /* rows p, q of S */
                    const double spk = S(p, k), sqk = S(q, k);
                    S(p, k) = c * spk - s * sqk;
                    S(q, k) = s * spk + c * sqk;
                }
                S(p, q) = S(q, p) = 0.0;
                for (Index k = 0; k < n; ++k) {
                    const double vkp = V(k, p), vkq = V(k, q);
                    V(k, p) = c * vkp - s * vkq;
                    V(k, q) = s * vkp + c * vkq;
                }
            }
    }
    std::vector<Index> order(static_cast<size_t>(n));
    std::iota(order.begin(), order.end(), Index(0));
    std::sort(order.begin(), order.end(), [&](Index a, Index b) { return S(a, a) < S(b, b); });
    w.resize(static_cast<size_t>(n));
    MatrixXd Vs(n, n);
    for (Index j = 0; j < n; ++j) {
        w[static_cast<size_t>(j)] = S(order[static_cast<size_t>(j)], order[static_cast<size_t>(j)]);
        for (Index i = 0; i < n; ++i) Vs(i, j) = V(i, order[static_cast<size_t>(j)]);
    }
    V = Vs;
}
}  // namespace stubdetail

}  // namespace Eigen
#endif
