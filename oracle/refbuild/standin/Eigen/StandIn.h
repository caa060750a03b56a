// STAND-IN for the Eigen API surface that the PFEM3D hot path uses.  TEST INFRASTRUCTURE ONLY.
//
// Why this exists: the reference (ImperatorS79/PFEM) cannot be built in this image — Eigen, CGAL, gmsh and Lua/sol2
// are absent and there is no network.  To pin the CPU oracle against the reference's OWN source code (not against a
// second restatement), oracle/refbuild compiles the reference's unmodified translation units and templates
// (MatricesBuilder.inl, Element.cpp, Mesh.cpp, MomContEquation*.inl, WCompNewton/{Cont,Mom}Equation.inl,
// PicardAlgo.cpp, Equation.cpp, StatesFromToQ.hpp) where they lie under /root/reference, against this header instead
// of Eigen.  This file is original code written for this repository; it contains nothing from Eigen.  It is an eager
// (no expression templates) dense/sparse mini library that implements only what those files call, with Eigen's
// documented semantics where they matter for parity:
//   * dense matrices are column-major, fixed or dynamic size; products/sums evaluate left to right as written;
//   * the comma initialiser fills row-block by row-block, scalars or whole matrices;
//   * Triplet default = (0, 0, 0.0); SparseMatrix::setFromTriplets sums duplicates in triplet order and yields a
//     compressed column-major matrix with sorted inner indices; explicit zeros are kept; InnerIterator walks a column;
//   * SparseLU is a direct solve: a dense partial-pivot LU here, or a host callback (SciPy SuperLU/COLAMD — the same
//     algorithm family as Eigen::SparseLU) when one is registered through standin_set_direct_solver();
//   * ConjugateGradient is a Jacobi-preconditioned CG with tolerance machine-epsilon and 2n iterations by default (the
//     Eigen 3.3 defaults as recalled in SURVEY.md §8c; only used by the fractional-step path, out of scope).
// Rounding can differ from real Eigen in the last bits (Eigen may vectorise or reassociate small products); parity
// tests therefore compare to 1e-12, not bit-exactly.
#pragma once
#include <algorithm>
#include <array>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <limits>
#include <numeric>
#include <stdexcept>
#include <type_traits>
#include <vector>
#ifdef _OPENMP
#include <omp.h>  // real Eigen pulls omp.h in when OpenMP is on; MomEquation.inl:344 relies on it
#endif

namespace Eigen {

using Index = std::ptrdiff_t;
constexpr int Dynamic = -1;
enum ComputationInfo { Success = 0, NumericalIssue = 1, NoConvergence = 2, InvalidInput = 3 };
enum UpLoType { Lower = 1, Upper = 2 };
inline void setNbThreads(int) {}
inline int nbThreads() { return 1; }
inline void initParallel() {}

template <typename T, int R, int C> class Matrix;
template <typename T, int N> class DiagonalMatrix;

namespace standin {

template <typename T, int R, int C, bool Fixed = (R >= 0 && C >= 0)> struct Storage;
template <typename T, int R, int C> struct Storage<T, R, C, true> {
    std::array<T, static_cast<std::size_t>(R) * C> v{};
    Index rows() const { return R; }
    Index cols() const { return C; }
    void resize(Index r, Index c) { assert(r == R && c == C); (void)r; (void)c; }
};
template <typename T, int R, int C> struct Storage<T, R, C, false> {
    std::vector<T> v;
    Index r_ = (R >= 0 ? R : 0), c_ = (C >= 0 ? C : 0);
    Index rows() const { return r_; }
    Index cols() const { return c_; }
    void resize(Index r, Index c) { r_ = r; c_ = c; v.assign(static_cast<std::size_t>(r * c), T(0)); }
};

constexpr int prodDim(int a, int b) { return (a < 0 || b < 0) ? Dynamic : a; }

template <typename M> class CommaInit {
  public:
    CommaInit(M& m) : m_(m) {}
    template <typename S, typename = std::enable_if_t<std::is_arithmetic<S>::value>> CommaInit& operator,(S s) {
        if (col_ == m_.cols()) { row_ += blockRows_; col_ = 0; blockRows_ = 1; }
        assert(row_ < m_.rows() && col_ < m_.cols());
        m_(row_, col_++) = static_cast<typename M::Scalar>(s);
        return *this;
    }
    template <typename T, int R, int C> CommaInit& operator,(const Matrix<T, R, C>& o) {
        if (col_ == m_.cols()) { row_ += blockRows_; col_ = 0; blockRows_ = o.rows(); }
        assert(row_ + o.rows() <= m_.rows() && col_ + o.cols() <= m_.cols());
        for (Index j = 0; j < o.cols(); ++j)
            for (Index i = 0; i < o.rows(); ++i) m_(row_ + i, col_ + j) = o(i, j);
        col_ += o.cols();
        return *this;
    }
    template <typename S, typename = std::enable_if_t<std::is_arithmetic<S>::value>> CommaInit& first(S s) {
        blockRows_ = 1;
        m_(0, 0) = static_cast<typename M::Scalar>(s);
        col_ = 1;
        return *this;
    }
    template <typename T, int R, int C> CommaInit& first(const Matrix<T, R, C>& o) {
        blockRows_ = o.rows();
        for (Index j = 0; j < o.cols(); ++j)
            for (Index i = 0; i < o.rows(); ++i) m_(i, j) = o(i, j);
        col_ = o.cols();
        return *this;
    }

  private:
    M& m_;
    Index row_ = 0, col_ = 0, blockRows_ = 1;
};

template <typename M> class BlockRef {
  public:
    BlockRef(M& m, Index i, Index j, Index r, Index c) : m_(m), i_(i), j_(j), r_(r), c_(c) {}
    template <typename T, int R, int C> BlockRef& operator=(const Matrix<T, R, C>& o) {
        assert(o.rows() == r_ && o.cols() == c_);
        for (Index b = 0; b < c_; ++b)
            for (Index a = 0; a < r_; ++a) m_(i_ + a, j_ + b) = o(a, b);
        return *this;
    }
    operator Matrix<typename M::Scalar, Dynamic, Dynamic>() const {
        Matrix<typename M::Scalar, Dynamic, Dynamic> out(r_, c_);
        for (Index b = 0; b < c_; ++b)
            for (Index a = 0; a < r_; ++a) out(a, b) = m_(i_ + a, j_ + b);
        return out;
    }

  private:
    M& m_;
    Index i_, j_, r_, c_;
};

template <typename T, int N> struct QrSolver {  // stands in for colPivHouseholderQr().solve(): Gauss, full row pivoting
    Matrix<T, N, N> A;
    Matrix<T, N, 1> solve(Matrix<T, N, 1> b) const {
        Matrix<T, N, N> a = A;
        const Index n = a.rows();
        for (Index k = 0; k < n; ++k) {
            Index p = k;
            for (Index i = k + 1; i < n; ++i)
                if (std::abs(a(i, k)) > std::abs(a(p, k))) p = i;
            if (p != k) {
                for (Index j = 0; j < n; ++j) std::swap(a(k, j), a(p, j));
                std::swap(b[k], b[p]);
            }
            for (Index i = k + 1; i < n; ++i) {
                const T f = a(i, k) / a(k, k);
                for (Index j = k; j < n; ++j) a(i, j) -= f * a(k, j);
                b[i] -= f * b[k];
            }
        }
        for (Index k = n - 1; k >= 0; --k) {
            T s = b[k];
            for (Index j = k + 1; j < n; ++j) s -= a(k, j) * b[j];
            b[k] = s / a(k, k);
        }
        return b;
    }
};

}  // namespace standin

template <typename T, int R, int C> class Matrix {
  public:
    using Scalar = T;
    static constexpr int RowsAtCompileTime = R, ColsAtCompileTime = C;

    Matrix() = default;
    template <int CC = C, typename = std::enable_if_t<CC == 1 || R == 1>> explicit Matrix(Index n) {
        if (C == 1) s_.resize(n, 1); else s_.resize(1, n);
    }
    Matrix(Index r, Index c) { s_.resize(r, c); }
    template <int R2, int C2> Matrix(const Matrix<T, R2, C2>& o) { assignFrom(o); }
    template <int R2, int C2> Matrix& operator=(const Matrix<T, R2, C2>& o) { assignFrom(o); return *this; }

    Index rows() const { return s_.rows(); }
    Index cols() const { return s_.cols(); }
    Index size() const { return rows() * cols(); }
    void resize(Index n) { if (C == 1) s_.resize(n, 1); else s_.resize(1, n); }
    void resize(Index r, Index c) { s_.resize(r, c); }
    Matrix& setZero() { std::fill(s_.v.begin(), s_.v.end(), T(0)); return *this; }
    Matrix& setZero(Index n) { resize(n); return setZero(); }
    Matrix& setOnes() { std::fill(s_.v.begin(), s_.v.end(), T(1)); return *this; }
    Matrix& setConstant(T x) { std::fill(s_.v.begin(), s_.v.end(), x); return *this; }
    static Matrix Zero() { Matrix m; m.setZero(); return m; }
    static Matrix Zero(Index n) { Matrix m(n); m.setZero(); return m; }
    static Matrix Zero(Index r, Index c) { Matrix m(r, c); m.setZero(); return m; }
    static Matrix Map(const T* p, Index n) {
        Matrix m;
        m.resize(R >= 0 ? R : n, C >= 0 ? C : 1);
        for (Index i = 0; i < n; ++i) m(i) = p[i];
        return m;
    }
    T* data() { return s_.v.data(); }
    const T* data() const { return s_.v.data(); }

    T& operator()(Index i, Index j) { assert(i >= 0 && i < rows() && j >= 0 && j < cols()); return s_.v[static_cast<std::size_t>(i + j * rows())]; }
    const T& operator()(Index i, Index j) const { assert(i >= 0 && i < rows() && j >= 0 && j < cols()); return s_.v[static_cast<std::size_t>(i + j * rows())]; }
    T& operator()(Index i) { assert(i >= 0 && i < size()); return s_.v[static_cast<std::size_t>(i)]; }
    const T& operator()(Index i) const { assert(i >= 0 && i < size()); return s_.v[static_cast<std::size_t>(i)]; }
    T& operator[](Index i) { return (*this)(i); }
    const T& operator[](Index i) const { return (*this)(i); }
    T& coeffRef(Index i, Index j) { return (*this)(i, j); }
    T& coeffRef(Index i) { return (*this)(i); }
    T value() const { assert(size() == 1); return s_.v[0]; }
    template <int RR = R, int CC = C, typename = std::enable_if_t<RR == 1 && CC == 1>> operator T() const { return s_.v[0]; }

    Matrix<T, C, R> transpose() const {
        Matrix<T, C, R> t(cols(), rows());
        for (Index j = 0; j < cols(); ++j)
            for (Index i = 0; i < rows(); ++i) t(j, i) = (*this)(i, j);
        return t;
    }
    Matrix<T, R, 1> diagonal() const {
        Matrix<T, R, 1> d(rows(), 1);
        for (Index i = 0; i < rows(); ++i) d[i] = (*this)(i, i);
        return d;
    }
    T squaredNorm() const { T s = 0; for (const T& x : s_.v) s += x * x; return s; }
    T norm() const { return std::sqrt(squaredNorm()); }
    T sum() const { T s = 0; for (const T& x : s_.v) s += x; return s; }
    template <int R2, int C2> T dot(const Matrix<T, R2, C2>& o) const {
        assert(size() == o.size());
        T s = 0;
        for (Index i = 0; i < size(); ++i) s += (*this)(i) * o(i);
        return s;
    }
    standin::BlockRef<Matrix> block(Index i, Index j, Index r, Index c) { return standin::BlockRef<Matrix>(*this, i, j, r, c); }
    standin::QrSolver<T, R> colPivHouseholderQr() const { return standin::QrSolver<T, R>{*this}; }

    template <typename S, typename = std::enable_if_t<std::is_arithmetic<S>::value>> standin::CommaInit<Matrix> operator<<(S s) {
        standin::CommaInit<Matrix> ci(*this);
        ci.first(s);
        return ci;
    }
    template <int R2, int C2> standin::CommaInit<Matrix> operator<<(const Matrix<T, R2, C2>& o) {
        standin::CommaInit<Matrix> ci(*this);
        ci.first(o);
        return ci;
    }

    template <int R2, int C2> Matrix& operator+=(const Matrix<T, R2, C2>& o) {
        assert(rows() == o.rows() && cols() == o.cols());
        for (Index i = 0; i < size(); ++i) (*this)(i) += o(i);
        return *this;
    }
    template <int R2, int C2> Matrix& operator-=(const Matrix<T, R2, C2>& o) {
        assert(rows() == o.rows() && cols() == o.cols());
        for (Index i = 0; i < size(); ++i) (*this)(i) -= o(i);
        return *this;
    }
    Matrix& operator*=(T s) { for (T& x : s_.v) x *= s; return *this; }
    Matrix& operator/=(T s) { for (T& x : s_.v) x /= s; return *this; }
    Matrix operator-() const { Matrix m = *this; for (T& x : m.s_.v) x = -x; return m; }

  private:
    template <int R2, int C2> void assignFrom(const Matrix<T, R2, C2>& o) {
        static_assert((R < 0 || R2 < 0 || R == R2) && (C < 0 || C2 < 0 || C == C2), "stand-in Eigen: size mismatch");
        s_.resize(o.rows(), o.cols());
        for (Index j = 0; j < o.cols(); ++j)
            for (Index i = 0; i < o.rows(); ++i) (*this)(i, j) = o(i, j);
    }
    standin::Storage<T, R, C> s_;
};

using VectorXd = Matrix<double, Dynamic, 1>;
using MatrixXd = Matrix<double, Dynamic, Dynamic>;
using Vector2d = Matrix<double, 2, 1>;
using Vector3d = Matrix<double, 3, 1>;
using Matrix2d = Matrix<double, 2, 2>;
using Matrix3d = Matrix<double, 3, 3>;

template <typename T, int R, int K, int K2, int C>
Matrix<T, standin::prodDim(R, R), standin::prodDim(C, C)> operator*(const Matrix<T, R, K>& a, const Matrix<T, K2, C>& b) {
    static_assert(K < 0 || K2 < 0 || K == K2, "stand-in Eigen: inner dimensions differ");
    assert(a.cols() == b.rows());
    Matrix<T, R, C> out(a.rows(), b.cols());
    for (Index j = 0; j < b.cols(); ++j)
        for (Index i = 0; i < a.rows(); ++i) {
            T s = 0;
            for (Index k = 0; k < a.cols(); ++k) s += a(i, k) * b(k, j);
            out(i, j) = s;
        }
    return out;
}
template <typename T, int R, int C, typename S, typename = std::enable_if_t<std::is_arithmetic<S>::value>>
Matrix<T, R, C> operator*(S s, const Matrix<T, R, C>& a) {
    Matrix<T, R, C> out = a;
    out *= static_cast<T>(s);
    return out;
}
template <typename T, int R, int C, typename S, typename = std::enable_if_t<std::is_arithmetic<S>::value>>
Matrix<T, R, C> operator*(const Matrix<T, R, C>& a, S s) {
    Matrix<T, R, C> out = a;
    out *= static_cast<T>(s);
    return out;
}
template <typename T, int R, int C, typename S, typename = std::enable_if_t<std::is_arithmetic<S>::value>>
Matrix<T, R, C> operator/(const Matrix<T, R, C>& a, S s) {
    Matrix<T, R, C> out = a;
    out /= static_cast<T>(s);
    return out;
}
template <typename T, int R, int C, int R2, int C2> Matrix<T, R, C> operator+(const Matrix<T, R, C>& a, const Matrix<T, R2, C2>& b) {
    Matrix<T, R, C> out = a;
    out += b;
    return out;
}
template <typename T, int R, int C, int R2, int C2> Matrix<T, R, C> operator-(const Matrix<T, R, C>& a, const Matrix<T, R2, C2>& b) {
    Matrix<T, R, C> out = a;
    out -= b;
    return out;
}

template <typename M> class Map : public M {
  public:
    using T = typename M::Scalar;
    Map(const T* p, Index n) {
        M::resize(M::RowsAtCompileTime >= 0 ? M::RowsAtCompileTime : n, M::ColsAtCompileTime >= 0 ? M::ColsAtCompileTime : 1);
        assert(M::size() == n);
        for (Index i = 0; i < n; ++i) (*this)(i) = p[i];
    }
    Map(const T* p, Index r, Index c) {
        M::resize(r, c);
        for (Index i = 0; i < r * c; ++i) (*this)(i) = p[i];
    }
};

template <typename T, int N> class DiagonalMatrix {
  public:
    DiagonalMatrix() = default;
    explicit DiagonalMatrix(Index n) { d_.resize(n); }
    Matrix<T, N, 1>& diagonal() { return d_; }
    const Matrix<T, N, 1>& diagonal() const { return d_; }
    void setZero() { d_.setZero(); }
    void resize(Index n) { d_.resize(n); }
    Index rows() const { return d_.rows(); }
    Index cols() const { return d_.rows(); }

  private:
    Matrix<T, N, 1> d_;
};
template <typename T, int N, typename S, typename = std::enable_if_t<std::is_arithmetic<S>::value>>
DiagonalMatrix<T, N> operator*(S s, const DiagonalMatrix<T, N>& D) {
    DiagonalMatrix<T, N> out = D;
    out.diagonal() *= static_cast<T>(s);
    return out;
}
template <typename T, int N, typename S, typename = std::enable_if_t<std::is_arithmetic<S>::value>>
DiagonalMatrix<T, N> operator*(const DiagonalMatrix<T, N>& D, S s) { return s * D; }
template <typename T, int N, int R, int C> Matrix<T, R, C> operator*(const DiagonalMatrix<T, N>& D, const Matrix<T, R, C>& x) {
    assert(D.rows() == x.rows());
    Matrix<T, R, C> out(x.rows(), x.cols());
    for (Index j = 0; j < x.cols(); ++j)
        for (Index i = 0; i < x.rows(); ++i) out(i, j) = D.diagonal()[i] * x(i, j);
    return out;
}

// ------------------------------------------------------------------------------------------------------------------
// sparse
template <typename T, typename I = int> class Triplet {
  public:
    Triplet() : r_(0), c_(0), v_(0) {}
    Triplet(const I& i, const I& j, const T& v = T(0)) : r_(i), c_(j), v_(v) {}
    const I& row() const { return r_; }
    const I& col() const { return c_; }
    const T& value() const { return v_; }

  private:
    I r_, c_;
    T v_;
};

template <typename T, int Options = 0, typename I = int> class SparseMatrix {
  public:
    using Scalar = T;
    using StorageIndex = I;
    SparseMatrix() = default;
    SparseMatrix(Index r, Index c) { resize(r, c); }
    void resize(Index r, Index c) {
        rows_ = r; cols_ = c;
        outer_.assign(static_cast<std::size_t>(c + 1), 0);
        inner_.clear(); val_.clear();
    }
    Index rows() const { return rows_; }
    Index cols() const { return cols_; }
    Index nonZeros() const { return static_cast<Index>(val_.size()); }
    Index outerSize() const { return cols_; }
    void setZero() { resize(rows_, cols_); }
    void makeCompressed() {}
    bool isCompressed() const { return true; }
    const I* outerIndexPtr() const { return outer_.data(); }
    const I* innerIndexPtr() const { return inner_.data(); }
    const T* valuePtr() const { return val_.data(); }
    T* valuePtr() { return val_.data(); }

    // duplicates are summed in the order the triplets appear; columns end up with sorted row indices
    template <typename It> void setFromTriplets(It begin, It end) {
        std::vector<I> rowCount(static_cast<std::size_t>(rows_) + 1, 0);
        std::size_t n = 0;
        for (It it = begin; it != end; ++it, ++n) {
            assert(it->row() >= 0 && it->row() < rows_ && it->col() >= 0 && it->col() < cols_);
            ++rowCount[static_cast<std::size_t>(it->row()) + 1];
        }
        // pass 1: bucket by row (stable), pass 2: bucket by column (stable) -> column-major, rows sorted, input order kept
        for (std::size_t i = 0; i < static_cast<std::size_t>(rows_); ++i) rowCount[i + 1] += rowCount[i];
        std::vector<I> tc(n); std::vector<T> tv(n); std::vector<I> tr(n);
        {
            std::vector<I> pos(rowCount.begin(), rowCount.end() - 1);
            for (It it = begin; it != end; ++it) {
                const std::size_t p = static_cast<std::size_t>(pos[static_cast<std::size_t>(it->row())]++);
                tr[p] = static_cast<I>(it->row()); tc[p] = static_cast<I>(it->col()); tv[p] = it->value();
            }
        }
        std::vector<I> colCount(static_cast<std::size_t>(cols_) + 1, 0);
        for (std::size_t k = 0; k < n; ++k) ++colCount[static_cast<std::size_t>(tc[k]) + 1];
        for (std::size_t j = 0; j < static_cast<std::size_t>(cols_); ++j) colCount[j + 1] += colCount[j];
        std::vector<I> r2(n); std::vector<T> v2(n);
        {
            std::vector<I> pos(colCount.begin(), colCount.end() - 1);
            for (std::size_t k = 0; k < n; ++k) {
                const std::size_t p = static_cast<std::size_t>(pos[static_cast<std::size_t>(tc[k])]++);
                r2[p] = tr[k]; v2[p] = tv[k];
            }
        }
        outer_.assign(static_cast<std::size_t>(cols_) + 1, 0);
        inner_.clear(); val_.clear();
        for (Index j = 0; j < cols_; ++j) {
            for (I k = colCount[static_cast<std::size_t>(j)]; k < colCount[static_cast<std::size_t>(j) + 1]; ++k) {
                if (!inner_.empty() && static_cast<Index>(inner_.size()) > outer_[static_cast<std::size_t>(j)] && inner_.back() == r2[static_cast<std::size_t>(k)])
                    val_.back() += v2[static_cast<std::size_t>(k)];
                else { inner_.push_back(r2[static_cast<std::size_t>(k)]); val_.push_back(v2[static_cast<std::size_t>(k)]); }
            }
            outer_[static_cast<std::size_t>(j) + 1] = static_cast<I>(inner_.size());
        }
    }

    class InnerIterator {
      public:
        InnerIterator(SparseMatrix& m, Index outer)
            : m_(&m), outer_(outer), k_(m.outer_[static_cast<std::size_t>(outer)]), end_(m.outer_[static_cast<std::size_t>(outer) + 1]) {}
        InnerIterator(const SparseMatrix& m, Index outer) : InnerIterator(const_cast<SparseMatrix&>(m), outer) {}
        InnerIterator& operator++() { ++k_; return *this; }
        operator bool() const { return k_ < end_; }
        Index row() const { return m_->inner_[static_cast<std::size_t>(k_)]; }
        Index col() const { return outer_; }
        Index index() const { return row(); }
        Index outer() const { return outer_; }
        const T& value() const { return m_->val_[static_cast<std::size_t>(k_)]; }
        T& valueRef() { return m_->val_[static_cast<std::size_t>(k_)]; }

      private:
        SparseMatrix* m_;
        Index outer_;
        I k_, end_;
    };

    T coeff(Index r, Index c) const {
        for (I k = outer_[static_cast<std::size_t>(c)]; k < outer_[static_cast<std::size_t>(c) + 1]; ++k)
            if (inner_[static_cast<std::size_t>(k)] == r) return val_[static_cast<std::size_t>(k)];
        return T(0);
    }

  private:
    Index rows_ = 0, cols_ = 0;
    std::vector<I> outer_{0}, inner_;
    std::vector<T> val_;
};

template <typename T, int O, typename I, int R> Matrix<T, R, 1> operator*(const SparseMatrix<T, O, I>& A, const Matrix<T, R, 1>& x) {
    assert(A.cols() == x.rows());
    Matrix<T, R, 1> y(A.rows());
    y.setZero();
    const I* op = A.outerIndexPtr();
    const I* ip = A.innerIndexPtr();
    const T* vp = A.valuePtr();
    for (Index j = 0; j < A.cols(); ++j)
        for (I k = op[j]; k < op[j + 1]; ++k) y[ip[k]] += vp[k] * x[j];
    return y;
}

template <typename I> class COLAMDOrdering {};
template <typename I> class AMDOrdering {};
template <typename I> class NaturalOrdering {};
template <typename T> class DiagonalPreconditioner {};
template <typename T> class IdentityPreconditioner {};

namespace standin {
// optional host hook: solve A x = b for a CSC matrix; return 0 on success.
using DirectSolverFn = int (*)(int n, const int* colPtr, const int* rowIdx, const double* val, const double* b, double* x);
inline DirectSolverFn& directSolverHook() { static DirectSolverFn fn = nullptr; return fn; }
inline long& directSolveCount() { static long n = 0; return n; }
struct CgRecord { long n, iterations; double error; int info; };
inline std::vector<CgRecord>& cgLog() { static std::vector<CgRecord> log; return log; }
}  // namespace standin

template <typename MatrixType, typename Ordering = COLAMDOrdering<int>> class SparseLU {
  public:
    using T = typename MatrixType::Scalar;
    void analyzePattern(const MatrixType&) {}
    void factorize(const MatrixType& A) { A_ = &A; info_ = (A.rows() == A.cols() && A.rows() > 0) ? Success : InvalidInput; }
    void compute(const MatrixType& A) { factorize(A); }
    ComputationInfo info() const { return info_; }
    Matrix<T, Dynamic, 1> solve(const Matrix<T, Dynamic, 1>& b) const {
        const Index n = A_->rows();
        Matrix<T, Dynamic, 1> x(n);
        ++standin::directSolveCount();
        if (standin::directSolverHook()) {
            const int rc = standin::directSolverHook()(static_cast<int>(n), A_->outerIndexPtr(), A_->innerIndexPtr(), A_->valuePtr(), b.data(), x.data());
            if (rc != 0) info_ = NumericalIssue;
            return x;
        }
        std::vector<T> a(static_cast<std::size_t>(n * n), T(0));  // row-major dense copy, partial-pivot LU
        for (Index j = 0; j < n; ++j)
            for (typename MatrixType::InnerIterator it(*A_, j); it; ++it) a[static_cast<std::size_t>(it.row() * n + j)] += it.value();
        std::vector<T> r(b.data(), b.data() + n);
        for (Index k = 0; k < n; ++k) {
            Index p = k;
            for (Index i = k + 1; i < n; ++i)
                if (std::abs(a[static_cast<std::size_t>(i * n + k)]) > std::abs(a[static_cast<std::size_t>(p * n + k)])) p = i;
            if (a[static_cast<std::size_t>(p * n + k)] == T(0)) { info_ = NumericalIssue; return x; }
            if (p != k) {
                std::swap_ranges(a.begin() + k * n, a.begin() + (k + 1) * n, a.begin() + p * n);
                std::swap(r[static_cast<std::size_t>(k)], r[static_cast<std::size_t>(p)]);
            }
            const T piv = a[static_cast<std::size_t>(k * n + k)];
            for (Index i = k + 1; i < n; ++i) {
                const T f = a[static_cast<std::size_t>(i * n + k)] / piv;
                if (f == T(0)) continue;
                T* ri = &a[static_cast<std::size_t>(i * n)];
                const T* rk = &a[static_cast<std::size_t>(k * n)];
                for (Index j = k + 1; j < n; ++j) ri[j] -= f * rk[j];
                r[static_cast<std::size_t>(i)] -= f * r[static_cast<std::size_t>(k)];
            }
        }
        for (Index k = n - 1; k >= 0; --k) {
            T s = r[static_cast<std::size_t>(k)];
            for (Index j = k + 1; j < n; ++j) s -= a[static_cast<std::size_t>(k * n + j)] * x[j];
            x[k] = s / a[static_cast<std::size_t>(k * n + k)];
        }
        return x;
    }

  private:
    const MatrixType* A_ = nullptr;
    mutable ComputationInfo info_ = InvalidInput;
};

template <typename MatrixType, int UpLo = Lower, typename Preconditioner = DiagonalPreconditioner<typename MatrixType::Scalar>>
class ConjugateGradient {
  public:
    using T = typename MatrixType::Scalar;
    using Vec = Matrix<T, Dynamic, 1>;
    ConjugateGradient& compute(const MatrixType& A) {
        A_ = &A;
        const Index n = A.cols();
        invDiag_.resize(n);
        for (Index j = 0; j < n; ++j) {
            const T d = A.coeff(j, j);
            invDiag_[j] = (d != T(0)) ? T(1) / d : T(1);
        }
        info_ = Success;
        return *this;
    }
    ConjugateGradient& setTolerance(T t) { tol_ = t; return *this; }
    ConjugateGradient& setMaxIterations(Index m) { maxIt_ = m; return *this; }
    ComputationInfo info() const { return info_; }
    Index iterations() const { return iters_; }
    T error() const { return err_; }
    Vec solve(const Vec& b) const { Vec x0(b.rows()); x0.setZero(); return solveWithGuess(b, x0); }
    Vec solveWithGuess(const Vec& b, const Vec& x0) const {
        const Index n = b.rows();
        const Index maxIt = maxIt_ >= 0 ? maxIt_ : 2 * n;
        Vec x = x0;
        Vec r = b - (*A_) * x;
        const T rhsNorm2 = b.squaredNorm();
        iters_ = 0; err_ = 0; info_ = Success;
        if (rhsNorm2 == T(0)) { x.setZero(); return x; }
        const T threshold = std::max(tol_ * tol_ * rhsNorm2, std::numeric_limits<T>::min());
        T resNorm2 = r.squaredNorm();
        if (resNorm2 < threshold) { err_ = std::sqrt(resNorm2 / rhsNorm2); return x; }
        Vec z(n), p(n), tmp(n);
        for (Index i = 0; i < n; ++i) p[i] = invDiag_[i] * r[i];
        T absNew = r.dot(p);
        Index i = 0;
        while (i < maxIt) {
            tmp = (*A_) * p;
            const T alpha = absNew / p.dot(tmp);
            for (Index k = 0; k < n; ++k) { x[k] += alpha * p[k]; r[k] -= alpha * tmp[k]; }
            resNorm2 = r.squaredNorm();
            if (resNorm2 < threshold) break;
            for (Index k = 0; k < n; ++k) z[k] = invDiag_[k] * r[k];
            const T absOld = absNew;
            absNew = r.dot(z);
            const T beta = absNew / absOld;
            for (Index k = 0; k < n; ++k) p[k] = z[k] + beta * p[k];
            ++i;
        }
        iters_ = i;
        err_ = std::sqrt(resNorm2 / rhsNorm2);
        info_ = (resNorm2 < threshold) ? Success : NoConvergence;
        standin::cgLog().push_back({static_cast<long>(n), static_cast<long>(iters_), static_cast<double>(err_), static_cast<int>(info_)});
        return x;
    }

  private:
    const MatrixType* A_ = nullptr;
    Vec invDiag_;
    T tol_ = std::numeric_limits<T>::epsilon();
    Index maxIt_ = -1;
    mutable Index iters_ = 0;
    mutable T err_ = 0;
    mutable ComputationInfo info_ = Success;
};

}  // namespace Eigen
