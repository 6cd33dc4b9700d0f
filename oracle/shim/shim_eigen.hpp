// shim_eigen.hpp — TEST INFRASTRUCTURE.  A minimal stand-in for the subset of Eigen 3 that the reference's
// linear propagator and PCCBlackmoreSVC touch, so that those reference SOURCES (compiled where they lie under
// /root/reference) can run here without Eigen.  Eager evaluation, coefficient (i,j) of a product is the sum
// over k in increasing order.  Every product on this path has inner dimension 2, where the order of the two
// terms cannot change the rounded sum, so the values equal what any Eigen build without FMA produces
// (the reference builds for baseline x86-64: CMakeLists.txt:6-8).  The 2x2 inverse follows Eigen's
// compute_inverse_size2_helper (adjugate times 1/det).  NOT a general Eigen replacement.
//
// Round 2 additions, for the unicycle propagator, the belief-space distance, the goal and the chi-squared /
// CDF-grid checkers (Eigen 3.4.0 is what the reference's Dockerfile installs; its algorithms are restated from
// the published sources, NOT bit-for-bit: these pins hold values to 1e-9 relative):
//   * 4x4 inverse: Eigen's generic compute_inverse_size4 (cofactor_4x4 / general_det3_helper, then division by
//     the determinant col(0).row(0)) - on x86-64 Eigen actually takes an SSE2 2x2-block variant whose rounding
//     differs in the last ulps;
//   * general n x n inverse: partial-pivot LU (Eigen's PartialPivLU for dynamic sizes);
//   * MatrixBase::exp(): unsupported/MatrixFunctions matrix_exp_compute for double: Pade (3,5,7,9,13) by l1 norm
//     with scaling and squaring, solved with partial-pivot LU;
//   * SelfAdjointEigenSolver<MatrixXd>: lower triangle only (as Eigen), cyclic Jacobi here instead of Eigen's
//     tridiagonal QL - eigenvalues agree to rounding, operatorSqrt = V sqrt(D) V^T;
//   * EigenSolver<Matrix2d>: RealSchur's 2x2 path (scale by the max coefficient, splitOffTwoRows with a Givens
//     rotation), eigenvectors by back substitution on T times U, normalised columns.
#pragma once
// Eigen/Core includes <emmintrin.h> on x86-64 (SSE2 vectorisation is on by default), which reaches <stdlib.h>
// through <mm_malloc.h>; libstdc++'s <stdlib.h> wrapper does `using std::abs;`.  That is what makes the reference's
// UNQUALIFIED `abs(t - current_time) < 1E-9` in *PVC::satisfiesConstraints the floating-point overload — with
// <cmath>/<cstdlib> alone it would be int abs(int) and match every constraint time within +-1 s.  The stand-in
// must reproduce that include, or the reference sources compiled here would behave differently.
#include <emmintrin.h>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <algorithm>
#include <limits>
#include <ostream>
#include <vector>

#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW

namespace Eigen {

enum { Dynamic = -1, DontAlign = 0x2 };

template <class T, int R, int C, int Opt = 0>
class Matrix {
public:
    Matrix() : r_(R == Dynamic ? 0 : R), c_(C == Dynamic ? 0 : C), d_((size_t)r_ * c_, T(0)) {}
    Matrix(int r, int c) : r_(r), c_(c), d_((size_t)r * c, T(0)) {}
    explicit Matrix(int n) : r_(n), c_(1), d_((size_t)n, T(0)) {}   // VectorXd mu(2)
    // Vector2d{a, b}
    Matrix(T a, T b) : r_(R == Dynamic ? 2 : R), c_(C == Dynamic ? 1 : C), d_((size_t)r_ * c_, T(0)) { d_[0] = a; d_[1] = b; }
    template <int R2, int C2, int O2>
    Matrix(const Matrix<T, R2, C2, O2>& o) : r_(o.rows()), c_(o.cols()), d_((size_t)r_ * c_) {
        assert((R == Dynamic || R == r_) && (C == Dynamic || C == c_));
        for (int i = 0; i < r_; ++i) for (int j = 0; j < c_; ++j) (*this)(i, j) = o(i, j);
    }
    template <int R2, int C2, int O2>
    Matrix& operator=(const Matrix<T, R2, C2, O2>& o) {
        r_ = o.rows(); c_ = o.cols(); d_.assign((size_t)r_ * c_, T(0));
        for (int i = 0; i < r_; ++i) for (int j = 0; j < c_; ++j) (*this)(i, j) = o(i, j);
        return *this;
    }
    int rows() const { return r_; }
    int cols() const { return c_; }
    void resize(int r, int c) {
        if (r != r_ || c != c_) { r_ = r; c_ = c; d_.assign((size_t)r * c, T(0)); }
        if (d_.empty()) std::vector<T>().swap(d_);   // resize(0,0) releases the storage (RealVectorBeliefSpace::freeState relies on it)
    }
    T& operator()(int i, int j) { return d_[(size_t)i * c_ + j]; }
    const T& operator()(int i, int j) const { return d_[(size_t)i * c_ + j]; }
    T& operator[](int i) { return d_[i]; }
    const T& operator[](int i) const { return d_[i]; }
    T value() const { assert(r_ == 1 && c_ == 1); return d_[0]; }

    static Matrix Zero() { return Matrix(); }
    static Matrix Zero(int r, int c) { return Matrix(r, c); }
    static Matrix Zero(int n) { return Matrix(n); }
    static Matrix Identity() { Matrix m; for (int i = 0; i < m.r_ && i < m.c_; ++i) m(i, i) = T(1); return m; }
    static Matrix Identity(int r, int c) { Matrix m(r, c); for (int i = 0; i < r && i < c; ++i) m(i, i) = T(1); return m; }

    // comma initialiser: M << a, b, c, d;
    struct Comma {
        Matrix& m; int k;
        Comma& operator,(T v) { m.d_[k++] = v; return *this; }
    };
    Comma operator<<(T v) { d_[0] = v; return Comma{*this, 1}; }

    Matrix<T, Dynamic, Dynamic> transpose() const {
        Matrix<T, Dynamic, Dynamic> t(c_, r_);
        for (int i = 0; i < r_; ++i) for (int j = 0; j < c_; ++j) t(j, i) = (*this)(i, j);
        return t;
    }
    Matrix<T, Dynamic, Dynamic> row(int i) const {
        Matrix<T, Dynamic, Dynamic> t(1, c_);
        for (int j = 0; j < c_; ++j) t(0, j) = (*this)(i, j);
        return t;
    }
    // Eigen::internal::compute_inverse_size2_helper: invdet = 1/det; [d -b; -c a] * invdet
    // size 4: generic compute_inverse_size4 (cofactors, then /= determinant); other sizes: partial-pivot LU
    Matrix<T, Dynamic, Dynamic> inverse() const {
        assert(r_ == c_);
        Matrix<T, Dynamic, Dynamic> m(r_, c_);
        if (r_ == 1) { m(0, 0) = T(1) / (*this)(0, 0); return m; }
        if (r_ == 2) {
            const T det = (*this)(0, 0) * (*this)(1, 1) - (*this)(1, 0) * (*this)(0, 1);
            const T invdet = T(1) / det;
            m(0, 0) = (*this)(1, 1) * invdet;
            m(1, 0) = -(*this)(1, 0) * invdet;
            m(0, 1) = -(*this)(0, 1) * invdet;
            m(1, 1) = (*this)(0, 0) * invdet;
            return m;
        }
        if (r_ == 4) {
            const Matrix& a = *this;
            auto det3 = [&](int i1, int i2, int i3, int j1, int j2, int j3) { return a(i1, j1) * (a(i2, j2) * a(i3, j3) - a(i2, j3) * a(i3, j2)); };
            auto cof = [&](int i, int j) {
                const int i1 = (i + 1) % 4, i2 = (i + 2) % 4, i3 = (i + 3) % 4, j1 = (j + 1) % 4, j2 = (j + 2) % 4, j3 = (j + 3) % 4;
                return det3(i1, i2, i3, j1, j2, j3) + det3(i2, i3, i1, j1, j2, j3) + det3(i3, i1, i2, j1, j2, j3);
            };
            for (int i = 0; i < 4; ++i)
                for (int j = 0; j < 4; ++j) m(j, i) = ((i + j) & 1) ? -cof(i, j) : cof(i, j);
            T det = T(0);
            for (int i = 0; i < 4; ++i) det = (i == 0) ? a(0, 0) * m(0, 0) : det + a(i, 0) * m(0, i);
            for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) m(i, j) = m(i, j) / det;
            return m;
        }
        return lu_solve(*this, Matrix<T, Dynamic, Dynamic>::Identity(r_, c_));
    }
    // partial-pivot LU solve A X = B (Eigen: PartialPivLU::solve)
    template <int R2, int C2, int O2>
    static Matrix<T, Dynamic, Dynamic> lu_solve(const Matrix& A, const Matrix<T, R2, C2, O2>& B) {
        const int n = A.rows();
        Matrix<T, Dynamic, Dynamic> lu(A), x(B);
        for (int k = 0; k < n; ++k) {
            int piv = k;
            for (int i = k + 1; i < n; ++i) if (std::fabs(lu(i, k)) > std::fabs(lu(piv, k))) piv = i;
            if (piv != k) { for (int j = 0; j < n; ++j) std::swap(lu(k, j), lu(piv, j)); for (int j = 0; j < x.cols(); ++j) std::swap(x(k, j), x(piv, j)); }
            for (int i = k + 1; i < n; ++i) {
                lu(i, k) = lu(i, k) / lu(k, k);
                for (int j = k + 1; j < n; ++j) lu(i, j) = lu(i, j) - lu(i, k) * lu(k, j);
            }
        }
        for (int c = 0; c < x.cols(); ++c) {
            for (int i = 0; i < n; ++i) for (int j = 0; j < i; ++j) x(i, c) = x(i, c) - lu(i, j) * x(j, c);
            for (int i = n - 1; i >= 0; --i) { for (int j = i + 1; j < n; ++j) x(i, c) = x(i, c) - lu(i, j) * x(j, c); x(i, c) = x(i, c) / lu(i, i); }
        }
        return x;
    }
    // unsupported/Eigen/MatrixFunctions: matrix exponential (Higham 2005 Pade + scaling and squaring, double thresholds)
    Matrix<T, Dynamic, Dynamic> exp() const;

    T squaredNorm() const { T a = T(0); for (size_t i = 0; i < d_.size(); ++i) a = (i == 0) ? d_[0] * d_[0] : a + d_[i] * d_[i]; return a; }
    T norm() const { return std::sqrt(squaredNorm()); }
    T trace() const { T a = T(0); for (int i = 0; i < r_ && i < c_; ++i) a = (i == 0) ? (*this)(0, 0) : a + (*this)(i, i); return a; }
    T maxCoeff() const { T a = d_[0]; for (size_t i = 1; i < d_.size(); ++i) if (d_[i] > a) a = d_[i]; return a; }
    T minCoeff() const { T a = d_[0]; for (size_t i = 1; i < d_.size(); ++i) if (d_[i] < a) a = d_[i]; return a; }
    T sum() const { T a = T(0); for (size_t i = 0; i < d_.size(); ++i) a = (i == 0) ? d_[0] : a + d_[i]; return a; }
    Matrix<T, Dynamic, Dynamic> operator-() const { Matrix<T, Dynamic, Dynamic> m(r_, c_); for (int i = 0; i < r_; ++i) for (int j = 0; j < c_; ++j) m(i, j) = -(*this)(i, j); return m; }
    Matrix<T, Dynamic, Dynamic> col(int j) const { Matrix<T, Dynamic, Dynamic> t(r_, 1); for (int i = 0; i < r_; ++i) t(i, 0) = (*this)(i, j); return t; }
    Matrix& operator*=(T s) { for (auto& v : d_) v = v * s; return *this; }
    const Matrix& real() const { return *this; }

    // diagonal() << vector  (CDFGridPVC.cpp:192)
    struct DiagProxy {
        Matrix* m;
        template <int R2, int C2, int O2> DiagProxy& operator<<(const Matrix<T, R2, C2, O2>& v) { for (int i = 0; i < m->r_ && i < m->c_; ++i) (*m)(i, i) = v[i]; return *this; }
    };
    DiagProxy diagonal() { return DiagProxy{this}; }
    // (M).colwise() + v  (CDFGridPVC.cpp:220)
    struct ColwiseProxy {
        Matrix<T, Dynamic, Dynamic> m;
        template <int R2, int C2, int O2> Matrix<T, Dynamic, Dynamic> operator+(const Matrix<T, R2, C2, O2>& v) const {
            Matrix<T, Dynamic, Dynamic> o(m.rows(), m.cols());
            for (int i = 0; i < m.rows(); ++i) for (int j = 0; j < m.cols(); ++j) o(i, j) = m(i, j) + v[i];
            return o;
        }
    };
    ColwiseProxy colwise() const { return ColwiseProxy{Matrix<T, Dynamic, Dynamic>(*this)}; }

    // dynamic sub-blocks: leftCols / rightCols / topRows / bottomRows, assignable and chainable
    struct DBlock {
        Matrix* m; int i0, j0, r, c;
        int rows() const { return r; }
        int cols() const { return c; }
        T operator()(int i, int j) const { return (*m)(i0 + i, j0 + j); }
        template <int R2, int C2, int O2> DBlock& operator=(const Matrix<T, R2, C2, O2>& o) { for (int i = 0; i < r; ++i) for (int j = 0; j < c; ++j) (*m)(i0 + i, j0 + j) = o(i, j); return *this; }
        operator Matrix<T, Dynamic, Dynamic>() const { Matrix<T, Dynamic, Dynamic> o(r, c); for (int i = 0; i < r; ++i) for (int j = 0; j < c; ++j) o(i, j) = (*this)(i, j); return o; }
        DBlock leftCols(int n) const { return DBlock{m, i0, j0, r, n}; }
        DBlock rightCols(int n) const { return DBlock{m, i0, j0 + c - n, r, n}; }
        DBlock topRows(int n) const { return DBlock{m, i0, j0, n, c}; }
        DBlock bottomRows(int n) const { return DBlock{m, i0 + r - n, j0, n, c}; }
    };
    DBlock leftCols(int n) { return DBlock{this, 0, 0, r_, n}; }
    DBlock rightCols(int n) { return DBlock{this, 0, c_ - n, r_, n}; }
    DBlock topRows(int n) { return DBlock{this, 0, 0, n, c_}; }
    DBlock bottomRows(int n) { return DBlock{this, r_ - n, 0, n, c_}; }

    // fixed-size block as an assignable proxy / readable value
    template <int BR, int BC>
    struct Block {
        Matrix* m; int i0, j0;
        operator Matrix<T, BR, BC>() const { Matrix<T, BR, BC> b; for (int i = 0; i < BR; ++i) for (int j = 0; j < BC; ++j) b(i, j) = (*m)(i0 + i, j0 + j); return b; }
        template <int R2, int C2, int O2>
        Block& operator=(const Matrix<T, R2, C2, O2>& o) { for (int i = 0; i < BR; ++i) for (int j = 0; j < BC; ++j) (*m)(i0 + i, j0 + j) = o(i, j); return *this; }
        T operator()(int i, int j) const { return (*m)(i0 + i, j0 + j); }
        int rows() const { return BR; }
        int cols() const { return BC; }
    };
    template <int BR, int BC> Block<BR, BC> block(int i0, int j0) { return Block<BR, BC>{this, i0, j0}; }
    template <int BR, int BC> Matrix<T, BR, BC> block(int i0, int j0) const {
        Matrix<T, BR, BC> b;
        for (int i = 0; i < BR; ++i) for (int j = 0; j < BC; ++j) b(i, j) = (*this)(i0 + i, j0 + j);
        return b;
    }

private:
    int r_, c_;
    std::vector<T> d_;
    template <class, int, int, int> friend class Matrix;
};

template <class T, int R1, int C1, int O1, int R2, int C2, int O2>
Matrix<T, Dynamic, Dynamic> operator*(const Matrix<T, R1, C1, O1>& a, const Matrix<T, R2, C2, O2>& b) {
    assert(a.cols() == b.rows());
    Matrix<T, Dynamic, Dynamic> m(a.rows(), b.cols());
    for (int i = 0; i < a.rows(); ++i)
        for (int j = 0; j < b.cols(); ++j) {
            T acc = a(i, 0) * b(0, j);
            for (int k = 1; k < a.cols(); ++k) acc = acc + a(i, k) * b(k, j);
            m(i, j) = acc;
        }
    return m;
}
template <class T, int R1, int C1, int O1, int R2, int C2, int O2>
Matrix<T, Dynamic, Dynamic> operator+(const Matrix<T, R1, C1, O1>& a, const Matrix<T, R2, C2, O2>& b) {
    Matrix<T, Dynamic, Dynamic> m(a.rows(), a.cols());
    for (int i = 0; i < a.rows(); ++i) for (int j = 0; j < a.cols(); ++j) m(i, j) = a(i, j) + b(i, j);
    return m;
}
template <class T, int R1, int C1, int O1, int R2, int C2, int O2>
Matrix<T, Dynamic, Dynamic> operator-(const Matrix<T, R1, C1, O1>& a, const Matrix<T, R2, C2, O2>& b) {
    Matrix<T, Dynamic, Dynamic> m(a.rows(), a.cols());
    for (int i = 0; i < a.rows(); ++i) for (int j = 0; j < a.cols(); ++j) m(i, j) = a(i, j) - b(i, j);
    return m;
}
template <class T, int R1, int C1, int O1>
Matrix<T, Dynamic, Dynamic> operator*(const Matrix<T, R1, C1, O1>& a, T s) {
    Matrix<T, Dynamic, Dynamic> m(a.rows(), a.cols());
    for (int i = 0; i < a.rows(); ++i) for (int j = 0; j < a.cols(); ++j) m(i, j) = a(i, j) * s;
    return m;
}
template <class T, int R1, int C1, int O1>
Matrix<T, Dynamic, Dynamic> operator*(T s, const Matrix<T, R1, C1, O1>& a) {
    Matrix<T, Dynamic, Dynamic> m(a.rows(), a.cols());
    for (int i = 0; i < a.rows(); ++i) for (int j = 0; j < a.cols(); ++j) m(i, j) = s * a(i, j);
    return m;
}

// int scalars (2 * cov3_sqrt, RealVectorBeliefSpace.cpp:42)
template <class T, int R1, int C1, int O1>
Matrix<T, Dynamic, Dynamic> operator*(int s, const Matrix<T, R1, C1, O1>& a) { return T(s) * a; }
template <class T, int R1, int C1, int O1>
Matrix<T, Dynamic, Dynamic> operator*(const Matrix<T, R1, C1, O1>& a, int s) { return a * T(s); }
template <class T, int R1, int C1, int O1>
std::ostream& operator<<(std::ostream& os, const Matrix<T, R1, C1, O1>& a) {
    for (int i = 0; i < a.rows(); ++i) { for (int j = 0; j < a.cols(); ++j) os << (j ? " " : "") << a(i, j); if (i + 1 < a.rows()) os << "\n"; }
    return os;
}

// matrix_exp_compute (unsupported/Eigen/src/MatrixFunctions/MatrixExponential.h), double
template <class T, int R, int C, int Opt>
Matrix<T, Dynamic, Dynamic> Matrix<T, R, C, Opt>::exp() const {
    using M = Matrix<T, Dynamic, Dynamic>;
    const int n = r_;
    M A(*this);
    T l1 = T(0);                                   // cwiseAbs().colwise().sum().maxCoeff()
    for (int j = 0; j < n; ++j) { T s = T(0); for (int i = 0; i < n; ++i) s += std::fabs(A(i, j)); if (s > l1) l1 = s; }
    const M I = M::Identity(n, n);
    M U, V;
    int squarings = 0;
    if (l1 < 1.495585217958292e-002) {             // matrix_exp_pade3
        const T b[] = {120.0, 60.0, 12.0, 1.0};
        const M A2 = A * A;
        const M tmp = b[3] * A2 + b[1] * I;
        U = A * tmp; V = b[2] * A2 + b[0] * I;
    } else if (l1 < 2.539398330063230e-001) {      // matrix_exp_pade5
        const T b[] = {30240.0, 15120.0, 3360.0, 420.0, 30.0, 1.0};
        const M A2 = A * A, A4 = A2 * A2;
        const M tmp = b[5] * A4 + b[3] * A2 + b[1] * I;
        U = A * tmp; V = b[4] * A4 + b[2] * A2 + b[0] * I;
    } else if (l1 < 9.504178996162932e-001) {      // matrix_exp_pade7
        const T b[] = {17297280.0, 8648640.0, 1995840.0, 277200.0, 25200.0, 1512.0, 56.0, 1.0};
        const M A2 = A * A, A4 = A2 * A2, A6 = A4 * A2;
        const M tmp = b[7] * A6 + b[5] * A4 + b[3] * A2 + b[1] * I;
        U = A * tmp; V = b[6] * A6 + b[4] * A4 + b[2] * A2 + b[0] * I;
    } else if (l1 < 2.097847961257068e+000) {      // matrix_exp_pade9
        const T b[] = {17643225600.0, 8821612800.0, 2075673600.0, 302702400.0, 30270240.0, 2162160.0, 110880.0, 3960.0, 90.0, 1.0};
        const M A2 = A * A, A4 = A2 * A2, A6 = A4 * A2, A8 = A6 * A2;
        const M tmp = b[9] * A8 + b[7] * A6 + b[5] * A4 + b[3] * A2 + b[1] * I;
        U = A * tmp; V = b[8] * A8 + b[6] * A6 + b[4] * A4 + b[2] * A2 + b[0] * I;
    } else {                                       // matrix_exp_pade13 with scaling
        const T maxnorm = 5.371920351148152;
        std::frexp(l1 / maxnorm, &squarings);
        if (squarings < 0) squarings = 0;
        A = A * std::ldexp(T(1), -squarings);
        const T b[] = {64764752532480000.0, 32382376266240000.0, 7771770303897600.0, 1187353796428800.0, 129060195264000.0, 10559470521600.0,
                       670442572800.0, 33522128640.0, 1323241920.0, 40840800.0, 960960.0, 16380.0, 182.0, 1.0};
        const M A2 = A * A, A4 = A2 * A2, A6 = A4 * A2;
        V = b[13] * A6 + b[11] * A4 + b[9] * A2;
        M tmp = A6 * V;
        tmp = tmp + (b[7] * A6 + b[5] * A4 + b[3] * A2 + b[1] * I);
        U = A * tmp;
        tmp = b[12] * A6 + b[10] * A4 + b[8] * A2;
        V = A6 * tmp;
        V = V + (b[6] * A6 + b[4] * A4 + b[2] * A2 + b[0] * I);
    }
    const M numer = U + V, denom = V - U;          // result = denom.partialPivLu().solve(numer)
    M Rm = M::lu_solve(denom, numer);
    for (int i = 0; i < squarings; ++i) Rm = Rm * Rm;
    return Rm;
}

using MatrixXd = Matrix<double, Dynamic, Dynamic>;
using VectorXd = Matrix<double, Dynamic, 1>;
using Matrix2d = Matrix<double, 2, 2>;
using Vector2d = Matrix<double, 2, 1>;
using Matrix4d = Matrix<double, 4, 4>;
using Vector4d = Matrix<double, 4, 1>;

// SelfAdjointEigenSolver: reads the LOWER triangle only (Eigen's convention), cyclic Jacobi rotations.
template <class MatrixType>
class SelfAdjointEigenSolver {
public:
    SelfAdjointEigenSolver() = default;
    template <class M> explicit SelfAdjointEigenSolver(const M& a) { compute(a); }
    template <class M> SelfAdjointEigenSolver& compute(const M& a) {
        const int n = a.rows();
        MatrixXd A(n, n);
        for (int i = 0; i < n; ++i) for (int j = 0; j <= i; ++j) { A(i, j) = a(i, j); A(j, i) = a(i, j); }
        vec_ = MatrixXd::Identity(n, n);
        for (int sweep = 0; sweep < 100; ++sweep) {
            double off = 0, diag = 0;
            for (int i = 0; i < n; ++i) { diag += A(i, i) * A(i, i); for (int j = 0; j < i; ++j) off += A(i, j) * A(i, j); }
            if (off <= 1e-40 * diag || off == 0) break;
            for (int p = 0; p < n; ++p)
                for (int q = p + 1; q < n; ++q) {
                    if (A(p, q) == 0) continue;
                    const double theta = (A(q, q) - A(p, p)) / (2 * A(p, q));
                    const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1));
                    const double c = 1 / std::sqrt(t * t + 1), s = t * c;
                    for (int k = 0; k < n; ++k) { const double akp = A(k, p), akq = A(k, q); A(k, p) = c * akp - s * akq; A(k, q) = s * akp + c * akq; }
                    for (int k = 0; k < n; ++k) { const double apk = A(p, k), aqk = A(q, k); A(p, k) = c * apk - s * aqk; A(q, k) = s * apk + c * aqk; }
                    for (int k = 0; k < n; ++k) { const double vkp = vec_(k, p), vkq = vec_(k, q); vec_(k, p) = c * vkp - s * vkq; vec_(k, q) = s * vkp + c * vkq; }
                }
        }
        val_ = MatrixXd(n, 1);
        for (int i = 0; i < n; ++i) val_(i, 0) = A(i, i);
        return *this;
    }
    const MatrixXd& eigenvalues() const { return val_; }
    const MatrixXd& eigenvectors() const { return vec_; }
    // m_eivec * m_eivalues.cwiseSqrt().asDiagonal() * m_eivec.adjoint()
    MatrixXd operatorSqrt() const {
        const int n = val_.rows();
        MatrixXd D(n, n);
        for (int i = 0; i < n; ++i) D(i, i) = std::sqrt(val_(i, 0));
        return vec_ * D * vec_.transpose();
    }
private:
    MatrixXd val_, vec_;
};

// EigenSolver<Matrix2d> (general real matrix): RealSchur on the matrix scaled by its largest coefficient; for 2x2 the
// Hessenberg step is trivial and the whole iteration is one splitOffTwoRows.  Complex pairs: real part = trace / 2.
template <class MatrixType>
class EigenSolver {
public:
    struct ComplexVec {
        double re[2], im[2];
        Matrix<double, 2, 1> real() const { return Matrix<double, 2, 1>(re[0], re[1]); }
    };
    struct ComplexMat {
        Matrix<double, 2, 2> re;
        Matrix<double, 2, 2> real() const { return re; }
    };
    template <class M> EigenSolver& compute(const M& a) {
        assert(a.rows() == 2 && a.cols() == 2);
        double scale = 0;
        for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) scale = std::max(scale, std::fabs(a(i, j)));
        vec_.re = Matrix<double, 2, 2>::Identity();
        if (scale == 0 || !std::isfinite(scale)) { val_ = ComplexVec{{a(0, 0), a(1, 1)}, {0, 0}}; return *this; }
        double t00 = a(0, 0) / scale, t01 = a(0, 1) / scale, t10 = a(1, 0) / scale, t11 = a(1, 1) / scale;
        double u00 = 1, u01 = 0, u10 = 0, u11 = 1;
        const double eps = std::numeric_limits<double>::epsilon();
        bool complex_pair = false;
        // findSmallSubdiagEntry: |T(1,0)| <= eps * (|T(0,0)| + |T(1,1)|) (or below the "consider as zero" floor) deflates
        double s = std::fabs(t00) + std::fabs(t11);
        const double normv = std::fabs(t00) + std::fabs(t01) + std::fabs(t10) + std::fabs(t11);
        const double considerAsZero = std::max(normv * eps * eps, std::numeric_limits<double>::min());
        if (!(std::fabs(t10) <= std::max(eps * s, considerAsZero))) {
            // splitOffTwoRows(iu = 1)
            const double p = 0.5 * (t00 - t11);
            const double q = p * p + t10 * t01;
            if (q >= 0) {
                const double z = std::sqrt(std::fabs(q));
                // JacobiRotation::makeGivens(p +- z, T(1,0))
                const double gp = (p >= 0) ? p + z : p - z, gq = t10;
                double c, sn;
                if (gq == 0) { c = gp < 0 ? -1 : 1; sn = 0; }
                else if (gp == 0) { c = 0; sn = gq < 0 ? 1 : -1; }
                else if (std::fabs(gp) > std::fabs(gq)) { const double t = gq / gp; double u = std::sqrt(1 + t * t); if (gp < 0) u = -u; c = 1 / u; sn = -t * c; }
                else { const double t = gp / gq; double u = std::sqrt(1 + t * t); if (gq < 0) u = -u; sn = -1 / u; c = -t * sn; }
                // T.applyOnTheLeft(0, 1, rot.adjoint()): adjoint = (c, -s); rows x, y: x' = c x - s y ... with j = (c, -sn):
                // x' = j.c x + j.s y, y' = -j.s x + j.c y
                { const double jc = c, js = -sn;
                  const double x0 = t00, x1 = t01, y0 = t10, y1 = t11;
                  t00 = jc * x0 + js * y0; t01 = jc * x1 + js * y1; t10 = -js * x0 + jc * y0; t11 = -js * x1 + jc * y1; }
                // T.applyOnTheRight(0, 1, rot): columns x, y with j = rot.transpose() = (c, -sn)
                { const double jc = c, js = -sn;
                  const double x0 = t00, x1 = t10, y0 = t01, y1 = t11;
                  t00 = jc * x0 + js * y0; t10 = jc * x1 + js * y1; t01 = -js * x0 + jc * y0; t11 = -js * x1 + jc * y1; }
                t10 = 0;
                { const double jc = c, js = -sn;
                  const double x0 = u00, x1 = u10, y0 = u01, y1 = u11;
                  u00 = jc * x0 + js * y0; u10 = jc * x1 + js * y1; u01 = -js * x0 + jc * y0; u11 = -js * x1 + jc * y1; }
            } else complex_pair = true;
        } else t10 = 0;
        t00 *= scale; t01 *= scale; t10 *= scale; t11 *= scale;
        if (complex_pair) {
            // EigenSolver::compute: p = 0.5 (T(0,0) - T(1,1)); z = maxVal * sqrt(|p0^2 + t0 t1|) ...; real part T(1,1) + p
            const double p = 0.5 * (t00 - t11);
            const double maxval = std::max(std::fabs(p), std::max(std::fabs(t10), std::fabs(t01)));
            double z = 0;
            if (maxval != 0) { const double t0 = t10 / maxval, t1 = t01 / maxval, p0 = p / maxval; z = maxval * std::sqrt(std::fabs(p0 * p0 + t0 * t1)); }
            val_ = ComplexVec{{t11 + p, t11 + p}, {z, -z}};
            vec_.re = Matrix<double, 2, 2>::Identity();   // complex eigenvectors: not needed by the reference's call sites
            return *this;
        }
        val_ = ComplexVec{{t00, t11}, {0, 0}};
        // doComputeEigenvectors, real case: back substitution on the upper-triangular T, then V = U * X, columns normalised
        // column 0: x = (1, 0); column 1: x = (r, 1) with (t00 - t11) r + t01 = 0
        double x01;
        {
            const double w = t00 - t11;
            const double norm = std::fabs(t00) + std::fabs(t01) + std::fabs(t11);
            x01 = (w != 0) ? -t01 / w : -t01 / (eps * norm);
            // overflow control of the reference implementation: t = |x01|; if (eps * t) * t > 1 -> rescale (never here)
        }
        double v00 = u00, v10 = u10, v01 = u00 * x01 + u01, v11 = u10 * x01 + u11;
        const double n0 = std::sqrt(v00 * v00 + v10 * v10), n1 = std::sqrt(v01 * v01 + v11 * v11);
        vec_.re(0, 0) = v00 / n0; vec_.re(1, 0) = v10 / n0; vec_.re(0, 1) = v01 / n1; vec_.re(1, 1) = v11 / n1;
        return *this;
    }
    const ComplexVec& eigenvalues() const { return val_; }
    const ComplexMat& eigenvectors() const { return vec_; }
private:
    ComplexVec val_{};
    ComplexMat vec_{};
};

// LLT (Cholesky), lower triangle only
template <class MatrixType>
class LLT {
public:
    template <class M> explicit LLT(const M& a) {
        const int n = a.rows();
        L_ = MatrixXd(n, n);
        for (int j = 0; j < n; ++j) {
            double d = a(j, j);
            for (int k = 0; k < j; ++k) d -= L_(j, k) * L_(j, k);
            L_(j, j) = std::sqrt(d);
            for (int i = j + 1; i < n; ++i) { double v = a(i, j); for (int k = 0; k < j; ++k) v -= L_(i, k) * L_(j, k); L_(i, j) = v / L_(j, j); }
        }
    }
    const MatrixXd& matrixL() const { return L_; }
private:
    MatrixXd L_;
};

}  // namespace Eigen
