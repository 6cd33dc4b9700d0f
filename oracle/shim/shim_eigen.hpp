// shim_eigen.hpp — TEST INFRASTRUCTURE.  A minimal stand-in for the subset of Eigen 3 that the reference's
// linear propagator and PCCBlackmoreSVC touch, so that those reference SOURCES (compiled where they lie under
// /root/reference) can run here without Eigen.  Eager evaluation, coefficient (i,j) of a product is the sum
// over k in increasing order.  Every product on this path has inner dimension 2, where the order of the two
// terms cannot change the rounded sum, so the values equal what any Eigen build without FMA produces
// (the reference builds for baseline x86-64: CMakeLists.txt:6-8).  The 2x2 inverse follows Eigen's
// compute_inverse_size2_helper (adjugate times 1/det).  NOT a general Eigen replacement.
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
#include <vector>

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
    void resize(int r, int c) { if (r != r_ || c != c_) { r_ = r; c_ = c; d_.assign((size_t)r * c, T(0)); } }
    T& operator()(int i, int j) { return d_[(size_t)i * c_ + j]; }
    const T& operator()(int i, int j) const { return d_[(size_t)i * c_ + j]; }
    T& operator[](int i) { return d_[i]; }
    const T& operator[](int i) const { return d_[i]; }
    T value() const { assert(r_ == 1 && c_ == 1); return d_[0]; }

    static Matrix Zero() { return Matrix(); }
    static Matrix Zero(int r, int c) { return Matrix(r, c); }
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
    Matrix inverse() const {
        assert(r_ == 2 && c_ == 2);
        const T det = (*this)(0, 0) * (*this)(1, 1) - (*this)(1, 0) * (*this)(0, 1);
        const T invdet = T(1) / det;
        Matrix m(2, 2);
        m(0, 0) = (*this)(1, 1) * invdet;
        m(1, 0) = -(*this)(1, 0) * invdet;
        m(0, 1) = -(*this)(0, 1) * invdet;
        m(1, 1) = (*this)(0, 0) * invdet;
        return m;
    }

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

using MatrixXd = Matrix<double, Dynamic, Dynamic>;
using VectorXd = Matrix<double, Dynamic, 1>;
using Matrix2d = Matrix<double, 2, 2>;
using Vector2d = Matrix<double, 2, 1>;

}  // namespace Eigen
