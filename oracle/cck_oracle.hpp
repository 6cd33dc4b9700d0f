// cck_oracle.hpp — CPU ORACLE (test infrastructure, NOT product code).
//
// A dependency-free C++17 restatement of the reference's belief propagation +
// chance-constraint checking hot path.  Only tests/, bench.py's
// cpu_baseline / --impl reference leg and __graft_entry__.smoke() may use it, and
// only as the checker / the CPU baseline.  The product (libcck_b200.so) never
// links or calls anything in this directory.
//
// PARITY: the reference ships no tests, golden vectors or fixtures and its CMake
// build needs OMPL 1.6.0, Eigen3 and Boost (absent, no network).
//   PINNED (bit for bit, tests/test_ref_pin.py): the default-configuration path — Instance/Robot/Obstacle parsing
//   and geometry, R2_UncertainLinearStatePropagator::propagate, PCCBlackmoreSVC (Minkowski half-planes, erf_inv
//   constant, isValid), MinkowskiSumBlackmorePVC / BoundingBoxBlackmorePVC / BeliefPVC (validatePlan,
//   createConstraint, satisfiesConstraints) — against the reference's OWN sources compiled where they lie against the stand-in
//   headers of oracle/shim/ (oracle/_ref, oracle/ref_capi.cpp) and against the vectors generated from that
//   build (tests/golden/ref_*.npz).
//   PARITY UNPINNED for everything else (unicycle, adaptive / chi-squared checkers, goal, distance): checked only against (i) closed-form known answers derived from the reference's
//   formulas, (ii) scipy for erf_inv / chi-square quantiles, (iii) self-consistency.
// Third-party arithmetic restated (source not under /root/reference):
//   Boost.Math erf_inv, quantile(chi_squared(2)); Eigen 2x2/4x4 inverse(),
//   EigenSolver<Matrix2d>, MatrixXd::exp() (nilpotent input); Boost.Geometry
//   correct / buffer(point_circle(10)) / disjoint / within; OMPL 1.6
//   propagateWhileValid / PathControl::interpolate / satisfiesBounds.
//
// Every function cites the reference file:line it follows (paths relative to
// /root/reference).  Arithmetic is written in the reference's operation order;
// compile with -ffp-contract=off (the reference builds for baseline x86-64:
// no FMA; CMakeLists.txt:6-8 has no -march flag).
#pragma once
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <limits>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <sstream>
#include <string>
#include <vector>

namespace cck_oracle {

// ---------------------------------------------------------------------------
// Special functions (Boost.Math restated)
// ---------------------------------------------------------------------------

// boost::math::erf_inv(z) for double (call sites:
// include/StateValidityCheckers/PCCBlackmoreSVC.h:34-36,
// src/PlanValidityCheckers/MinkowskiSumBlackmorePVC.cpp:178, ...).
// Boost evaluates in long double (promote_double policy) and rounds; we solve
// erf(x)=z by Newton/Halley iterations on erfl/erfcl in long double (<1 ulp).
inline double erf_inv(double z) {
    if (std::isnan(z) || z < -1.0 || z > 1.0) return std::nan("");
    if (z == 1.0) return HUGE_VAL;
    if (z == -1.0) return -HUGE_VAL;
    if (z == 0.0) return 0.0;
    const bool neg = z < 0;
    const long double p = neg ? -(long double)z : (long double)z;
    const long double q = 1.0L - p;  // exact for p >= 0.5 in long double
    // initial guess: Winitzki-style approximation
    const long double a = 0.147L;
    const long double ln1 = logl(1.0L - p * p);
    const long double t1 = 2.0L / (3.14159265358979323846264338327950288L * a) + ln1 / 2.0L;
    long double x = sqrtl(sqrtl(t1 * t1 - ln1 / a) - t1);
    const long double two_over_sqrtpi = 1.12837916709551257389615890312154517L;
    for (int it = 0; it < 60; ++it) {
        long double f;
        if (p > 0.5L) f = q - erfcl(x);   // erf(x) - p == q - erfc(x)
        else          f = erfl(x) - p;
        const long double d = two_over_sqrtpi * expl(-x * x);
        // Halley step: x_{n+1} = x - f/(d + x f)
        const long double dx = f / (d + x * f);
        x -= dx;
        if (fabsl(dx) <= 1e-21L * fabsl(x)) break;
    }
    const double r = (double)x;
    return neg ? -r : r;
}

// quantile(boost::math::chi_squared(v=2), p)
// (include/StateValidityCheckers/ChiSquaredBoundarySVC.h:24-27,
//  include/PlanValidityCheckers/ChiSquaredBoundaryPVC.h:22-25).
// For 2 degrees of freedom the CDF is 1-exp(-x/2), so the quantile is -2 ln(1-p).
inline double chi2_quantile_2dof(double p) { return -2.0 * std::log1p(-p); }

// ---------------------------------------------------------------------------
// Belief state  (include/Spaces/RealVectorBeliefSpace.h:21-96,
//                src/Spaces/RealVectorBeliefSpace.cpp:7-14)
// AoS record: values[dim], sigma[dim*dim] row-major, lambda[dim*dim] row-major.
// ---------------------------------------------------------------------------
template <int D>
struct Belief {
    double v[D];
    double S[D * D];
    double L[D * D];
};
using Belief2 = Belief<2>;
using Belief4 = Belief<4>;

// allocState default: sigma_=lambda_=0.01*I  (RealVectorBeliefSpace.cpp:11-12)
template <int D>
inline void belief_default(Belief<D>& b) {
    for (int i = 0; i < D; ++i) b.v[i] = 0.0;
    for (int i = 0; i < D * D; ++i) b.S[i] = b.L[i] = 0.0;
    for (int i = 0; i < D; ++i) b.S[i * D + i] = b.L[i * D + i] = 0.01;
}

// ompl::base::RealVectorStateSpace::satisfiesBounds (OMPL 1.6.0, restated; called
// first by every SVC, e.g. PCCBlackmoreSVC.cpp:35).  A NaN component passes.
inline bool satisfies_bounds(const double* v, const double* low, const double* high, int dim) {
    const double eps = DBL_EPSILON;
    for (int i = 0; i < dim; ++i)
        if (v[i] - eps > high[i] || v[i] + eps < low[i]) return false;
    return true;
}

// ---------------------------------------------------------------------------
// a1: R2_UncertainLinearStatePropagator (src/StatePropogators/UncertainLinearSP.cpp)
// ---------------------------------------------------------------------------
struct LinearParams {
    double dt;       // duration_ = si->getPropagationStepSize()   (:5)
    double Q[4];     // pow(0.1,2)*I   (:37-38)
    double R[4];     // pow(0.1,2)*I   (:39-40)
    double Acl[4];   // A_ol - B_ol*K_ = (1-0.3)*I   (:22, hdr:50)
};

inline LinearParams make_linear_params(double dt, double process_noise = 0.1,
                                       double measurement_noise = 0.1, double K = 0.3) {
    LinearParams p{};
    p.dt = dt;
    const double q = std::pow(process_noise, 2), r = std::pow(measurement_noise, 2);
    p.Q[0] = q * 1.0; p.Q[1] = q * 0.0; p.Q[2] = q * 0.0; p.Q[3] = q * 1.0;
    p.R[0] = r * 1.0; p.R[1] = r * 0.0; p.R[2] = r * 0.0; p.R[3] = r * 1.0;
    // A_cl_ = A_ol_ - B_ol_*K_ with A_ol_=B_ol_=I
    p.Acl[0] = 1.0 - 1.0 * K; p.Acl[1] = 0.0 - 0.0 * K;
    p.Acl[2] = 0.0 - 0.0 * K; p.Acl[3] = 1.0 - 1.0 * K;
    return p;
}

// 2x2 row-major product C = A*B, coefficient (i,j) = A(i,0)B(0,j) + A(i,1)B(1,j)
inline void mm2(const double* A, const double* B, double* C) {
    double t[4];
    t[0] = A[0] * B[0] + A[1] * B[2];
    t[1] = A[0] * B[1] + A[1] * B[3];
    t[2] = A[2] * B[0] + A[3] * B[2];
    t[3] = A[2] * B[1] + A[3] * B[3];
    C[0] = t[0]; C[1] = t[1]; C[2] = t[2]; C[3] = t[3];
}

// Eigen fixed 2x2 inverse(): invdet = 1/det, adjugate * invdet
// (Eigen/src/LU/InverseImpl.h compute_inverse_size2_helper, restated)
inline void inv2(const double* M, double* R) {
    const double det = M[0] * M[3] - M[2] * M[1];
    const double invdet = 1.0 / det;
    R[0] = M[3] * invdet;
    R[2] = -M[2] * invdet;
    R[1] = -M[1] * invdet;
    R[3] = M[0] * invdet;
}

// UncertainLinearSP.cpp:58-111.  NOTE (quirk 1): the `duration` argument is
// ignored, the member duration_ is used (:92-93).  F_=H_=I_ (hdr:45-47) make
// F*S*F and H*S*H^T exact no-ops, so they are elided.
inline void linear_propagate(const Belief2& s, const double* u, const LinearParams& p, Belief2& out) {
    const double mx = s.v[0] + p.dt * u[0];                     // :92
    const double my = s.v[1] + p.dt * u[1];                     // :93
    double sp[4], S[4], Si[4], K[4], lp[4], t[4], ImK[4], ns[4], KP[4];
    for (int i = 0; i < 4; ++i) sp[i] = s.S[i] + p.Q[i];        // :98
    for (int i = 0; i < 4; ++i) S[i] = sp[i] + p.R[i];          // :100
    inv2(S, Si);
    mm2(sp, Si, K);                                             // :101
    mm2(p.Acl, s.L, t); mm2(t, p.Acl, lp);                      // :102 (A_cl*L*A_cl, no transpose)
    ImK[0] = 1.0 - K[0]; ImK[1] = 0.0 - K[1];
    ImK[2] = 0.0 - K[2]; ImK[3] = 1.0 - K[3];
    mm2(ImK, sp, ns);                                           // :103
    mm2(K, sp, KP);                                             // :104
    out.v[0] = mx; out.v[1] = my;
    for (int i = 0; i < 4; ++i) out.S[i] = ns[i];
    for (int i = 0; i < 4; ++i) out.L[i] = lp[i] + KP[i];
}

// ---------------------------------------------------------------------------
// a2: DynUnicycleControlSpace (src/StatePropogators/UncertainUnicycleSP.cpp)
// ---------------------------------------------------------------------------
struct UnicycleParams {
    double F[16];      // A_ol_d = exp(A_ol*dt) = I + dt*(e0e1^T + e2e3^T)   (:88-91,114)
    double Acl[16];    // A_cl_d_ = A_ol_d - B_ol_d*K_   (:106)
    double Q[16], R[16];   // pow(0.1,2)*I4   (:123-127)
    double gains[8];   // controller_parameters_  (:76-83)
    double acc_lo, acc_hi, turn_lo, turn_hi;   // (:22-25)
};

inline UnicycleParams make_unicycle_params(double dt) {
    UnicycleParams p{};
    // exp of the nilpotent A_ol*dt is I + A_ol*dt exactly; Eigen's Pade-5 path
    // reproduces it to rounding (see DESIGN.md "oracle notes").
    for (int i = 0; i < 16; ++i) p.F[i] = 0.0;
    for (int i = 0; i < 4; ++i) p.F[i * 4 + i] = 1.0;
    p.F[0 * 4 + 1] = dt; p.F[2 * 4 + 3] = dt;
    // zero-order-hold B_ol_d from exp([[A,B],[0,0]]*dt) = I + N + N^2/2  (:95-104)
    double Bd[4][2] = {{dt * dt / 2.0, 0.0}, {dt, 0.0}, {0.0, dt * dt / 2.0}, {0.0, dt}};
    const double Kg[2][4] = {{2.5857, 3.4434, 0.0, 0.0}, {0.0, 0.0, 2.5857, 3.4434}};  // :58-65
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            const double bk = Bd[i][0] * Kg[0][j] + Bd[i][1] * Kg[1][j];
            p.Acl[i * 4 + j] = p.F[i * 4 + j] - bk;
        }
    const double q = std::pow(0.1, 2), r = std::pow(0.1, 2);
    for (int i = 0; i < 16; ++i) p.Q[i] = p.R[i] = 0.0;
    for (int i = 0; i < 4; ++i) { p.Q[i * 4 + i] = q * 1.0; p.R[i * 4 + i] = r * 1.0; }
    const double g[8] = {1.0, 1.0, 0.0, 0.0, 0.0, 0.0, 1.0, 1.0};
    for (int i = 0; i < 8; ++i) p.gains[i] = g[i];
    p.acc_lo = -0.5; p.acc_hi = 0.5; p.turn_lo = -0.5; p.turn_hi = 0.5;
    return p;
}

inline void mm4(const double* A, const double* B, double* C) {
    double t[16];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double acc = A[i * 4 + 0] * B[0 * 4 + j];
            acc = acc + A[i * 4 + 1] * B[1 * 4 + j];
            acc = acc + A[i * 4 + 2] * B[2 * 4 + j];
            acc = acc + A[i * 4 + 3] * B[3 * 4 + j];
            t[i * 4 + j] = acc;
        }
    for (int i = 0; i < 16; ++i) C[i] = t[i];
}
inline void mm4_bt(const double* A, const double* B, double* C) {  // C = A * B^T
    double t[16];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double acc = A[i * 4 + 0] * B[j * 4 + 0];
            acc = acc + A[i * 4 + 1] * B[j * 4 + 1];
            acc = acc + A[i * 4 + 2] * B[j * 4 + 2];
            acc = acc + A[i * 4 + 3] * B[j * 4 + 3];
            t[i * 4 + j] = acc;
        }
    for (int i = 0; i < 16; ++i) C[i] = t[i];
}

// General 4x4 inverse by cofactors (Eigen's fixed-size 4x4 inverse() is also a
// cofactor expansion, vectorised; results agree to rounding, not bit-for-bit).
inline void inv4(const double* m, double* inv) {
    double a[16];
    a[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[14] - m[9] * m[6] * m[15] + m[9] * m[7] * m[14] + m[13] * m[6] * m[11] - m[13] * m[7] * m[10];
    a[4] = -m[4] * m[10] * m[15] + m[4] * m[11] * m[14] + m[8] * m[6] * m[15] - m[8] * m[7] * m[14] - m[12] * m[6] * m[11] + m[12] * m[7] * m[10];
    a[8] = m[4] * m[9] * m[15] - m[4] * m[11] * m[13] - m[8] * m[5] * m[15] + m[8] * m[7] * m[13] + m[12] * m[5] * m[11] - m[12] * m[7] * m[9];
    a[12] = -m[4] * m[9] * m[14] + m[4] * m[10] * m[13] + m[8] * m[5] * m[14] - m[8] * m[6] * m[13] - m[12] * m[5] * m[10] + m[12] * m[6] * m[9];
    a[1] = -m[1] * m[10] * m[15] + m[1] * m[11] * m[14] + m[9] * m[2] * m[15] - m[9] * m[3] * m[14] - m[13] * m[2] * m[11] + m[13] * m[3] * m[10];
    a[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[14] - m[8] * m[2] * m[15] + m[8] * m[3] * m[14] + m[12] * m[2] * m[11] - m[12] * m[3] * m[10];
    a[9] = -m[0] * m[9] * m[15] + m[0] * m[11] * m[13] + m[8] * m[1] * m[15] - m[8] * m[3] * m[13] - m[12] * m[1] * m[11] + m[12] * m[3] * m[9];
    a[13] = m[0] * m[9] * m[14] - m[0] * m[10] * m[13] - m[8] * m[1] * m[14] + m[8] * m[2] * m[13] + m[12] * m[1] * m[10] - m[12] * m[2] * m[9];
    a[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[14] - m[5] * m[2] * m[15] + m[5] * m[3] * m[14] + m[13] * m[2] * m[7] - m[13] * m[3] * m[6];
    a[6] = -m[0] * m[6] * m[15] + m[0] * m[7] * m[14] + m[4] * m[2] * m[15] - m[4] * m[3] * m[14] - m[12] * m[2] * m[7] + m[12] * m[3] * m[6];
    a[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[13] - m[4] * m[1] * m[15] + m[4] * m[3] * m[13] + m[12] * m[1] * m[7] - m[12] * m[3] * m[5];
    a[14] = -m[0] * m[5] * m[14] + m[0] * m[6] * m[13] + m[4] * m[1] * m[14] - m[4] * m[2] * m[13] - m[12] * m[1] * m[6] + m[12] * m[2] * m[5];
    a[3] = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] - m[9] * m[2] * m[7] + m[9] * m[3] * m[6];
    a[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[4] * m[2] * m[11] + m[4] * m[3] * m[10] + m[8] * m[2] * m[7] - m[8] * m[3] * m[6];
    a[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[9] + m[4] * m[1] * m[11] - m[4] * m[3] * m[9] - m[8] * m[1] * m[7] + m[8] * m[3] * m[5];
    a[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[9] - m[4] * m[1] * m[10] + m[4] * m[2] * m[9] + m[8] * m[1] * m[6] - m[8] * m[2] * m[5];
    const double det = m[0] * a[0] + m[1] * a[4] + m[2] * a[8] + m[3] * a[12];
    const double invdet = 1.0 / det;
    for (int i = 0; i < 16; ++i) inv[i] = a[i] * invdet;
}

inline void saturate(double& value, double lo, double hi) {   // UncertainUnicycleSP.cpp:132-135
    if (value < lo) value = lo;
    if (value > hi) value = hi;
}
inline double wrap_angle(double angle) {                       // :137-144
    const double k = std::floor((M_PI - angle) / (2 * M_PI));
    return angle + (2 * k * M_PI);
}

// UncertainUnicycleSP.cpp:146-227.  Uses the `duration` ARGUMENT (quirk 1);
// Sigma_pred = F*Sigma*F + Q without transpose (quirk 2); locals replace the
// reference's `mutable` members (u_bar_0, u_bar_1, surge_final).
inline void unicycle_propagate(const Belief4& s, const double* u, double duration,
                               const UnicycleParams& p, Belief4& out) {
    const double x = s.v[0], y = s.v[1], yaw = s.v[2], surge = s.v[3];
    const double cos_y = std::cos(yaw), sin_y = std::sin(yaw);
    const double xr = u[0], yr = u[1], yawr = u[2], surger = u[3];
    const double u0 = p.gains[0] * (xr - x) + p.gains[1] * (surger * std::cos(yawr) - surge * cos_y);   // :173
    const double u1 = p.gains[6] * (yr - y) + p.gains[7] * (surger * std::sin(yawr) - surge * sin_y);   // :174
    double ub0 = cos_y * u0 + sin_y * u1;                                                              // :179
    double ub1 = (-sin_y * u0 + cos_y * u1) / surge;                                                   // :180
    saturate(ub0, p.acc_lo, p.acc_hi);
    saturate(ub1, p.turn_lo, p.turn_hi);
    const double surge_final = surge + duration * ub0;                                                 // :190
    double S[16], Si[16], K[16], sp[16], t[16], lp[16], ImK[16], ns[16], KP[16];
    mm4(p.F, s.S, t); mm4(t, p.F, sp);                                                                  // :213
    for (int i = 0; i < 16; ++i) sp[i] = sp[i] + p.Q[i];
    for (int i = 0; i < 16; ++i) S[i] = sp[i] + p.R[i];                                                 // :215 (H=I)
    inv4(S, Si);
    mm4(sp, Si, K);                                                                                     // :216
    mm4(p.Acl, s.L, t); mm4_bt(t, p.Acl, lp);                                                           // :217
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) ImK[i * 4 + j] = (i == j ? 1.0 : 0.0) - K[i * 4 + j];
    mm4(ImK, sp, ns);                                                                                   // :219
    mm4(K, sp, KP);                                                                                     // :220
    out.v[0] = x + duration * cos_y * surge_final;                                                      // :202
    out.v[1] = y + duration * sin_y * surge_final;                                                      // :203
    out.v[2] = wrap_angle(yaw + duration * ub1);                                                        // :204
    out.v[3] = surge_final;                                                                             // :205
    for (int i = 0; i < 16; ++i) out.S[i] = ns[i];
    for (int i = 0; i < 16; ++i) out.L[i] = lp[i] + KP[i];
}

// ---------------------------------------------------------------------------
// a13: world model (src/utils/Instance.cpp, src/utils/common.cpp)
// ---------------------------------------------------------------------------
struct Pt { double x, y; };
using Ring = std::vector<Pt>;   // closed CCW ring: first == last (include/utils/common.h:19)

// RectangularObstacle ring (common.cpp:25-42): BL, BR, TR, TL, BL
inline Ring rect_ring(double x, double y, double len, double width) {
    Pt bl{x - (len / 2), y - (width / 2)}, br{x + (len / 2), y - (width / 2)};
    Pt tl{x - (len / 2), y + (width / 2)}, tr{x + (len / 2), y + (width / 2)};
    return Ring{bl, br, tr, tl, bl};
}

struct RobotShape {
    double bounding_rad;
    Ring bounding;   // square of half-side bounding_rad centred at the origin
};
// RectangularRobot + createBoundingShape (common.cpp:117-136, 61-108)
inline RobotShape rectangular_robot(double sx, double sy, double len, double width) {
    Ring shape = rect_ring(sx, sy, len, width);
    double max_rad = 0.0;
    for (const Pt& p : shape) {
        // matrix_transformer(cos0, sin0, 0-r_ix; -sin0, cos0, 0-r_iy)   (:72-76)
        const double x = 1.0 * p.x + 0.0 * p.y + (0 - sx);
        const double y = -0.0 * p.x + 1.0 * p.y + (0 - sy);
        const double r = std::sqrt((x * x) + (y * y));
        if (r > max_rad) max_rad = r;
    }
    RobotShape rs;
    rs.bounding_rad = max_rad;
    rs.bounding = Ring{{-max_rad, -max_rad}, {max_rad, -max_rad}, {max_rad, max_rad}, {-max_rad, max_rad}, {-max_rad, -max_rad}};
    return rs;
}

// getMinkowskiSum_ (PCCBlackmoreSVC.cpp:110-145; identical loop in
// MinkowskiSumBlackmorePVC.cpp:214-248)
inline Ring minkowski_sum(const Ring& shape1, const Ring& shape2) {
    Ring s1 = shape1, s2 = shape2;
    s1.push_back(s1[1]);
    s2.push_back(s2[1]);
    Ring result;
    size_t i = 0, j = 0;
    while (i < s1.size() - 2 || j < s2.size() - 2) {
        result.push_back(Pt{s1[i].x + s2[j].x, s1[i].y + s2[j].y});
        const double ax = s1[i + 1].x - s1[i].x, ay = s1[i + 1].y - s1[i].y;
        const double bx = s2[j + 1].x - s2[j].x, by = s2[j + 1].y - s2[j].y;
        const double cross = (ax * by) - (ay * bx);
        if (cross >= 0) ++i;
        if (cross <= 0) ++j;
        if (result.size() > 64) break;  // guard: malformed input
    }
    result.push_back(result.front());
    return result;
}

struct HalfPlanes { double A[4][2]; double B[4]; };
// getHalfPlanes_ (PCCBlackmoreSVC.cpp:77-108)
inline HalfPlanes half_planes(const Ring& ring) {
    HalfPlanes hp{};
    for (size_t i = 0; i + 1 < ring.size() && i < 4; ++i) {
        const double x1 = ring[i].x, y1 = ring[i].y, x2 = ring[i + 1].x, y2 = ring[i + 1].y;
        const double a = y2 - y1;
        const double b = x1 - x2;
        const double c = a * x1 + b * y1;
        hp.A[i][0] = a; hp.A[i][1] = b; hp.B[i] = c;
    }
    return hp;
}

struct Obstacle { double x, y; Ring ring; };
struct RobotSpec { std::string name, model; double sx, sy, gx, gy; RobotShape shape; };

struct Instance {
    double x_max = -1, y_max = -1;
    std::vector<Obstacle> obstacles;
    std::vector<RobotSpec> robots;
    double p_safe = 0, p_safe_obs = -1, p_safe_agents = -1;
    bool map_overrun = false;   // pos > size() hit (UB in the reference, quirk 13)
};

// Instance::load_map_ (Instance.cpp:78-115), incl. quirk 13: line[r] with
// r == line.size() reads '\0' (!= '.') => obstacle; r > size() is UB => flagged.
inline bool load_map_text(const std::string& text, Instance& inst) {
    std::istringstream f(text);
    std::string line;
    if (!std::getline(f, line)) return false;
    auto second_token = [](const std::string& l) {
        std::istringstream ls(l); std::string a, b; ls >> a >> b; return std::atoi(b.c_str());
    };
    if (!line.empty() && line[0] == 't') {
        std::getline(f, line); inst.y_max = second_token(line);
        std::getline(f, line); inst.x_max = second_token(line);
        std::getline(f, line);
    }
    if (!(inst.x_max > 0 && inst.y_max > 0)) return false;
    int c = 0;
    while (std::getline(f, line)) {
        for (int r = 0; r < inst.x_max; ++r) {
            char ch;
            if ((size_t)r < line.size()) ch = line[r];
            else if ((size_t)r == line.size()) ch = '\0';
            else { inst.map_overrun = true; continue; }
            if (ch != '.') inst.obstacles.push_back(Obstacle{(double)r, (double)c, rect_ring(r, c, 1, 1)});
        }
        c++;
    }
    return true;
}

// Instance::load_agents_ (Instance.cpp:117-174): tab-separated, header skipped
inline bool load_scen_text(const std::string& text, int num_agents, Instance& inst) {
    std::istringstream f(text);
    std::string line;
    std::getline(f, line);
    for (int i = 0; i < num_agents; ++i) {
        if (!std::getline(f, line) || line.empty()) return false;
        std::vector<std::string> tok;
        std::string cur;
        for (char ch : line) {
            if (ch == '\t') { if (!cur.empty()) tok.push_back(cur); cur.clear(); }
            else if (ch != '\r') cur.push_back(ch);
        }
        if (!cur.empty()) tok.push_back(cur);
        if (tok.size() < 10) return false;
        RobotSpec r;
        r.sx = std::atoi(tok[4].c_str()); r.sy = std::atoi(tok[5].c_str());
        r.gx = std::atoi(tok[6].c_str()); r.gy = std::atoi(tok[7].c_str());
        r.model = tok[9];
        r.name = "Robot " + std::to_string(i);
        if (tok[8] != "Rectangle") return false;   // quirk 11: belief checkers need Rectangle
        r.shape = rectangular_robot(r.sx, r.sy, 0.25, 0.25);   // :166
        inst.robots.push_back(r);
    }
    return true;
}

// risk split (Instance.cpp:38-44)
inline void split_risk(Instance& inst, double p_safe) {
    inst.p_safe = p_safe;
    const double total_things = (double)(inst.obstacles.size() + inst.robots.size());
    const double p_coll = 1 - p_safe;
    const double p_coll_obs = (p_coll * inst.obstacles.size()) / total_things;
    const double p_coll_agnts = (p_coll * inst.robots.size()) / total_things;
    inst.p_safe_agents = 1 - p_coll_agnts;
    inst.p_safe_obs = 1 - p_coll_obs;
}

// ---------------------------------------------------------------------------
// f2: RealVectorBeliefSpace::distance (src/Spaces/RealVectorBeliefSpace.cpp:16-62) and the two
// nearest-neighbour rules of the low-level planner (ConstraintRespectingBSST.cpp:119-178).
// Eigen's SelfAdjointEigenSolver (lower triangle, operatorSqrt = V sqrt(D) V^T) is restated as a
// cyclic Jacobi iteration; Eigen's tridiagonal QL agrees to rounding, not bit for bit.
// ---------------------------------------------------------------------------
inline void sym_operator_sqrt(const double* Ain, int n, double* out) {
    double A[16], V[16];
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) { A[i * n + j] = Ain[std::max(i, j) * n + std::min(i, j)]; V[i * n + j] = (i == j) ? 1.0 : 0.0; }
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0;
        for (int i = 0; i < n; ++i) for (int j = 0; j < i; ++j) off += A[i * n + j] * A[i * n + j];
        double dg = 0;                                    // same convergence test as the kernel (csrc/cck_tree.cuh)
        for (int i = 0; i < n; ++i) dg += A[i * n + i] * A[i * n + i];
        if (off < 1e-300 || off <= 1e-30 * dg) break;
        for (int p = 0; p < n; ++p)
            for (int q = p + 1; q < n; ++q) {
                if (std::fabs(A[p * n + q]) < 1e-300) continue;
                const double th = (A[q * n + q] - A[p * n + p]) / (2 * A[p * n + q]);
                const double t = (th >= 0 ? 1.0 : -1.0) / (std::fabs(th) + std::sqrt(th * th + 1));
                const double c = 1 / std::sqrt(t * t + 1), sn = t * c;
                for (int k = 0; k < n; ++k) { const double akp = A[k * n + p], akq = A[k * n + q]; A[k * n + p] = c * akp - sn * akq; A[k * n + q] = sn * akp + c * akq; }
                for (int k = 0; k < n; ++k) { const double apk = A[p * n + k], aqk = A[q * n + k]; A[p * n + k] = c * apk - sn * aqk; A[q * n + k] = sn * apk + c * aqk; }
                for (int k = 0; k < n; ++k) { const double vkp = V[k * n + p], vkq = V[k * n + q]; V[k * n + p] = c * vkp - sn * vkq; V[k * n + q] = sn * vkp + c * vkq; }
            }
    }
    double sq[4];
    for (int k = 0; k < n; ++k) sq[k] = std::sqrt(A[k * n + k]);   // m_eivalues.cwiseSqrt(): NaN for a negative eigenvalue, exactly as Eigen (pinned by tests/test_ref_pin2.py)
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            double acc = 0;
            for (int k = 0; k < n; ++k) acc += V[i * n + k] * sq[k] * V[j * n + k];
            out[i * n + j] = acc;
        }
}

// b1, b2: belief records [D + 2 D^2]
inline double belief_distance(const double* b1, const double* b2, int D) {
    double d2 = 0;                                                                    // :22-27
    for (int i = 0; i < D; ++i) { const double e = b1[i] - b2[i]; d2 += e * e; }
    if (d2 == 0) return 0.0;
    double C1[16], C2[16], S2[16], t[16], m[16], S3[16];
    for (int f = 0; f < D * D; ++f) { C1[f] = b1[D + f] + b1[D + D * D + f]; C2[f] = b2[D + f] + b2[D + D * D + f]; }   // getCovariance, :30-31
    sym_operator_sqrt(C2, D, S2);                                                     // :35-36
    for (int i = 0; i < D; ++i) for (int j = 0; j < D; ++j) { double a = 0; for (int k = 0; k < D; ++k) a += S2[i * D + k] * C1[k * D + j]; t[i * D + j] = a; }
    for (int i = 0; i < D; ++i) for (int j = 0; j < D; ++j) { double a = 0; for (int k = 0; k < D; ++k) a += t[i * D + k] * S2[k * D + j]; m[i * D + j] = a; }   // :37
    sym_operator_sqrt(m, D, S3);                                                      // :38-39
    double tr = 0;
    for (int i = 0; i < D; ++i) tr += C1[i * D + i] + C2[i * D + i] - 2 * S3[i * D + i];
    return std::fabs(d2 + tr);                                                        // :41, :62
}

// selectNode (:119-148): best-cost active node within the selection radius of the sample, else the nearest
// active one.  distance(sample, node): the node is the second argument.
inline int select_node(const double* nodes, const double* cost, const unsigned char* active, long n, const double* sample, int D,
                       double selection_radius, double* dist_out) {
    const int W = D + 2 * D * D;
    int best = -1, nearest = -1;
    double best_cost = HUGE_VAL, best_d = 0, nd = HUGE_VAL;
    for (long i = 0; i < n; ++i) {
        if (!active[i]) continue;
        const double d = belief_distance(sample, nodes + i * W, D);
        if (d <= selection_radius && (best < 0 || cost[i] < best_cost)) { best_cost = cost[i]; best = (int)i; best_d = d; }
        if (nearest < 0 || d < nd) { nd = d; nearest = (int)i; }
    }
    if (dist_out) *dist_out = best >= 0 ? best_d : (nearest >= 0 ? nd : HUGE_VAL);
    return best >= 0 ? best : nearest;
}

// nearest witness (:150-178): distance(witness, node), the query node is the second argument
inline int nearest_witness(const double* witnesses, long n, const double* node, int D, double* dist_out) {
    const int W = D + 2 * D * D;
    int closest = -1;
    double cd = HUGE_VAL;
    for (long i = 0; i < n; ++i) {
        const double d = belief_distance(witnesses + i * W, node, D);
        if (closest < 0 || d < cd) { cd = d; closest = (int)i; }
    }
    if (dist_out) *dist_out = cd;
    return closest;
}

// ---------------------------------------------------------------------------
// a4/a5/a6: state validity checkers
// ---------------------------------------------------------------------------
struct Counters { uint64_t n_steps = 0, n_halfplane = 0, n_obstacle = 0; };

struct ObstacleTable {
    std::vector<HalfPlanes> hp;       // A_list_, B_list_
    std::vector<Pt> centre;           // Obstacle::x_, y_
    std::vector<Ring> minkowski;      // inflated rings (ChiSquaredBoundarySVC obs_list_)
};
inline ObstacleTable build_obstacle_table(const std::vector<Obstacle>& obs, const RobotShape& robot) {
    ObstacleTable t;
    for (const Obstacle& o : obs) {
        Ring m = minkowski_sum(robot.bounding, o.ring);      // PCCBlackmoreSVC.cpp:20
        t.minkowski.push_back(m);
        t.hp.push_back(half_planes(m));                      // :21
        t.centre.push_back(Pt{o.x, o.y});
    }
    return t;
}

// HyperplaneCCValidityChecker_ (PCCBlackmoreSVC.cpp:63-75)
inline bool hyperplane_cc(const HalfPlanes& hp, double x, double y, const double* PX, double erf_inv_result, Counters* cnt) {
    for (int i = 0; i < 4; ++i) {
        if (cnt) cnt->n_halfplane++;
        const double a0 = hp.A[i][0], a1 = hp.A[i][1];
        const double t0 = a0 * PX[0] + a1 * PX[2];        // (A.row(i) * PX)(0)
        const double t1 = a0 * PX[1] + a1 * PX[3];        // (A.row(i) * PX)(1)
        const double tmp = t0 * a0 + t1 * a1;             // ... * A.row(i)^T   (:67)
        const double PV = std::sqrt(tmp);                 // :68
        const double b_bar = std::sqrt(2) * PV * erf_inv_result;   // :69
        if (x * a0 + y * a1 >= (hp.B[i] + b_bar)) return true;     // :70
    }
    return false;
}

enum SvcKind { SVC_BLACKMORE = 0, SVC_ADAPTIVE = 1, SVC_CHISQUARED = 2 };

struct Svc {
    int kind = SVC_BLACKMORE;
    int dim = 2;
    double low[4] = {0, 0, 0, 0}, high[4] = {0, 0, 0, 0};
    ObstacleTable tab;
    double p_collision = 0;       // 1 - accep_prob
    double erf_inv_result = 0;    // PCCBlackmoreSVC.cpp:10
    double sc = 0;                // chi2 quantile (ChiSquaredBoundarySVC.cpp:7)
    double max_rad = 0;           // robot bounding radius (ChiSquaredBoundarySVC.cpp:4)
};

inline Svc make_svc(int kind, int dim, const double* low, const double* high,
                    const std::vector<Obstacle>& obs, const RobotShape& robot, double accep_prob) {
    Svc s;
    s.kind = kind; s.dim = dim;
    for (int i = 0; i < dim; ++i) { s.low[i] = low[i]; s.high[i] = high[i]; }
    s.tab = build_obstacle_table(obs, robot);
    s.p_collision = 1 - accep_prob;                                           // PCCBlackmoreSVC.cpp:5
    const size_t n = obs.size();
    if (kind == SVC_BLACKMORE && n > 0)
        s.erf_inv_result = erf_inv(1 - 2 * (s.p_collision / (double)(int)n));   // :10
    if (kind == SVC_CHISQUARED) s.sc = chi2_quantile_2dof(accep_prob);        // ChiSquaredBoundarySVC.cpp:7
    s.max_rad = robot.bounding_rad;
    return s;
}

// max real part of the eigenvalues of a general 2x2 (Eigen::EigenSolver restated:
// ChiSquaredBoundarySVC.cpp:45-48)
inline double max_eig_real_2x2(const double* M) {
    const double tr = M[0] + M[3];
    const double det = M[0] * M[3] - M[1] * M[2];
    const double h = tr / 2;
    const double disc = h * h - det;
    if (disc < 0) return h;
    return h + std::sqrt(disc);
}

// bg::buffer(point, point_circle(10)) (ChiSquaredBoundarySVC.cpp:58-62):
// vertices centre + r*(cos a, sin a), a = 0, -2pi/10, ... (clockwise), closed.
inline Ring decagon(double cx, double cy, double r) {
    Ring out;
    const double two_pi = 2.0 * M_PI;
    const double diff = two_pi / 10.0;
    double a = 0;
    for (int i = 0; i < 10; ++i, a -= diff) out.push_back(Pt{cx + r * std::cos(a), cy + r * std::sin(a)});
    out.push_back(out.front());
    return out;
}

// !bg::disjoint for two convex closed rings: separating-axis test, touching
// counts as intersecting.
inline bool convex_disjoint(const Ring& P, const Ring& Q) {
    auto separated = [](const Ring& A, const Ring& B) {
        for (size_t i = 0; i + 1 < A.size(); ++i) {
            const double nx = A[i + 1].y - A[i].y, ny = A[i].x - A[i + 1].x;
            if (nx == 0 && ny == 0) continue;
            double amin = HUGE_VAL, amax = -HUGE_VAL, bmin = HUGE_VAL, bmax = -HUGE_VAL;
            for (const Pt& p : A) { const double d = nx * p.x + ny * p.y; amin = std::min(amin, d); amax = std::max(amax, d); }
            for (const Pt& p : B) { const double d = nx * p.x + ny * p.y; bmin = std::min(bmin, d); bmax = std::max(bmax, d); }
            if (amax < bmin || bmax < amin) return true;
        }
        return false;
    };
    return separated(P, Q) || separated(Q, P);
}

// bg::within(A, B) for convex rings: every vertex of A inside or on B.
inline bool convex_within(const Ring& A, const Ring& B) {
    // orientation of B
    double area2 = 0;
    for (size_t i = 0; i + 1 < B.size(); ++i) area2 += B[i].x * B[i + 1].y - B[i + 1].x * B[i].y;
    const double sgn = area2 >= 0 ? 1.0 : -1.0;
    for (const Pt& p : A)
        for (size_t i = 0; i + 1 < B.size(); ++i) {
            const double cr = (B[i + 1].x - B[i].x) * (p.y - B[i].y) - (B[i + 1].y - B[i].y) * (p.x - B[i].x);
            if (sgn * cr < 0) return false;
        }
    return true;
}

// isValid for the three SVCs.  `cov2` = (Sigma+Lambda).block<2,2>(0,0), row-major
// (PCCBlackmoreSVC.cpp:41: for the unicycle this is the (x, xdot) block, quirk 2).
inline bool svc_is_valid(const Svc& s, const double* values, const double* cov2, Counters* cnt = nullptr) {
    if (!satisfies_bounds(values, s.low, s.high, s.dim)) return false;          // :35
    const double x = values[0], y = values[1];
    const size_t M = s.tab.hp.size();
    if (s.kind == SVC_BLACKMORE) {                                               // PCCBlackmoreSVC.cpp:54-60
        for (size_t o = 0; o < M; ++o) {
            if (cnt) cnt->n_obstacle++;
            if (!hyperplane_cc(s.tab.hp[o], x, y, cov2, s.erf_inv_result, cnt)) return false;
        }
        return true;
    }
    if (s.kind == SVC_ADAPTIVE) {                                                // AdaptiveRiskBlackmoreSVC.cpp:24-92
        if (M == 0) return true;
        // createEtaList_ (:69-92); only eta_list[0] is used (quirk 4, :48)
        double sum = 0;
        for (size_t o = 0; o < M; ++o) {
            const double dx = x - s.tab.centre[o].x, dy = y - s.tab.centre[o].y;
            const double di = std::sqrt(dx * dx + dy * dy);
            sum += (1 / di);
        }
        const double alpha = 1 / sum;
        const double dx0 = x - s.tab.centre[0].x, dy0 = y - s.tab.centre[0].y;
        const double d0 = std::sqrt(dx0 * dx0 + dy0 * dy0);
        const double eta0 = (alpha / d0) * s.p_collision;
        const double c = erf_inv(1 - (2 * eta0));                                // :61 (same value for every plane)
        for (size_t o = 0; o < M; ++o) {
            if (cnt) cnt->n_obstacle++;
            const HalfPlanes& hp = s.tab.hp[o];
            bool safe = false;
            for (int i = 0; i < 4; ++i) {                                        // isSafe_ :55-67
                if (cnt) cnt->n_halfplane++;
                const double a0 = hp.A[i][0], a1 = hp.A[i][1];
                const double t0 = a0 * cov2[0] + a1 * cov2[2];
                const double t1 = a0 * cov2[1] + a1 * cov2[3];
                const double tmp = t0 * a0 + t1 * a1;
                const double Pv = std::sqrt(tmp);
                const double vbar = std::sqrt(2) * Pv * c;
                if (a0 * x + a1 * y - hp.B[i] >= vbar) { safe = true; break; }
            }
            if (!safe) return false;
        }
        return true;
    }
    // SVC_CHISQUARED: ChiSquaredBoundarySVC.cpp:19-69
    if (M == 0) return true;                                                     // :30-33
    const double max_lambda = max_eig_real_2x2(cov2);                            // :45-48
    const double boundary = (max_lambda * s.sc + s.max_rad);                     // :49 (variance, quirk 3)
    const Ring disk = decagon(x, y, boundary);
    for (size_t o = 0; o < M; ++o) {
        if (cnt) cnt->n_obstacle++;
        if (!convex_disjoint(disk, s.tab.minkowski[o])) return false;            // :64-66
    }
    return true;
}

// CentralizedChiSquaredBoundarySVC's agent-agent test (CentralizedChiSquaredBoundarySVC.cpp:70-99): decagons of radius
// lambda_max * sc + r around the two means must be disjoint.  NOTE the reference computes the second radius from
// `solver_a2.eigenvalues()` although it called `solver_a1.compute(sigma_a2)` (:79-82): solver_a2 is never computed, so the
// value is undefined there; the evident intent (lambda_max of sigma_a2) is what is restated here and in the kernel.
inline bool chisq_pair_disjoint(const double* mu_a, const double* cov_a, double rad_a, const double* mu_b, const double* cov_b, double rad_b, double sc) {
    const double ra = (max_eig_real_2x2(cov_a) * sc + rad_a), rb = (max_eig_real_2x2(cov_b) * sc + rad_b);
    return convex_disjoint(decagon(mu_a[0], mu_a[1], ra), decagon(mu_b[0], mu_b[1], rb));
}

template <int D>
inline void cov_block2(const Belief<D>& b, double* c) {   // (Sigma+Lambda).block<2,2>(0,0)
    c[0] = b.S[0] + b.L[0];         c[1] = b.S[1] + b.L[1];
    c[2] = b.S[D] + b.L[D];         c[3] = b.S[D + 1] + b.L[D + 1];
}

// ---------------------------------------------------------------------------
// OMPL glue: propagateWhileValid, propagate-n (interpolate)
// ---------------------------------------------------------------------------
struct Model2 {
    LinearParams p;
    void step(const Belief2& s, const double* u, double duration, Belief2& o) const { (void)duration; linear_propagate(s, u, p, o); }
};
struct Model4 {
    UnicycleParams p;
    void step(const Belief4& s, const double* u, double duration, Belief4& o) const { unicycle_propagate(s, u, duration, p, o); }
};

// oc::SpaceInformation::propagateWhileValid (OMPL 1.6.0, restated; called at
// src/Planners/ConstraintRespectingBSST.cpp:248).  Returns the number of leading
// valid steps; `result` = last valid state (the start state if step 1 is invalid).
template <int D, class Model>
inline int propagate_while_valid(const Model& m, const Svc& svc, double step_size, const Belief<D>& state,
                                 const double* control, int steps, Belief<D>& result, Counters* cnt = nullptr,
                                 Belief<D>* traj = nullptr) {
    if (steps == 0) { result = state; return 0; }
    Belief<D> cur = state, nxt;
    int r = steps;
    for (int i = 0; i < steps; ++i) {
        m.step(cur, control, step_size, nxt);
        if (cnt) cnt->n_steps++;
        double c2[4];
        cov_block2(nxt, c2);
        if (!svc_is_valid(svc, nxt.v, c2, cnt)) { r = i; break; }
        cur = nxt;
        if (traj) traj[i] = cur;
    }
    result = cur;
    return r;
}

// oc::SpaceInformation::propagate(state, control, steps, out, alloc) as used by
// PathControl::interpolate (no validity checks): out[0..steps-1].
template <int D, class Model>
inline void propagate_n(const Model& m, double step_size, const Belief<D>& state, const double* control, int steps, Belief<D>* out) {
    Belief<D> cur = state;
    for (int i = 0; i < steps; ++i) { m.step(cur, control, step_size, out[i]); cur = out[i]; }
}

// PathControl::interpolate (OMPL 1.6.0, restated): nodes[0] is the root state,
// segment i applies controls[i] for steps[i] steps starting from nodes[i].
// Output: one state per step: nodes[0], then for each segment its re-propagated
// intermediate states followed by the stored node state nodes[i+1].
template <int D, class Model>
inline std::vector<Belief<D>> interpolate_path(const Model& m, double step_size, const std::vector<Belief<D>>& nodes,
                                               const std::vector<std::vector<double>>& controls, const std::vector<int>& steps) {
    std::vector<Belief<D>> out;
    for (size_t i = 0; i < controls.size(); ++i) {
        out.push_back(nodes[i]);
        if (steps[i] <= 1) continue;
        std::vector<Belief<D>> tmp(steps[i]);
        propagate_n<D>(m, step_size, nodes[i], controls[i].data(), steps[i], tmp.data());
        for (int j = 0; j + 1 < steps[i]; ++j) out.push_back(tmp[j]);
    }
    out.push_back(nodes[controls.size()]);
    return out;
}

// ---------------------------------------------------------------------------
// a7-a10: plan validity checkers
// ---------------------------------------------------------------------------
struct Dist2 { double mu[2]; double cov[4]; };

// BeliefPVC::getDistribution_ (BeliefPVC.cpp:84-113): 2-D -> Sigma+Lambda;
// 4-D -> entries (0,0),(0,2),(2,0),(2,2) of Sigma+Lambda (quirk 2).
template <int D>
inline Dist2 get_distribution(const Belief<D>& b) {
    Dist2 d;
    d.mu[0] = b.v[0]; d.mu[1] = b.v[1];
    if (D == 4) {
        d.cov[0] = b.S[0] + b.L[0];   d.cov[1] = b.S[2] + b.L[2];
        d.cov[2] = b.S[8] + b.L[8];   d.cov[3] = b.S[10] + b.L[10];
    } else {
        for (int i = 0; i < 4; ++i) d.cov[i] = b.S[i] + b.L[i];
    }
    return d;
}

enum PvcKind { PVC_BLACKMORE = 0, PVC_BOUNDINGBOX = 1, PVC_CHISQUARED = 2, PVC_ADAPTIVE_BLACKMORE = 3, PVC_ADAPTIVE_BOUNDINGBOX = 4,
               PVC_BLACKMORE2 = 5, PVC_CDFGRID = 6 };   // Blackmore2PVC, CDFGridPVC(n)

struct Pvc {
    int kind = PVC_BLACKMORE;
    int k = 0;                         // number of robots
    std::vector<std::string> names;    // "Robot i"
    std::vector<double> radii;         // bounding radii
    std::vector<HalfPlanes> pair_hp;   // indexed [min(a,b)*k + max(a,b)]
    double p_coll_agnts = 0;           // 1 - p_safe_agnts   (BeliefPVC.cpp:7)
    double p_coll_dist = 0;            // p_coll_agnts / (k-1)  (MinkowskiSumBlackmorePVC.cpp:7-8)
    double c_pair = 0;                 // erf_inv(1 - 2 p_coll_dist)   (:178)
    double sc = 0;                     // chi2 quantile of the ctor arg (ChiSquaredBoundaryPVC.cpp:19)
    int grid_n = 2;                    // CDFGridPVC: disk_steps_ (CDFGridPVC.cpp:5, `--pvc CDFGrid-n`, demos/main.cpp:126-134)
    std::vector<int> map_order;        // robot indices in std::map<std::string,...> order (quirk 6)
};

inline Ring bounding_box_ring(double r1, double r2) {      // BoundingBoxBlackmorePVC.cpp:174-195
    const double r = r1 + r2;
    return Ring{{-r, -r}, {r, -r}, {r, r}, {-r, r}, {-r, -r}};
}

inline Pvc make_pvc(int kind, const std::vector<RobotShape>& robots, double p_safe_arg) {
    Pvc p;
    p.kind = kind;
    p.k = (int)robots.size();
    p.p_coll_agnts = 1 - p_safe_arg;
    const int norm = p.k - 1;
    p.p_coll_dist = norm > 0 ? p.p_coll_agnts / norm : p.p_coll_agnts;
    p.c_pair = erf_inv(1 - (2 * p.p_coll_dist));
    p.sc = chi2_quantile_2dof(p_safe_arg);
    p.pair_hp.resize((size_t)p.k * p.k);
    for (int i = 0; i < p.k; ++i) {
        p.names.push_back("Robot " + std::to_string(i));
        p.radii.push_back(robots[i].bounding_rad);
    }
    for (int a = 0; a < p.k; ++a)
        for (int b = a + 1; b < p.k; ++b) {
            Ring poly = (kind == PVC_BOUNDINGBOX || kind == PVC_ADAPTIVE_BOUNDINGBOX || kind == PVC_CDFGRID)   // CDFGridPVC.cpp:32: getBoundingBox_
                            ? bounding_box_ring(robots[a].bounding_rad, robots[b].bounding_rad)
                            : minkowski_sum(robots[a].bounding, robots[b].bounding);
            p.pair_hp[(size_t)a * p.k + b] = half_planes(poly);
        }
    p.map_order.resize(p.k);
    for (int i = 0; i < p.k; ++i) p.map_order[i] = i;
    std::sort(p.map_order.begin(), p.map_order.end(), [&](int a, int b) { return p.names[a] < p.names[b]; });
    return p;
}

// isSafe_ (MinkowskiSumBlackmorePVC.cpp:169-184; BoundingBoxBlackmorePVC.cpp:197-212;
// AdaptiveRiskBlackmorePVC.cpp:236-251 with c = erf_inv(1-2*eta))
inline bool pair_is_safe_halfplanes(const HalfPlanes& hp, const double* mu_ab, const double* S_ab, double c, Counters* cnt = nullptr) {
    for (int i = 0; i < 4; ++i) {
        if (cnt) cnt->n_halfplane++;
        const double a0 = hp.A[i][0], a1 = hp.A[i][1];
        const double t0 = a0 * S_ab[0] + a1 * S_ab[2];
        const double t1 = a0 * S_ab[1] + a1 * S_ab[3];
        const double tmp = t0 * a0 + t1 * a1;
        const double Pv = std::sqrt(tmp);
        const double vbar = std::sqrt(2) * Pv * c;
        if (a0 * mu_ab[0] + a1 * mu_ab[1] - hp.B[i] >= vbar) return true;
    }
    return false;
}

// ChiSquaredBoundaryPVC::isSafe_ (ChiSquaredBoundaryPVC.cpp:139-168)
inline bool pair_is_safe_chisq(const Dist2& a, double rad_a, const Dist2& b, double rad_b, double sc) {
    const double la = max_eig_real_2x2(a.cov), lb = max_eig_real_2x2(b.cov);
    const double dx = a.mu[0] - b.mu[0], dy = a.mu[1] - b.mu[1];
    const double dist = std::sqrt(dx * dx + dy * dy);
    const double boundary = (la * sc + rad_a) + (lb * sc + rad_b);
    return dist > boundary;
}

// Blackmore2PVC::isSafe_ / ImprovedHyperplaneCCValidityChecker (Blackmore2PVC.cpp:169-192): the smallest of the four
// half-plane collision probabilities 1/2 + 1/2 erf((B_i - a_i.mu) / sqrt(2 a_i^T Sigma a_i)) against p_coll_agnts_ (NOT
// p_coll_dist_, which the constructor computes and never uses).
inline bool pair_is_safe_blackmore2(const HalfPlanes& hp, const double* mu_ab, const double* S_ab, double p_coll_agnts) {
    double p_obs = std::numeric_limits<double>::max();
    for (int i = 0; i < 4; ++i) {
        const double a0 = hp.A[i][0], a1 = hp.A[i][1];
        const double t0 = a0 * S_ab[0] + a1 * S_ab[2];
        const double t1 = a0 * S_ab[1] + a1 * S_ab[3];
        const double pv2 = t0 * a0 + t1 * a1;
        const double PV = std::sqrt(2 * pv2);
        const double pc = 0.5 + 0.5 * std::erf((hp.B[i] - (mu_ab[0] * a0 + mu_ab[1] * a1)) / PV);
        if (pc < p_obs) p_obs = pc;
    }
    return !(p_obs > p_coll_agnts);
}

// Eigen::EigenSolver<Matrix2d> as CDFGridPVC uses it (CDFGridPVC.cpp:189-193): real parts of the eigenvalues and of the
// normalised eigenvectors of a general real 2x2.  Restated in closed form; Eigen's RealSchur may return the pair in the other
// order or an eigenvector with the other sign - the grid probability below is invariant under both (the standard normal is
// symmetric under axis swaps and reflections).  Diagonal input: (a, d) and the identity, as Eigen.  Complex pair (cannot
// happen for a sum of covariances): real part tr/2 twice, identity - not pinned.
inline void eig2_general(const double* M, double* lam, double* V) {
    const double a = M[0], b = M[1], c = M[2], d = M[3];
    V[0] = 1; V[1] = 0; V[2] = 0; V[3] = 1;
    if (b == 0 && c == 0) { lam[0] = a; lam[1] = d; return; }
    const double h = (a - d) / 2, disc = h * h + b * c, m = (a + d) / 2;
    if (!(disc >= 0)) { lam[0] = lam[1] = m; return; }
    const double sq = std::sqrt(disc);
    lam[0] = m + sq; lam[1] = m - sq;
    for (int j = 0; j < 2; ++j) {
        double vx, vy;
        if (std::fabs(b) >= std::fabs(c)) { vx = b; vy = lam[j] - a; } else { vx = lam[j] - d; vy = c; }
        const double n = std::sqrt(vx * vx + vy * vy);
        V[0 + j] = vx / n; V[2 + j] = vy / n;          // column j
    }
}

inline double cdf_rect(double x1, double y1, double x2, double y2) {           // CDFGridPVC.cpp:293-306
    auto p = [](double x, double y) { return (1.0 / 2.0) * (1 + std::erf(x / std::sqrt(2))) * (1.0 / 2.0) * (1 + std::erf(y / std::sqrt(2))); };
    const double p1 = p(x1, y1), p2 = p(x2, y1), p3 = p(x2, y2), p4 = p(x1, y2);
    return p3 - p2 - p4 + p1;
}

// CDFGridPVC::isSafe_ (CDFGridPVC.cpp:182-291): whiten the difference distribution, map the bounding box of r_a + r_b into
// the whitened frame, lay an n x n grid over ITS bounding box and sum the standard-normal mass of every cell that has a corner
// inside (or on) the mapped box; n == 2: the whole bounding box.  `eigen_vals` is a default-constructed Matrix2d in the
// reference, whose off-diagonal entries are never written: they are taken as 0 here (what a zero-initialised stack gives).
inline bool pair_is_safe_cdfgrid(const double* mu_a, const double* cov_a, const double* mu_b, const double* cov_b, double r_sum, int n, double p_coll_agnts,
                                 double* prob_out = nullptr) {
    const double T[2] = {mu_a[0] - mu_b[0], mu_a[1] - mu_b[1]};
    double S[4], lam[2], V[4];
    for (int i = 0; i < 4; ++i) S[i] = cov_a[i] + cov_b[i];
    eig2_general(S, lam, V);
    const double l0 = std::sqrt(lam[0]), l1 = std::sqrt(lam[1]);            // LLT of diag(lam): L = diag(sqrt)
    const double invdet = 1.0 / (l0 * l1 - 0.0 * 0.0);                      // L.inverse(): 2x2 adjugate * (1/det)
    const double Li00 = l1 * invdet, Li11 = l0 * invdet;
    const double R[4] = {V[0] * Li00 + V[1] * 0.0, V[0] * 0.0 + V[1] * Li11, V[2] * Li00 + V[3] * 0.0, V[2] * 0.0 + V[3] * Li11};   // eigen_vecs * L^-1
    const double vx[4] = {-r_sum, r_sum, r_sum, -r_sum}, vy[4] = {-r_sum, -r_sum, r_sum, r_sum};   // getBoundingBox_ after correct(): BL, BR, TR, TL
    double px[4], py[4];
    for (int k = 0; k < 4; ++k) {                                            // R^T * ((-V_mat).colwise() + T)
        const double ex = -vx[k] + T[0], ey = -vy[k] + T[1];
        px[k] = R[0] * ex + R[2] * ey;
        py[k] = R[1] * ex + R[3] * ey;
    }
    double x_low = px[0], x_high = px[0], y_low = py[0], y_high = py[0];
    for (int k = 1; k < 4; ++k) { x_low = std::min(x_low, px[k]); x_high = std::max(x_high, px[k]); y_low = std::min(y_low, py[k]); y_high = std::max(y_high, py[k]); }
    double prob = 0;
    if (n == 2) {
        prob += cdf_rect(x_low, y_low, x_high, y_high);                     // the box meets the polygon it bounds
    } else if (n > 2) {
        // orientation of the mapped quadrilateral (R^T may reflect): point inside-or-on <=> all edge cross products have its sign or are 0
        double area2 = 0;
        for (int k = 0; k < 4; ++k) { const int k1 = (k + 1) & 3; area2 += px[k] * py[k1] - px[k1] * py[k]; }
        const double sg = area2 >= 0 ? 1.0 : -1.0;
        auto inside = [&](double x, double y) {
            for (int k = 0; k < 4; ++k) {
                const int k1 = (k + 1) & 3;
                const double cr = (px[k1] - px[k]) * (y - py[k]) - (py[k1] - py[k]) * (x - px[k]);
                if (sg * cr < 0) return false;
            }
            return true;
        };
        const double dx = (x_high - x_low) / (n - 1), dy = (y_high - y_low) / (n - 1);       // linspace_ (:308-328)
        auto gx = [&](int i) { return i == n - 1 ? x_high : x_low + dx * i; };
        auto gy = [&](int i) { return i == n - 1 ? y_high : y_low + dy * i; };
        for (int ix = 0; ix != n - 1; ++ix)
            for (int iy = 0; iy != n - 1; ++iy) {
                const double x1 = gx(ix), x2 = gx(ix + 1), y1 = gy(iy), y2 = gy(iy + 1);
                if (inside(x1, y1) || inside(x2, y1) || inside(x2, y2) || inside(x1, y2)) prob += cdf_rect(x1, y1, x2, y2);
            }
    }
    if (prob_out) *prob_out = prob;
    return prob < p_coll_agnts;
}

// createEtaMap_ (AdaptiveRiskBlackmorePVC.cpp:179-206): `active` = robot indices
// present in states_map, in map order; returns eta for robot `b` seen from `a`.
inline double adaptive_eta(const Pvc& p, const std::vector<int>& active, const std::vector<Dist2>& dist, int a, int b) {
    double sum = 0;
    double d_b = 0;
    for (int j : active) {
        if (j == a) continue;
        const double dx = dist[a].mu[0] - dist[j].mu[0], dy = dist[a].mu[1] - dist[j].mu[1];
        const double di = std::sqrt(dx * dx + dy * dy);
        if (j == b) d_b = di;
        sum += (1 / di);
    }
    const double alpha = 1 / sum;
    return (alpha / d_b) * p.p_coll_agnts;
}

struct Conflict { int a, b, step; };

// checkForConflicts_ (MinkowskiSumBlackmorePVC.cpp:125-167 and the variants):
// `active` lists robot indices in std::map name order; first unsafe pair wins.
inline bool check_for_conflicts(const Pvc& p, const std::vector<int>& active, const std::vector<Dist2>& dist, int step, Conflict& out, Counters* cnt = nullptr) {
    for (size_t ia = 0; ia < active.size(); ++ia) {
        const int a = active[ia];
        for (size_t ib = ia + 1; ib < active.size(); ++ib) {
            const int b = active[ib];
            bool safe;
            if (p.kind == PVC_CHISQUARED) {
                safe = pair_is_safe_chisq(dist[a], p.radii[a], dist[b], p.radii[b], p.sc);
            } else if (p.kind == PVC_CDFGRID) {
                safe = pair_is_safe_cdfgrid(dist[a].mu, dist[a].cov, dist[b].mu, dist[b].cov, p.radii[std::min(a, b)] + p.radii[std::max(a, b)], p.grid_n, p.p_coll_agnts);
            } else if (p.kind == PVC_BLACKMORE2) {
                const double mu_ab[2] = {dist[a].mu[0] - dist[b].mu[0], dist[a].mu[1] - dist[b].mu[1]};
                double S_ab[4];
                for (int i = 0; i < 4; ++i) S_ab[i] = dist[a].cov[i] + dist[b].cov[i];
                safe = pair_is_safe_blackmore2(p.pair_hp[(size_t)std::min(a, b) * p.k + std::max(a, b)], mu_ab, S_ab, p.p_coll_agnts);
            } else {
                const double mu_ab[2] = {dist[a].mu[0] - dist[b].mu[0], dist[a].mu[1] - dist[b].mu[1]};
                double S_ab[4];
                for (int i = 0; i < 4; ++i) S_ab[i] = dist[a].cov[i] + dist[b].cov[i];
                const HalfPlanes& hp = p.pair_hp[(size_t)std::min(a, b) * p.k + std::max(a, b)];
                double c = p.c_pair;
                if (p.kind == PVC_ADAPTIVE_BLACKMORE || p.kind == PVC_ADAPTIVE_BOUNDINGBOX)
                    c = erf_inv(1 - (2 * adaptive_eta(p, active, dist, a, b)));
                safe = pair_is_safe_halfplanes(hp, mu_ab, S_ab, c, cnt);
            }
            if (!safe) { out = Conflict{a, b, step}; return true; }
        }
    }
    return false;
}

// validatePlan (MinkowskiSumBlackmorePVC.cpp:32-62 and the identical scaffolding
// of the other PVCs).  trajs[r] = interpolated states of robot r (one per dt).
template <int D>
inline std::vector<Conflict> validate_plan(const Pvc& p, const std::vector<std::vector<Belief<D>>>& trajs, Counters* cnt = nullptr) {
    int maxStates = 0;
    for (const auto& t : trajs) maxStates = std::max(maxStates, (int)t.size());
    auto dist_at = [&](int r, int step) {                        // getActiveRobots_ (BeliefPVC.cpp:42-82)
        const auto& t = trajs[r];
        return get_distribution<D>(step < (int)t.size() ? t[step] : t.back());
    };
    std::vector<Dist2> dist(p.k);
    for (int k = 0; k < maxStates; ++k) {
        for (int r = 0; r < p.k; ++r) dist[r] = dist_at(r, k);
        Conflict c;
        if (check_for_conflicts(p, p.map_order, dist, k, c, cnt)) {
            std::vector<Conflict> confs;
            int step = k;
            bool has = true;
            // the pair is re-inserted into a std::map keyed by name (quirk 6)
            std::vector<int> pair_active;
            for (int r : p.map_order) if (r == c.a || r == c.b) pair_active.push_back(r);
            while (has && step < maxStates) {                    // :52-57 (quirk 7)
                confs.push_back(c);
                step++;
                dist[c.a] = dist_at(c.a, step);
                dist[c.b] = dist_at(c.b, step);
                Conflict c2;
                has = check_for_conflicts(p, pair_active, dist, step, c2, cnt);
                if (has) c = c2;
            }
            return confs;
        }
    }
    return {};
}

// BeliefPVC::createConstraint (BeliefPVC.cpp:10-40)
template <int D>
struct BeliefConstraint {
    int constrained, constraining;
    std::vector<int> steps;              // timeStep_ of each conflict
    std::vector<double> times;           // timeStep_ * step_duration
    std::vector<Belief<D>> states;       // copies of the constraining robot's states (quirk 8)
};
template <int D>
inline BeliefConstraint<D> create_constraint(const std::vector<std::vector<Belief<D>>>& trajs, const std::vector<Conflict>& conflicts,
                                             int constrained_robot, double step_duration) {
    BeliefConstraint<D> c;
    c.constrained = constrained_robot;
    c.constraining = (constrained_robot == conflicts.front().a) ? conflicts.front().b : conflicts.front().a;
    for (const Conflict& cf : conflicts) {
        c.steps.push_back(cf.step);
        c.times.push_back(cf.step * step_duration);
        const auto& t = trajs[c.constraining];
        c.states.push_back(cf.step < (int)t.size() ? t[cf.step] : t.back());
    }
    return c;
}

// satisfiesConstraints (MinkowskiSumBlackmorePVC.cpp:64-104; Adaptive :61-107;
// ChiSquaredBoundaryPVC.cpp:54-87: the chi-squared isSafe_ of the planning robot's state
// against the stored state of the constraining robot, radii paired with their beliefs :81).
template <int D>
inline bool satisfies_constraints(const Pvc& p, const std::vector<Belief<D>>& path, double step_size,
                                  const std::vector<BeliefConstraint<D>>& constraints, Counters* cnt = nullptr) {
    double current_time = 0.0;
    for (size_t s = 0; s < path.size(); ++s) {
        for (const auto& c : constraints) {
            int idx = -1;
            for (size_t i = 0; i < c.times.size(); ++i)
                if (std::fabs(c.times[i] - current_time) < 1E-9) { idx = (int)i; break; }   // :75-76 (floating abs)
            if (idx < 0) continue;
            const int a = c.constrained, b = c.constraining;
            const Dist2 da = get_distribution<D>(path[s]);
            const Dist2 db = get_distribution<D>(c.states[idx]);
            if (p.kind == PVC_CHISQUARED) {
                if (!pair_is_safe_chisq(da, p.radii[a], db, p.radii[b], p.sc)) return false;
                continue;
            }
            if (p.kind == PVC_CDFGRID) {      // CDFGridPVC.cpp:72-109: constrained robot's belief first
                if (!pair_is_safe_cdfgrid(da.mu, da.cov, db.mu, db.cov, p.radii[std::min(a, b)] + p.radii[std::max(a, b)], p.grid_n, p.p_coll_agnts)) return false;
                continue;
            }
            if (p.kind == PVC_BLACKMORE2) {   // Blackmore2PVC.cpp:64-104
                const double m2[2] = {da.mu[0] - db.mu[0], da.mu[1] - db.mu[1]};
                double S2[4];
                for (int i = 0; i < 4; ++i) S2[i] = da.cov[i] + db.cov[i];
                if (!pair_is_safe_blackmore2(p.pair_hp[(size_t)std::min(a, b) * p.k + std::max(a, b)], m2, S2, p.p_coll_agnts)) return false;
                continue;
            }
            const double mu_ab[2] = {da.mu[0] - db.mu[0], da.mu[1] - db.mu[1]};
            double S_ab[4];
            for (int i = 0; i < 4; ++i) S_ab[i] = da.cov[i] + db.cov[i];
            const HalfPlanes& hp = p.pair_hp[(size_t)std::min(a, b) * p.k + std::max(a, b)];
            double cc = p.c_pair;
            if (p.kind == PVC_ADAPTIVE_BLACKMORE || p.kind == PVC_ADAPTIVE_BOUNDINGBOX) {
                // states_map holds only the two robots (:92-101): eta = (alpha/d)*p_coll_agnts
                const double dx = mu_ab[0], dy = mu_ab[1];
                const double di = std::sqrt(dx * dx + dy * dy);
                double sum = 0; sum += (1 / di);
                const double alpha = 1 / sum;
                cc = erf_inv(1 - (2 * ((alpha / di) * p.p_coll_agnts)));
            }
            if (!pair_is_safe_halfplanes(hp, mu_ab, S_ab, cc, cnt)) return false;
        }
        current_time += step_size;
    }
    return true;
}

// ---------------------------------------------------------------------------
// f1: ChanceConstrainedGoal::isSafe_ (src/Goals/BeliefSpaceGoals.cpp:90-147)
// ---------------------------------------------------------------------------
inline bool goal_is_satisfied(const double* values, const double* cov2, double gx, double gy, double threshold, double p_safe, double* distance) {
    const double dx = values[0] - gx, dy = values[1] - gy;
    if (distance) *distance = (dx * dx + dy * dy);
    const double max_lambda = max_eig_real_2x2(cov2);
    const double sc = chi2_quantile_2dof(p_safe);
    const double boundary = (max_lambda * sc + 0.0);
    const Ring disk = decagon(values[0], values[1], boundary);
    const Ring goal_disk = decagon(gx, gy, threshold);
    return convex_within(disk, goal_disk);
}

}  // namespace cck_oracle
