// cck_device.cuh — device-side building blocks of the belief propagate + chance-check path.
//
// Compiled with -fmad=false: every decision inequality and every propagated value is
// produced by the same sequence of IEEE-754 double +,-,*,/,sqrt as the reference's
// x86-64 (no-FMA) build, so verdicts and the linear model's values are bit-identical
// to a CPU evaluation of the reference formulas.
#pragma once
#include <cfloat>
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/cck_b200.h"

namespace cck {

// ---------------------------------------------------------------------------
// Obstacle table as staged into shared memory (one TMA bulk copy per CTA).
// Obstacles are grouped into CLASSES of bit-identical half-plane normals A[4][2]:
// sqrt(a^T P a) and a.mu depend on the belief and on A only, so they are hoisted
// out of the per-obstacle loop (4 sqrt per class instead of 4 per obstacle) while
// every comparison still sees exactly the reference's operands.
// ---------------------------------------------------------------------------
struct TabHeader {
    int32_t n_classes, M;
    int32_t off_classes, off_B, off_centres, off_rects;   // byte offsets from the table base
    int32_t total_bytes, pad;
};
struct ClassRec {
    double A[8];            // [4][2]
    int32_t first, count;   // range in the class-sorted B array (rows [first, first+count) are real)
    int32_t count_pad;      // count rounded up to a multiple of 4; padding rows hold B = -inf
    int32_t reserved;
};
static_assert(sizeof(TabHeader) == 32, "header");
static_assert(sizeof(ClassRec) == 80, "class record");

struct SvcDev {
    int32_t kind, dim;
    double low[4], high[4];
    double erf_inv_result, p_collision, sc, max_rad;
    double dec_cos[10], dec_sin[10];   // decagon directions cos/sin(-i*2pi/10), accumulated as Boost does
};

struct LinDev { double dt; double Q[4], R[4], Acl[4]; int diag, pad; double dg[6]; };   // diag: Q, R, A_cl are diagonal with |A_cl| <= 1 (lin_step_diag is usable); dg = {Q00, Q11, R00, R11, A00, A11} packed for three 16-byte constant loads
struct UniDev { double F[16], Acl[16], Q[16], R[16], gains[8], acc_lo, acc_hi, turn_lo, turn_hi; int sparse, pad; };   // sparse: F, A_cl_d, Q, R have the reference's exact-zero pattern (uni_cov_step_sparse is usable)

// ---------------------------------------------------------------------------
// PTX helpers: mbarrier + TMA 1-D bulk copy (cp.async.bulk -> SASS UBLKCP)
// ---------------------------------------------------------------------------
#ifndef CCK_HOSTSIM
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
#define CCK_OPAQUE4(a, b, c, d) asm volatile("" : "+d"(a), "+d"(b), "+d"(c), "+d"(d))
#else   // tests/hostsim (CPU execution of the kernels, test infrastructure): the copy is made at once, the barrier is a flag
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t) { *bar = 0; }
__device__ __forceinline__ void mbar_expect_tx(uint64_t*, uint32_t) {}
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) { memcpy(smem_dst, gmem_src, bytes); *bar = 1; }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t) { while (*(volatile uint64_t*)bar == 0) sim::yield(); }
#define CCK_OPAQUE4(a, b, c, d) ((void)0)
#endif

// Stage `bytes` (multiple of 16) of the obstacle table into shared memory.  Thread 0
// issues the bulk copy; callers overlap their own global loads and then call
// stage_wait().  The barrier is single-use (parity 0).
__device__ __forceinline__ void stage_begin(unsigned char* smem_tab, const unsigned char* gmem_tab, uint32_t bytes, uint64_t* bar) {
    if (threadIdx.x == 0) mbar_init(bar, 1);
    __syncthreads();
    if (threadIdx.x == 0 && bytes > 0) {
        mbar_expect_tx(bar, bytes);
        tma_bulk_g2s(smem_tab, gmem_tab, bytes, bar);
    }
}
__device__ __forceinline__ void stage_wait(uint32_t bytes, uint64_t* bar) {
    if (bytes > 0) mbar_wait(bar, 0);
}

// ---------------------------------------------------------------------------
// a1: linear model step (UncertainLinearSP.cpp:58-111), registers only
// ---------------------------------------------------------------------------
struct Lin { double mx, my, S[4], L[4]; };

__device__ __forceinline__ void mm2(const double* A, const double* B, double* C) {
    C[0] = A[0] * B[0] + A[1] * B[2];
    C[1] = A[0] * B[1] + A[1] * B[3];
    C[2] = A[2] * B[0] + A[3] * B[2];
    C[3] = A[2] * B[1] + A[3] * B[3];
}

// dux = dt*u0, duy = dt*u1 are step-invariant (the control is held), hoisted by the caller.
__device__ __forceinline__ void lin_step(const Lin& s, double dux, double duy, const LinDev& p, Lin& o) {
    o.mx = s.mx + dux;                                   // :92
    o.my = s.my + duy;                                   // :93
    double sp[4], S[4], Si[4], K[4], t[4], lp[4], ImK[4], KP[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) sp[i] = s.S[i] + p.Q[i]; // :98 (F = I)
#pragma unroll
    for (int i = 0; i < 4; ++i) S[i] = sp[i] + p.R[i];   // :100 (H = I)
    const double det = S[0] * S[3] - S[2] * S[1];
    const double invdet = 1.0 / det;
    Si[0] = S[3] * invdet; Si[2] = -S[2] * invdet; Si[1] = -S[1] * invdet; Si[3] = S[0] * invdet;
    mm2(sp, Si, K);                                      // :101
    mm2(p.Acl, s.L, t); mm2(t, p.Acl, lp);               // :102
    ImK[0] = 1.0 - K[0]; ImK[1] = 0.0 - K[1]; ImK[2] = 0.0 - K[2]; ImK[3] = 1.0 - K[3];
    mm2(ImK, sp, o.S);                                   // :103
    mm2(K, sp, KP);                                      // :104
#pragma unroll
    for (int i = 0; i < 4; ++i) o.L[i] = lp[i] + KP[i];
}

// The reference's constants are diagonal (Q = q I, R = r I, A_cl = (1 - K) I: UncertainLinearSP.h:40-55).  A skipped term
// is 0 * x or x + 0 with x finite: an exact zero, and adding an exact zero changes nothing - same values in the same
// accumulation order as lin_step (at most the SIGN of an exact zero differs: (-0) + (+0) = +0).  Preconditions, checked
// by the caller: every entry of s.S and s.L finite, |A_cl| <= 1 (so A_cl * L stays finite); anything else takes lin_step.
// 70 FP64 instructions instead of 90.
__device__ __forceinline__ void lin_step_diag(const Lin& s, double dux, double duy, const LinDev& p, Lin& o) {
    o.mx = s.mx + dux;
    o.my = s.my + duy;
    double sp[4], S[4], Si[4], K[4], lp[4], ImK[4], KP[4];
    const double q0 = p.dg[0], q3 = p.dg[1], r0 = p.dg[2], r3 = p.dg[3], a0 = p.dg[4], a3 = p.dg[5];
    sp[0] = s.S[0] + q0; sp[1] = s.S[1]; sp[2] = s.S[2]; sp[3] = s.S[3] + q3;
    S[0] = sp[0] + r0; S[1] = sp[1]; S[2] = sp[2]; S[3] = sp[3] + r3;
    const double det = S[0] * S[3] - S[2] * S[1];
    const double invdet = 1.0 / det;
    Si[0] = S[3] * invdet; Si[2] = -S[2] * invdet; Si[1] = -S[1] * invdet; Si[3] = S[0] * invdet;
    mm2(sp, Si, K);
    // (A_cl L) A_cl with A_cl = diag(a0, a3): entry (i, j) = (a_i * L_ij) * a_j
    lp[0] = (a0 * s.L[0]) * a0; lp[1] = (a0 * s.L[1]) * a3;
    lp[2] = (a3 * s.L[2]) * a0; lp[3] = (a3 * s.L[3]) * a3;
    // 0 - K_ij is a sign flip (a free operand modifier instead of a DADD) except for K_ij = +0, where 0 - (+0) = +0 but
    // -(+0) = -0: again only the sign of an exact zero can differ (see above)
    ImK[0] = 1.0 - K[0]; ImK[1] = -K[1]; ImK[2] = -K[2]; ImK[3] = 1.0 - K[3];
    mm2(ImK, sp, o.S);
    mm2(K, sp, KP);
#pragma unroll
    for (int i = 0; i < 4; ++i) o.L[i] = lp[i] + KP[i];
}

// covariance P = Sigma + Lambda with finite entries below 2^100 and a non-negative diagonal, by the high words (a sign
// bit or a NaN/inf exponent compares as a huge unsigned number).  True implies every entry of Sigma and Lambda is finite.
__device__ __forceinline__ bool cov_small(double P0, double P1, double P2, double P3) {
    const unsigned T = 0x46300000u;
    return ((unsigned)__double2hiint(P0) < T) & ((unsigned)__double2hiint(P3) < T) & ((unsigned)__double2hiint(fabs(P1) + fabs(P2)) < T);
}

// ---------------------------------------------------------------------------
// a2: unicycle model step (UncertainUnicycleSP.cpp:146-227)
// ---------------------------------------------------------------------------
struct Uni { double v[4], S[16], L[16]; };

__device__ __forceinline__ void mm4(const double* A, const double* B, double* C) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double acc = A[i * 4 + 0] * B[0 * 4 + j];
            acc = acc + A[i * 4 + 1] * B[1 * 4 + j];
            acc = acc + A[i * 4 + 2] * B[2 * 4 + j];
            acc = acc + A[i * 4 + 3] * B[3 * 4 + j];
            C[i * 4 + j] = acc;
        }
}
__device__ __forceinline__ void mm4_bt(const double* A, const double* B, double* C) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double acc = A[i * 4 + 0] * B[j * 4 + 0];
            acc = acc + A[i * 4 + 1] * B[j * 4 + 1];
            acc = acc + A[i * 4 + 2] * B[j * 4 + 2];
            acc = acc + A[i * 4 + 3] * B[j * 4 + 3];
            C[i * 4 + j] = acc;
        }
}
__device__ __forceinline__ void inv4(const double* m, double* inv) {
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
#pragma unroll
    for (int i = 0; i < 16; ++i) inv[i] = a[i] * invdet;
}

// The covariance recursion does not depend on the mean or the control, and within one
// held control cr = cos(yaw_ref), sr = sin(yaw_ref) are step-invariant (hoisted).
__device__ __forceinline__ void uni_cov_step(const double* Sg, const double* Lm, const UniDev& p, double* So, double* Lo) {
    double sp[16], t[16], S[16], Si[16], K[16], lp[16], ImK[16], KP[16];
    mm4(p.F, Sg, t); mm4(t, p.F, sp);                    // :213  F*Sigma*F (no transpose, quirk 2)
#pragma unroll
    for (int i = 0; i < 16; ++i) sp[i] = sp[i] + p.Q[i];
#pragma unroll
    for (int i = 0; i < 16; ++i) S[i] = sp[i] + p.R[i];  // :215
    inv4(S, Si);
    mm4(sp, Si, K);                                      // :216
    mm4(p.Acl, Lm, t); mm4_bt(t, p.Acl, lp);             // :217
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) ImK[i * 4 + j] = (i == j ? 1.0 : 0.0) - K[i * 4 + j];
    mm4(ImK, sp, So);                                    // :219
    mm4(K, sp, KP);                                      // :220
#pragma unroll
    for (int i = 0; i < 16; ++i) Lo[i] = lp[i] + KP[i];
}

__device__ __forceinline__ void uni_mean_step(const double* v, const double* u, double cr, double sr, double duration,
                                              const UniDev& p, double* o) {
    const double x = v[0], y = v[1], yaw = v[2], surge = v[3];
    double sin_y, cos_y;
    sincos(yaw, &sin_y, &cos_y);
    const double u0 = p.gains[0] * (u[0] - x) + p.gains[1] * (u[3] * cr - surge * cos_y);    // :173
    const double u1 = p.gains[6] * (u[1] - y) + p.gains[7] * (u[3] * sr - surge * sin_y);    // :174
    double ub0 = cos_y * u0 + sin_y * u1;                                                    // :179
    double ub1 = (-sin_y * u0 + cos_y * u1) / surge;                                         // :180
    if (ub0 < p.acc_lo) ub0 = p.acc_lo;                                                      // :132-135 (NaN passes)
    if (ub0 > p.acc_hi) ub0 = p.acc_hi;
    if (ub1 < p.turn_lo) ub1 = p.turn_lo;
    if (ub1 > p.turn_hi) ub1 = p.turn_hi;
    const double surge_final = surge + duration * ub0;                                       // :190
    o[0] = x + duration * cos_y * surge_final;                                               // :202
    o[1] = y + duration * sin_y * surge_final;                                               // :203
    const double ang = yaw + duration * ub1;
    const double kk = floor((M_PI - ang) / (2 * M_PI));                                       // :139
    o[2] = ang + (2 * kk * M_PI);                                                            // :140
    o[3] = surge_final;
}

// ---------------------------------------------------------------------------
// a4/a5/a6: validity of one belief against the staged obstacle table
// ---------------------------------------------------------------------------
__device__ __forceinline__ bool bounds_ok(const SvcDev& sp, const double* v) {
    const double eps = DBL_EPSILON;
    bool ok = true;
#pragma unroll
    for (int d = 0; d < 4; ++d)
        if (d < sp.dim) ok = ok && !(v[d] - eps > sp.high[d] || v[d] + eps < sp.low[d]);
    return ok;
}

__host__ __device__ __forceinline__ double max_eig_real_2x2(double m0, double m1, double m2, double m3) {
    const double tr = m0 + m3;
    const double det = m0 * m3 - m1 * m2;
    const double h = tr / 2;
    const double disc = h * h - det;
    if (disc < 0) return h;
    return h + sqrt(disc);
}

// Blackmore / Adaptive half-plane test over all obstacles.  ADAPT=false:
//   safe_o = OR_i  a_i.mu >= B_oi + sqrt2*sqrt(a_i^T P a_i)*c       (PCCBlackmoreSVC.cpp:66-71)
// ADAPT=true:
//   safe_o = OR_i  a_i.mu - B_oi >= sqrt2*sqrt(a_i^T P a_i)*c(mu)   (AdaptiveRiskBlackmoreSVC.cpp:58-63)
template <bool ADAPT>
__device__ __forceinline__ bool halfplane_check(const unsigned char* tab, double x, double y, double P0, double P1,
                                                double P2, double P3, double c) {
    const TabHeader* h = reinterpret_cast<const TabHeader*>(tab);
    const ClassRec* cls = reinterpret_cast<const ClassRec*>(tab + h->off_classes);
    const double2* Bt = reinterpret_cast<const double2*>(tab + h->off_B);
    const int nc = h->n_classes;
    bool valid = true;
    for (int k = 0; k < nc && valid; ++k) {
        double lhs[4], bb[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const double a0 = cls[k].A[2 * i], a1 = cls[k].A[2 * i + 1];
            const double t0 = a0 * P0 + a1 * P2;
            const double t1 = a0 * P1 + a1 * P3;
            const double tmp = t0 * a0 + t1 * a1;
            const double PV = sqrt(tmp);
            bb[i] = 1.4142135623730951 * PV * c;
            lhs[i] = x * a0 + y * a1;
        }
        const int first = cls[k].first, cnt = cls[k].count;
        int o = 0;
        // two obstacles per iteration for ILP; the table rows are warp-uniform (LDS broadcast)
        for (; o + 2 <= cnt; o += 2) {
            const double2 b01 = Bt[(first + o) * 2], b23 = Bt[(first + o) * 2 + 1];
            const double2 c01 = Bt[(first + o) * 2 + 2], c23 = Bt[(first + o) * 2 + 3];
            bool s0, s1;
            if (ADAPT) {
                s0 = (lhs[0] - b01.x >= bb[0]) | (lhs[1] - b01.y >= bb[1]) | (lhs[2] - b23.x >= bb[2]) | (lhs[3] - b23.y >= bb[3]);
                s1 = (lhs[0] - c01.x >= bb[0]) | (lhs[1] - c01.y >= bb[1]) | (lhs[2] - c23.x >= bb[2]) | (lhs[3] - c23.y >= bb[3]);
            } else {
                s0 = (lhs[0] >= b01.x + bb[0]) | (lhs[1] >= b01.y + bb[1]) | (lhs[2] >= b23.x + bb[2]) | (lhs[3] >= b23.y + bb[3]);
                s1 = (lhs[0] >= c01.x + bb[0]) | (lhs[1] >= c01.y + bb[1]) | (lhs[2] >= c23.x + bb[2]) | (lhs[3] >= c23.y + bb[3]);
            }
            valid = valid & s0 & s1;
            if (!valid) break;
        }
        if (valid && o < cnt) {
            const double2 b01 = Bt[(first + o) * 2], b23 = Bt[(first + o) * 2 + 1];
            bool s0;
            if (ADAPT) s0 = (lhs[0] - b01.x >= bb[0]) | (lhs[1] - b01.y >= bb[1]) | (lhs[2] - b23.x >= bb[2]) | (lhs[3] - b23.y >= bb[3]);
            else s0 = (lhs[0] >= b01.x + bb[0]) | (lhs[1] >= b01.y + bb[1]) | (lhs[2] >= b23.x + bb[2]) | (lhs[3] >= b23.y + bb[3]);
            valid = valid & s0;
        }
    }
    return valid;
}

// Adaptive risk allocation: only eta_list[0] is used (quirk 4).
__device__ __forceinline__ double adaptive_c(const SvcDev& sp, const unsigned char* tab, double x, double y) {
    const TabHeader* h = reinterpret_cast<const TabHeader*>(tab);
    const double2* ctr = reinterpret_cast<const double2*>(tab + h->off_centres);
    double sum = 0;
    for (int o = 0; o < h->M; ++o) {
        const double dx = x - ctr[o].x, dy = y - ctr[o].y;
        const double di = sqrt(dx * dx + dy * dy);
        sum += (1 / di);
    }
    const double alpha = 1 / sum;
    const double dx0 = x - ctr[0].x, dy0 = y - ctr[0].y;
    const double d0 = sqrt(dx0 * dx0 + dy0 * dy0);
    const double eta0 = (alpha / d0) * sp.p_collision;
    return erfinv(1 - (2 * eta0));
}

// ChiSquaredBoundarySVC: decagon of circumradius lambda_max*sc + r_robot vs every inflated
// rectangle, separating-axis test (touching = not disjoint).
// The separating-axis test of decagon (px, py: 10 vertices + the closing one) vs rectangle, written exactly as the oracle's
// convex_disjoint: running minima / maxima that start at +-inf and are replaced only when a comparison is TRUE, so NaN
// projections are never taken.  A decagon with NaN vertices (NaN / inf covariance) is therefore "separated" from every
// rectangle - the state is valid - whereas fmin / fmax of the fast path below would carry the NaN into comparisons that are
// all false (found by tests/hostsim/fuzz_rollout.py --generic).  Used for non-finite or huge operands only.
__device__ __noinline__ bool chisq_disjoint_literal(const double* px, const double* py, double xlo, double ylo, double xhi, double yhi) {
    const double rx[5] = {xlo, xhi, xhi, xlo, xlo}, ry[5] = {ylo, ylo, yhi, yhi, ylo};
    for (int pass = 0; pass < 2; ++pass) {
        const double* ex = pass ? rx : px; const double* ey = pass ? ry : py;
        const int ne = pass ? 4 : 10;
        for (int i = 0; i < ne; ++i) {
            const double nx = ey[i + 1] - ey[i], ny = ex[i] - ex[i + 1];
            if (nx == 0 && ny == 0) continue;
            double amin = HUGE_VAL, amax = -HUGE_VAL, bmin = HUGE_VAL, bmax = -HUGE_VAL;
            for (int j = 0; j < 11; ++j) { const double d = nx * px[j] + ny * py[j]; amin = d < amin ? d : amin; amax = amax < d ? d : amax; }
            for (int j = 0; j < 5; ++j) { const double d = nx * rx[j] + ny * ry[j]; bmin = d < bmin ? d : bmin; bmax = bmax < d ? d : bmax; }
            if (amax < bmin || bmax < amin) return true;
        }
    }
    return false;
}

__device__ __forceinline__ bool chisq_check(const SvcDev& sp, const unsigned char* tab, double x, double y, double P0,
                                            double P1, double P2, double P3) {
    const TabHeader* h = reinterpret_cast<const TabHeader*>(tab);
    const double4* rc = reinterpret_cast<const double4*>(tab + h->off_rects);
    const double max_lambda = max_eig_real_2x2(P0, P1, P2, P3);
    const double boundary = (max_lambda * sp.sc + sp.max_rad);
    double px[11], py[11];
#pragma unroll
    for (int i = 0; i < 10; ++i) { px[i] = x + boundary * sp.dec_cos[i]; py[i] = y + boundary * sp.dec_sin[i]; }
    px[10] = px[0]; py[10] = py[0];
    if (!(fabs(x) < 1e150 && fabs(y) < 1e150 && fabs(boundary) < 1e150)) {     // NaN compares false: the literal test decides
        for (int o = 0; o < h->M; ++o) {
            const double4 r = rc[o];
            if (!chisq_disjoint_literal(px, py, r.x, r.y, r.z, r.w)) return false;
        }
        return true;
    }
    double dxmin = px[0], dxmax = px[0], dymin = py[0], dymax = py[0];
#pragma unroll
    for (int i = 1; i < 10; ++i) {
        dxmin = fmin(dxmin, px[i]); dxmax = fmax(dxmax, px[i]);
        dymin = fmin(dymin, py[i]); dymax = fmax(dymax, py[i]);
    }
    for (int o = 0; o < h->M; ++o) {
        const double4 r = rc[o];   // xlo, ylo, xhi, yhi
        // rectangle edge normals are (0,-w),(w,0),(0,w),(-w,0): projections reduce to x / y extents
        // scaled by +-w; the strict comparisons below are the same predicates.
        const double wx = r.z - r.x, wy = r.w - r.y;
        bool sep = false;
        // normals (0, x1-x2) of the bottom/top edges and (y2-y1, 0) of the right/left edges
        {
            const double n = r.x - r.z;   // bottom edge: ny = x1 - x2 = -(wx)
            const double a_lo = fmin(n * r.y, n * r.w), a_hi = fmax(n * r.y, n * r.w);
            const double b_lo = fmin(n * dymin, n * dymax), b_hi = fmax(n * dymin, n * dymax);
            sep = sep || (a_hi < b_lo) || (b_hi < a_lo);
        }
        {
            const double n = r.w - r.y;   // right edge: nx = y2 - y1 = wy
            const double a_lo = fmin(n * r.x, n * r.z), a_hi = fmax(n * r.x, n * r.z);
            const double b_lo = fmin(n * dxmin, n * dxmax), b_hi = fmax(n * dxmin, n * dxmax);
            sep = sep || (a_hi < b_lo) || (b_hi < a_lo);
        }
        (void)wx; (void)wy;
        if (!sep) {
            // decagon edge normals
            for (int i = 0; i < 10 && !sep; ++i) {
                const double nx = py[i + 1] - py[i], ny = px[i] - px[i + 1];
                if (nx == 0 && ny == 0) continue;
                double amin = nx * px[0] + ny * py[0], amax = amin;
#pragma unroll
                for (int j = 1; j < 10; ++j) { const double d = nx * px[j] + ny * py[j]; amin = fmin(amin, d); amax = fmax(amax, d); }
                const double q0 = nx * r.x + ny * r.y, q1 = nx * r.z + ny * r.y, q2 = nx * r.z + ny * r.w, q3 = nx * r.x + ny * r.w;
                const double bmin = fmin(fmin(q0, q1), fmin(q2, q3)), bmax = fmax(fmax(q0, q1), fmax(q2, q3));
                sep = (amax < bmin) || (bmax < amin);
            }
        }
        if (!sep) return false;
    }
    return true;
}

// isValid: bounds + obstacle test.  P = (Sigma+Lambda).block<2,2>(0,0).
template <int KIND>
__device__ __forceinline__ bool svc_valid(const SvcDev& sp, const unsigned char* tab, const double* v, double P0, double P1,
                                          double P2, double P3) {
    if (!bounds_ok(sp, v)) return false;
    const TabHeader* h = reinterpret_cast<const TabHeader*>(tab);
    if (h->M == 0) return true;
    if (KIND == CCK_SVC_BLACKMORE) return halfplane_check<false>(tab, v[0], v[1], P0, P1, P2, P3, sp.erf_inv_result);
    if (KIND == CCK_SVC_ADAPTIVE) return halfplane_check<true>(tab, v[0], v[1], P0, P1, P2, P3, adaptive_c(sp, tab, v[0], v[1]));
    return chisq_check(sp, tab, v[0], v[1], P0, P1, P2, P3);
}

}  // namespace cck
