// cck_rollout_grid.cuh — third generation of the fused propagateWhileValid kernel for the default
// configuration (2DUncertainLinear + PCCBlackmoreSVC).
//
// ncu on the second generation (profiles/r1c_*) showed the FP64 pipe at ~55 % and the ALU pipe at
// ~75 %: the fp32 threshold computation (4 DSQRT per belief-step) cost more FP64 issue than the
// Kalman recursion itself, and the 32 group rows per belief-step saturated the ALU pipe (FSETP).
// This version keeps FP64 for what must be bit-exact (the recursion, the bounds test, the exact
// decision on unproven obstacles) and moves everything else to a CONSERVATIVE fp32 broad phase:
//
//  * every half-plane whose normal is axis-aligned, (±w, 0) or (0, ±w), says "safe if x - ex >= xhi",
//    "x + ex <= xlo", … in distance units, with ex = sqrt2*c*sqrt(P00), ey = sqrt2*c*sqrt(P11).  Each
//    obstacle therefore owns an fp32 box (xlo, ylo, xhi, yhi), rounded OUTWARD on the host; a belief
//    whose fp32 interval [x -+ ex] x [y -+ ey] (widened by an explicit rounding margin) misses the box
//    is PROVEN safe for the reference's fp64 inequality.  Planes that are not axis-aligned simply do
//    not contribute a side (side = +-inf): the obstacle is then decided by the exact path.
//  * the boxes' lower-left corners are binned into a 64-column grid in shared memory, stored as a dense row-major CSR
//    (u16 start per cell, the box arrays sorted in the same order): a belief-step looks at the 3-4 grid rows its interval
//    overlaps and, per row, at ONE contiguous range of table rows (usually empty) - two nested loops, no bitmap
//    arithmetic.  The cell function is the same monotone fp32 expression on host and device, so no margin is needed
//    for the binning.
//  * whatever the box test cannot prove (a real collision, the ~1e-5-wide fp32 band, non-finite
//    operands) is decided by the reference's exact fp64 expression on the obstacle's own normals
//    (bm_exact_obstacle).  Verdicts and first-violation steps stay bit-identical to the oracle.
//  * lanes are refilled from a CTA-wide work counter when their rollout ends - in batches of idle lanes, result stores
//    included - and the previous (last valid) state is stashed in shared memory instead of a second register copy.
//  * the common outcome of a step (bounds proven in fp32, covariance finite, window misses the grid, no obstacle without
//    a box) leaves grid_check through ONE branch; the Kalman step skips the exact zeros of the reference's diagonal
//    Q, R, A_cl (lin_step_diag) while the lane's covariance is known finite.  171 instructions per warp-step, 75 of
//    them FP64 at the end of round 1 (profiles/r1o_*).
// Round 2 (profiles/r2c_*):
//  * a CLEARANCE FIELD is the first test of a step: 128 x 128 cells over the state bounds, per cell the quantised Chebyshev
//    clearances to the nearest and second-nearest obstacle box (lower bounds) and the nearest box's row.  One fp32 compare
//    proves the step against ALL obstacles, a second one leaves a single obstacle to decide; only then the cell window.
//  * a WARP-UNIFORM fast section: while every lane has work and finite covariance, the warp runs min(steps left) steps of
//    nothing but the Kalman step, the field lookup and that compare, with one vote per step; the first unproven step falls
//    through to the general path.  120 executed instructions per warp-step, 69 of them FP64 (34 DMUL, 30 DADD, 5 DFMA);
//    statically 124 / 69 (tools/sass_loop_mix.py, tests/test_sass_static.py).
//  * the refill is warp-aggregated: a ballot counts the lanes that want work, ONE atomicAdd per warp hands out consecutive
//    beliefs.
#pragma once
#include "cck_kernels.cuh"

namespace cck {

// exact reference test of one obstacle (rare path; PCCBlackmoreSVC.cpp:63-75)
__device__ __noinline__ bool bm_exact_obstacle(const double* A8, const double* B, double x, double y, double P0, double P1,
                                               double P2, double P3, double c) {
    bool safe = false;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const double a0 = A8[2 * i], a1 = A8[2 * i + 1];
        const double t0 = a0 * P0 + a1 * P2;
        const double t1 = a0 * P1 + a1 * P3;
        const double tmp = t0 * a0 + t1 * a1;
        const double PV = sqrt(tmp);
        const double bb = 1.4142135623730951 * PV * c;
        safe = safe | (x * a0 + y * a1 >= B[i] + bb);
    }
    return safe;
}

struct GHeader {                 // 80 bytes, then dstart[R*64+1] (u16, dense row-major CSR over the cells), outer boxes[M], inner boxes[M] (float4)
    int32_t total_bytes, M, n_grid, n_cells;
    int32_t R, off_occ, off_rowbase, off_cstart;
    int32_t off_boxes, pad0;
    float gx0, gy0, icx, icy;    // cell = floor((v - g0) * ic), clamped by the caller
    float SX, SY;                // >= max box width / height of the gridded obstacles
    float k;                     // >= sqrt(2) * erf_inv_result
    float Rm1;                   // (float)(R - 1)
    int32_t pad1[2];
};
static_assert(sizeof(GHeader) == 80, "grid header");

// header scalars as kernel constants (constant-bank operands instead of ~10 LDS per belief-step) + the state
// bounds rounded INWARD to fp32 (pre-check of satisfiesBounds)
struct GridConst {
    float k, SX, SY, gx0, gy0, icx, icy, Rm1;
    int32_t M, n_grid, off_occ, off_rowbase, off_cstart, off_boxes;
    float blx, bhx, bly, bhy;
    int32_t off_classes;        // byte offset of the ClassRec array inside the (global) class table
    float Rf;                   // (float)R = Rm1 + 1
    const float4* boxes_global; // non-null: the box arrays are too large for shared memory and are read from global (L1/L2)
    // clearance field (first test of a belief-step): CCK_DF_N x CCK_DF_N cells over the fp32 state bounds, one 32-bit word per
    // cell: bits 0-7 floor(16 * Chebyshev clearance of the cell to the NEAREST obstacle box), bits 8-15 the same for the
    // SECOND nearest box, bits 16-31 the table row of the nearest box.  Global memory (64 KB, L1/L2-resident).
    const uint32_t* df;         // null: no field (no obstacles, an obstacle without a box, unusable bounds)
    float df_icx, df_icy;       // cells per unit, shrunk so that a mean strictly inside the bounds lands in [0, CCK_DF_N)
    float df_k2;                // >= 256 * 2 c^2 (1 + 2^-18): variance -> (16 * sqrt2 c sqrt(P))^2, rounded up
};
#ifndef CCK_DF_N
#define CCK_DF_N 128
#endif
#define CCK_DF_SCALE 16.0

#define CCK_GRID_COLS 64
#define CCK_GRID_ROWS 64

// The binning function.  Monotone non-decreasing in v (IEEE RN subtraction of a constant, RN
// multiplication by a positive constant, floor), evaluated identically on host (table build) and
// device (query): an anchor inside a query interval lands inside the interval's cell range.
__host__ __device__ __forceinline__ float grid_cellf(float v, float g0, float ic) {
#ifdef __CUDA_ARCH__
    return floorf(__fmul_rn(__fsub_rn(v, g0), ic));
#else
    volatile float d = v - g0;
    volatile float t = d * ic;
    return floorf(t);
#endif
}

// MUFU.SQRT: max relative error 2^-23 (PTX ISA, sqrt.approx.f32), denormal inputs flushed to zero
__device__ __forceinline__ float sqrt_approx(float v) {
#ifndef CCK_HOSTSIM
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
#else
    return sqrtf(v);   // tests/hostsim: within the error bound the callers already allow for
#endif
}

// One obstacle the broad phase could not skip: outer box (proves safe), inner box (proves unsafe:
// every plane fails with margin; +-inf for obstacles that are not one-plane-per-side boxes), else the
// reference's fp64 expression.
template <class Exact>
__device__ __forceinline__ bool grid_obstacle(const GridConst& g, const unsigned char* gt, int j, float xm, float xp, float ym, float yp, float mx,
                                              float my, const Exact& exact) {
    // the box arrays live in shared memory behind the grid, or in global memory when they are too large for it; the
    // choice is made here, on the rare path, not once per belief-step
    const float4* boxes = g.boxes_global ? g.boxes_global : reinterpret_cast<const float4*>(gt + g.off_boxes);
    const float4 b = boxes[j];
    if ((xm >= b.z) | (xp <= b.x) | (ym >= b.w) | (yp <= b.y)) return true;
    const float4 q = boxes[g.M + j];
    const float mx2 = mx + mx, my2 = my + my;
    const float xmi = xm + mx2, xpi = xp - mx2, ymi = ym + my2, ypi = yp - my2;
    if ((xmi < q.z) & (xpi > q.x) & (ymi < q.w) & (ypi > q.y)) return false;
    return exact(j);
}

// "the belief is safe w.r.t. every obstacle" (PCCBlackmoreSVC.cpp:54-58), bit-exact.
// SPLIT: the covariance arrives as its two summands (P = Sa + Sb, the reference's Sigma + Lambda).  The sums are formed
// here for the fp32 broad phase and formed AGAIN (same operands, same result) where the exact path needs them, behind an
// opaque asm so that the compiler cannot keep the first copies alive: 8 registers less across the hot loop.
// The caller's fp32 bounds pre-check (`pre`, which must imply a FINITE mean: 0 * NaN poisons every plane of the
// reference's expression) and its exact fp64 bounds test (`exact_bounds`, evaluated only when pre is false) are folded
// in, so that the common outcome - bounds proven, operands finite, no gridded obstacle near, no obstacle without a box -
// leaves through ONE branch (ncu: 21 of 215 instructions per warp-step were BRA/BSSY/BSYNC of five separate rare-path
// tests).  `enabled` = there is an obstacle table at all.
// Clearance field lookup for a mean the caller has PROVEN strictly inside the fp32 bounds (`pre`): the word of its cell, or 0
// (clearances 0: nothing provable) when there is no field / the mean is not proven inside.  Issued right after the mean
// update so that the load overlaps the covariance recursion.
__device__ __forceinline__ uint32_t df_lookup(const GridConst& g, bool pre, float xf, float yf) {
    if (!g.df || !pre) return 0u;
    const int cx = __float2int_rz(__fmul_rn(__fsub_rn(xf, g.blx), g.df_icx));
    const int cy = __float2int_rz(__fmul_rn(__fsub_rn(yf, g.bly), g.df_icy));
    return __ldg(g.df + (cy * CCK_DF_N + cx));
}

// `dfw` (optional): df_lookup's word for this mean.  With t = (16 sqrt2 c sqrt(max(P00, P11)))^2:
//   t < q1^2: the belief's interval stays clear of EVERY obstacle box by Chebyshev distance - proven safe with one compare,
//             before any window arithmetic;
//   t < q2^2: it stays clear of every box except possibly the nearest one - that single obstacle is decided by its box
//             tests / the exact expression, without the cell window and its loops.
template <bool SPLIT = false, class ExactBounds>
__device__ __forceinline__ bool grid_check_b(const GridConst& g, const unsigned char* gt, const unsigned char* ctab, const double* exB,
                                             const int32_t* exCls, double c, bool pre, const ExactBounds& exact_bounds, bool enabled, double x, double y,
                                             double Sa0, double Sa1, double Sa2, double Sa3, double Sb0 = 0, double Sb1 = 0, double Sb2 = 0,
                                             double Sb3 = 0, bool* cov_small_out = nullptr, uint32_t dfw = 0u) {
    const GridConst* h = &g;
    const int M = h->M;
    const double P0 = SPLIT ? Sa0 + Sb0 : Sa0, P1 = SPLIT ? Sa1 + Sb1 : Sa1, P2 = SPLIT ? Sa2 + Sb2 : Sa2, P3 = SPLIT ? Sa3 + Sb3 : Sa3;
    const float p0 = __double2float_rn(P0), p3 = __double2float_rn(P3);
    // covariance entries finite, non-negative on the diagonal and below 2^100, by their high words (sign bit or NaN
    // exponent compare as huge unsigned numbers)
    const bool cov_ok = cov_small(P0, P1, P2, P3);
    if (cov_small_out) *cov_small_out = cov_ok;
    // a non-zero clearance implies `pre` (df_lookup): bounds proven, mean finite.  NaN variances compare false.
    const float tq = __fmul_rn(fmaxf(p0, p3), h->df_k2);
    const float q1 = (float)(dfw & 255u);
    if (cov_ok & (tq < q1 * q1)) return true;
    const float xf = __double2float_rn(x), yf = __double2float_rn(y);
    // the reference's fp64 expression on obstacle j's own normals
    auto exact = [&](int j) -> bool {
        const ClassRec* cls = reinterpret_cast<const ClassRec*>(ctab + h->off_classes);   // a kernel constant: no header load
        if (SPLIT) {
            double s0 = Sa0, s1 = Sa1, s2 = Sa2, s3 = Sa3;
            CCK_OPAQUE4(s0, s1, s2, s3);
            return bm_exact_obstacle(cls[exCls[j]].A, exB + (size_t)j * 4, x, y, s0 + Sb0, s1 + Sb1, s2 + Sb2, s3 + Sb3, c);
        }
        return bm_exact_obstacle(cls[exCls[j]].A, exB + (size_t)j * 4, x, y, P0, P1, P2, P3, c);
    };
    const float ex = h->k * sqrt_approx(p0), ey = h->k * sqrt_approx(p3);
    // margin: 2^-18 relative covers the fp64->fp32 conversions, the fp32 sqrt/mul/add roundings (<2^-20)
    // and the fp64 roundings of the reference's expression (2^-50); 1e-17 absolute covers denormal inputs
    // (flushed by sqrt.approx.ftz: ex_true <= k * sqrt(1.2e-38) < 1e-18)
    const float mx = __fmaf_rn(fabsf(xf) + ex, 3.814697265625e-06f, 1e-17f);
    const float my = __fmaf_rn(fabsf(yf) + ey, 3.814697265625e-06f, 1e-17f);
    const float wx = ex + mx, wy = ey + my;
    const float xm = xf - wx, xp = xf + wx, ym = yf - wy, yp = yf + wy;       // outer interval (contains the true one)
    // the inner interval (contained in the true one) is outer -+ 2*margin, formed only for an obstacle the outer test
    // could not skip: the two extra roundings (2^-24 relative each) are far inside the 2^-18 margin
    // gridded obstacles: anchor (xlo, ylo) must lie in (xm - SX, xp) x (ym - SY, yp) to be unproven
    // un-floored cell coordinates of the window's corners (grid_cellf before its floorf)
    const float uclo = __fmul_rn(__fsub_rn(__fsub_rd(xm, h->SX), h->gx0), h->icx), uchi = __fmul_rn(__fsub_rn(xp, h->gx0), h->icx);
    const float urlo = __fmul_rn(__fsub_rn(__fsub_rd(ym, h->SY), h->gy0), h->icy), urhi = __fmul_rn(__fsub_rn(yp, h->gy0), h->icy);
    // the window misses the grid <=> floor(uchi) < 0 or floor(uclo) > 63 or floor(urhi) < 0 or floor(urlo) > R - 1
    // (floor(v) < 0 <=> v < 0, floor(v) > n <=> v >= n + 1); every comparison is false for NaN: not proven empty
    const bool empty = (uchi < 0.f) | (uclo >= (float)CCK_GRID_COLS) | (urhi < 0.f) | (urlo >= h->Rf);
    // with `pre` (mean strictly inside the fp32 bounds) and a small finite covariance this is all the second fast exit needs -
    // a non-finite or overflowing fp32 intermediate above can only make the window look NON-empty
    if (pre & cov_ok & empty & (h->n_grid == M)) return true;
    {   // only the nearest box can matter (tq < q2^2 bounds the variances, the field bounds the mean: every operand of the
        // box tests is finite and small)
        const float q2 = (float)((dfw >> 8) & 255u);
        if (cov_ok & (tq < q2 * q2)) return grid_obstacle(g, gt, (int)(dfw >> 16), xm, xp, ym, yp, mx, my, exact);
    }
    if (!pre && !exact_bounds()) return false;
    if (!enabled) return true;
    const float offd = __double2float_rn(fabs(P1) + fabs(P2));
    // every operand finite and far from fp32 overflow, variances non-negative (false for NaN)
    const bool fin = (fabsf(xf) + fabsf(yf)) + (p0 + p3) + offd < 1e30f && p0 >= 0.f && p3 >= 0.f;
    bool ok = true;
    if (!fin) {   // rare: decide every obstacle with the reference's expression
        for (int j = 0; j < M && ok; ++j) ok = exact(j);
        return ok;
    }
    const float clo = fmaxf(floorf(uclo), 0.f), chi = fminf(floorf(uchi), (float)(CCK_GRID_COLS - 1));
    const float rlo = fmaxf(floorf(urlo), 0.f), rhi = fminf(floorf(urhi), h->Rm1);
    if ((clo <= chi) & (rlo <= rhi)) {
        // dense row-major CSR: the anchors of row r, columns c0..c1 are the contiguous table rows ds[r*64+c0] .. ds[r*64+c1+1]-1
        const uint16_t* ds = reinterpret_cast<const uint16_t*>(gt + h->off_cstart);
        const int c0 = (int)clo, c1 = (int)chi + 1, r1 = (int)rhi;
        for (int r = (int)rlo; r <= r1 && ok; ++r) {
            int j = ds[r * CCK_GRID_COLS + c0];
            const int j1 = ds[r * CCK_GRID_COLS + c1];
            for (; j < j1 && ok; ++j) ok = grid_obstacle(g, gt, j, xm, xp, ym, yp, mx, my, exact);
        }
    }
    // obstacles without a finite box (non-axis-aligned planes, unusable constants): box sides are +-inf
    for (int j = h->n_grid; j < M && ok; ++j)
        ok = grid_obstacle(g, gt, j, xm, xp, ym, yp, mx, my, exact);
    return ok;
}

template <bool SPLIT = false>
__device__ __forceinline__ bool grid_check(const GridConst& g, const unsigned char* gt, const unsigned char* ctab, const double* exB,
                                           const int32_t* exCls, double c, double x, double y, double Sa0, double Sa1, double Sa2, double Sa3,
                                           double Sb0 = 0, double Sb1 = 0, double Sb2 = 0, double Sb3 = 0) {
    // the caller has done the bounds test, which a NaN mean passes (and 0 * NaN poisons every plane of the reference's
    // expression): the fast exit needs a finite mean
    const bool pre = fabs(x) + fabs(y) < 1e30;
    // the clearance field needs the mean strictly inside the fp32 bounds it was built over
    const float xf = __double2float_rn(x), yf = __double2float_rn(y);
    const uint32_t dfw = df_lookup(g, (xf > g.blx) & (xf < g.bhx) & (yf > g.bly) & (yf < g.bhy), xf, yf);
    return grid_check_b<SPLIT>(g, gt, ctab, exB, exCls, c, pre, [] { return true; }, true, x, y, Sa0, Sa1, Sa2, Sa3, Sb0, Sb1, Sb2, Sb3, nullptr, dfw);
}

// PCCBlackmoreSVC::isValid, batched (row a4), through the same broad phase: bounds, then grid_check.
__global__ void __launch_bounds__(256) k_is_valid_grid(int64_t n, int dim, const double* __restrict__ bel, int64_t ld, uint8_t* __restrict__ valid,
                                                       const unsigned char* __restrict__ ggtab, int32_t gtab_bytes, const unsigned char* __restrict__ ctab,
                                                       const double* __restrict__ exB, const int32_t* __restrict__ exCls, const __grid_constant__ SvcDev sp,
                                                       const __grid_constant__ GridConst g) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    unsigned char* gt = smem_raw + 16;
    stage_begin(gt, ggtab, (uint32_t)gtab_bytes, bar);
    bool waited = false;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double v[4] = {0, 0, 0, 0};
        const int dd = dim * dim;
        for (int f = 0; f < dim; ++f) v[f] = bel[f * ld + i];
        // (Sigma+Lambda).block<2,2>(0,0)  (PCCBlackmoreSVC.cpp:41)
        const double P0 = bel[(dim + 0) * ld + i] + bel[(dim + dd + 0) * ld + i];
        const double P1 = bel[(dim + 1) * ld + i] + bel[(dim + dd + 1) * ld + i];
        const double P2 = bel[(dim + dim) * ld + i] + bel[(dim + dd + dim) * ld + i];
        const double P3 = bel[(dim + dim + 1) * ld + i] + bel[(dim + dd + dim + 1) * ld + i];
        if (!waited) { stage_wait((uint32_t)gtab_bytes, bar); waited = true; }
        bool ok = bounds_ok(sp, v);
        if (ok && gtab_bytes > 0) ok = grid_check(g, gt, ctab, exB, exCls, sp.erf_inv_result, v[0], v[1], P0, P1, P2, P3);
        valid[i] = ok ? 1 : 0;
    }
    if (!waited) stage_wait((uint32_t)gtab_bytes, bar);   // never exit with a copy in flight
}

// ---------------------------------------------------------------------------
// cp.async (LDGSTS): per-lane prefetch of the NEXT belief into shared memory while the current
// rollout runs, so a refill never exposes HBM latency to the warp.
// ---------------------------------------------------------------------------
#ifndef CCK_HOSTSIM
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(smem)), "l"(__cvta_generic_to_global(gmem)) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem)), "l"(__cvta_generic_to_global(gmem)) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
#else   // tests/hostsim: plain copies
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) { memcpy(smem, gmem, 8); }
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) { memcpy(smem, gmem, 4); }
__device__ __forceinline__ void cp_async_commit() {}
__device__ __forceinline__ void cp_async_wait_all() {}
#endif

// shared-memory carve-up (T = CCK_GRID_THREADS): [mbarrier 16 B][grid table][work counter 16 B][stash T x 80 B]
//                         [prefetch slots: 13 planes x T doubles][2 planes x T ints]
// Sweeps (EXPAND = false, profiles/r2c_so_sweep.txt): 4 warps x 4 CTAs at 122 registers is the shipped shape; 5 warps x 4 CTAs
// needs a 96-register cap, under which the round-2 kernel spills (164.6 -> 143.6 G belief-steps/s).  The expansion kernel
// carries the constraint and goal tests: 128 threads x 3 CTAs leaves it 170 registers (160 used, no spills); its launches
// are planner-sized and latency-bound.
#ifndef CCK_GRID_THREADS
#define CCK_GRID_THREADS 128
#endif
#ifndef CCK_GRID_MINB
#define CCK_GRID_MINB 4
#endif
__host__ __device__ constexpr int grid_threads(bool expand) { return expand ? 128 : CCK_GRID_THREADS; }
__host__ __device__ constexpr int grid_minb(bool expand) { return expand ? 3 : CCK_GRID_MINB; }
#ifndef CCK_REFILL_MIN
#define CCK_REFILL_MIN 10     // idle lanes of a warp that trigger a refill ...
#endif
#ifndef CCK_REFILL_WAIT
#define CCK_REFILL_WAIT 8    // ... or this many loop iterations with an idle lane (power of two)
#endif
// the work counter hands out ranges of 2^range_log2 consecutive beliefs, round-robin over the CTAs: 256 for the
// big sweeps (coalesced refills), 128 = one belief per lane for planner-sized batches (latency: every candidate of a
// K <= 75k launch starts immediately on its own lane)
__host__ inline int rollout_range_log2(int64_t n) { return n >= ((int64_t)1 << 22) ? 8 : 7; }
__host__ __device__ inline int grid_smem_bytes(int gtab_bytes, int threads) {
    return 16 + gtab_bytes + 16 + threads * 80 + 13 * threads * 8 + 2 * threads * 4;
}

template <bool TRAJ, bool EXPAND>
__global__ void __launch_bounds__(grid_threads(EXPAND), grid_minb(EXPAND)) k_rollout_lin_grid(const __grid_constant__ RolloutArgs a, const __grid_constant__ LinDev p,
                                                                          const __grid_constant__ SvcDev sp, const __grid_constant__ GridConst g,
                                                                          const int range_log2) {
    constexpr int T = grid_threads(EXPAND);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    unsigned char* gt = smem_raw + 16;
    int* next_k = reinterpret_cast<int*>(smem_raw + 16 + a.ftab_bytes);
    double2* stash = reinterpret_cast<double2*>(smem_raw + 16 + a.ftab_bytes + 16) + threadIdx.x * 5;   // 80 B per thread, conflict-free STS.128
    double* slot = reinterpret_cast<double*>(smem_raw + 16 + a.ftab_bytes + 16 + T * 80) + threadIdx.x;   // plane f at slot[f*T]
    int* islot = reinterpret_cast<int*>(smem_raw + 16 + a.ftab_bytes + 16 + T * 80 + 13 * T * 8) + threadIdx.x;
    if (threadIdx.x == 0) *next_k = 0;
    // EXPAND: a small constraint table is copied behind the prefetch slots (a dependent global load per step and
    // candidate is what a planner-sized launch would otherwise wait for)
    const unsigned char* cons = a.cons;
    if (EXPAND && a.cons_smem_bytes > 0) {
        unsigned char* sc = smem_raw + grid_smem_bytes(a.ftab_bytes, T);
        for (int o = threadIdx.x * 16; o < a.cons_smem_bytes; o += T * 16)
            *reinterpret_cast<uint4*>(sc + o) = *reinterpret_cast<const uint4*>(a.cons + o);
        cons = sc;
    }
    stage_begin(gt, a.ftab, (uint32_t)a.ftab_bytes, bar);   // __syncthreads inside
    const double eps = DBL_EPSILON;
    const double c = sp.erf_inv_result;

    int nxt = -1;                  // belief indices fit 32 bits (the launcher refuses n >= 2^31)
    // Claim the next beliefs of this CTA (ranges blockIdx, blockIdx+gridDim, ... of 2^range_log2 beliefs) for the lanes that
    // `want` one and start copying them.  WARP-SYNCHRONOUS (every lane of the warp calls it): ONE atomicAdd per warp hands
    // out consecutive beliefs to the wanting lanes in lane order, so their 8-byte copies of a plane are contiguous - four
    // lanes fill one 32-byte sector (ncu, round 1: one atomicAdd per lane and 1.58x the algorithmic DRAM traffic on the
    // realistic sweep, because a sector fetched for one lane was evicted before its neighbours asked for the rest).
    const unsigned lane = threadIdx.x & 31u;
    auto prefetch = [&](bool want) {
        const unsigned mw = __ballot_sync(0xffffffffu, want);
        if (mw == 0) return;
        const int cnt = __popc(mw);
        int base = 0;
        if (lane == (unsigned)(__ffs(mw) - 1)) base = atomicAdd(next_k, cnt);
        base = __shfl_sync(0xffffffffu, base, __ffs(mw) - 1);
        if (!want) return;
        nxt = -1;
        const int kk = base + __popc(mw & ((1u << lane) - 1u));
        const int rg = kk >> range_log2;
        const int64_t rbase = ((int64_t)blockIdx.x + (int64_t)rg * gridDim.x) << range_log2;
        const int64_t idx = rbase + (kk & ((1 << range_log2) - 1));
        if (idx >= a.n) return;          // past the end of the batch (later ranges start even further out)
        nxt = (int)idx;
#pragma unroll
        for (int f = 0; f < 10; ++f) cp_async8(slot + f * T, a.bel + f * a.ld + nxt);
        cp_async8(slot + 10 * T, a.ctrl + nxt);
        cp_async8(slot + 11 * T, a.ctrl + a.ldc + nxt);
        cp_async4(islot, a.steps + nxt);
        if (EXPAND) cp_async4(islot + T, a.parent_step + nxt);
        cp_async_commit();
    };
    prefetch(true);
    stage_wait((uint32_t)a.ftab_bytes, bar);

    bool have = false, pending = false;
    bool dense = true;             // this lane's next step must be the dense lin_step (see lin_step_diag's preconditions)
    bool use_tab = false;          // EXPAND: this candidate's covariances come from the tree's covariance table (RolloutArgs::cov_tab)
    Lin cur;
    double dux = 0, duy = 0;
    int steps = 0, t = 0, pstep = 0;
    int i = 0;
    bool cons_ok = true;
    cur.mx = cur.my = 0;
#pragma unroll
    for (int f = 0; f < 4; ++f) cur.S[f] = cur.L[f] = 0;
    // store the result of a finished rollout (a lane that is idle keeps its registers and its stash until then)
    auto flush = [&](Lin& B) {
        if (!pending) return;
        pending = false;
        if (t < steps) {
            const double2 s0 = stash[0], s1 = stash[1], s2 = stash[2], s3 = stash[3], s4 = stash[4];
            B.mx = s0.x; B.my = s0.y; B.S[0] = s1.x; B.S[1] = s1.y; B.S[2] = s2.x; B.S[3] = s2.y;
            B.L[0] = s3.x; B.L[1] = s3.y; B.L[2] = s4.x; B.L[3] = s4.y;
        }
        a.out[i] = B.mx; a.out[a.ldo + i] = B.my;
#pragma unroll
        for (int f = 0; f < 4; ++f) { a.out[(2 + f) * a.ldo + i] = B.S[f]; a.out[(6 + f) * a.ldo + i] = B.L[f]; }
        a.valid[i] = t;
        if (EXPAND) {
            // the rollout's start and the parent's cost are re-read here (L2 hits) instead of riding in six registers
            const double d0 = a.bel[i] - B.mx, d1 = a.bel[a.ld + i] - B.my;
            a.cost[i] = a.parent_cost[i] + sqrt(d0 * d0 + d1 * d1);
            double gd;
            const bool g = goal_ok(a.goal, sp, B.mx, B.my, B.S[0] + B.L[0], B.S[1] + B.L[1], B.S[2] + B.L[2], B.S[3] + B.L[3], &gd);
            a.goal_dist[i] = gd;
            a.flags[i] = (uint8_t)((t == steps ? 1 : 0) | (cons_ok ? 2 : 0) | (g ? 4 : 0) | (use_tab ? 8 : 0));
        }
    };
    // refill (warp-synchronous): idle lanes take their prefetched belief and the warp claims the ones after them with one
    // atomicAdd.  (Serving whole quads only, so that four lanes always fill one 32-byte sector per plane, was measured:
    // the up-to-three lanes left waiting cost more lane-time than the sectors saved - the sweep is not DRAM-bound.)
    // A belief with steps <= 0 is answered on the spot; its lane picks up real work in the next round.
    auto take = [&](Lin& A) {
        flush(A);
        const bool cand = !have && nxt >= 0;
        const bool serve = cand;
        if (serve) {
            cp_async_wait_all();
            i = nxt;
            A.mx = slot[0]; A.my = slot[T];
#pragma unroll
            for (int f = 0; f < 4; ++f) { A.S[f] = slot[(2 + f) * T]; A.L[f] = slot[(6 + f) * T]; }
            dux = p.dt * slot[10 * T]; duy = p.dt * slot[11 * T];
            steps = islot[0];
            if (EXPAND) pstep = islot[T];
        }
        prefetch(serve);
        if (serve) {
            if (steps <= 0) {         // propagateWhileValid(.., 0, ..): copy, return 0
                a.out[i] = A.mx; a.out[a.ldo + i] = A.my;
#pragma unroll
                for (int f = 0; f < 4; ++f) { a.out[(2 + f) * a.ldo + i] = A.S[f]; a.out[(6 + f) * a.ldo + i] = A.L[f]; }
                a.valid[i] = 0;
                if (EXPAND) {
                    double gd;
                    const bool g = goal_ok(a.goal, sp, A.mx, A.my, A.S[0] + A.L[0], A.S[1] + A.L[1], A.S[2] + A.L[2], A.S[3] + A.L[3], &gd);
                    a.cost[i] = a.parent_cost[i] + 0.0;
                    a.goal_dist[i] = gd;
                    a.flags[i] = (uint8_t)(1 | 2 | (g ? 4 : 0));
                }
            } else {
                t = 0; cons_ok = true;
                dense = !p.diag || !cov_small(A.S[0] + A.L[0], A.S[1] + A.L[1], A.S[2] + A.L[2], A.S[3] + A.L[3]);
                if (EXPAND) {      // parent covariance bit-identical to the table's entry for its depth, and the segment inside the table
                    use_tab = false;
                    if (a.cov_tab && pstep >= 0 && pstep + steps < a.cov_tab_n) {
                        const double* e = a.cov_tab + (size_t)pstep * 8;
                        bool same = true;
#pragma unroll
                        for (int f = 0; f < 4; ++f)
                            same = same && __double_as_longlong(A.S[f]) == __double_as_longlong(__ldg(e + f)) &&
                                   __double_as_longlong(A.L[f]) == __double_as_longlong(__ldg(e + 4 + f));
                        use_tab = same;
                    }
                }
                have = true;
            }
        }
    };
    // One belief-step of a lane that has work, in place (lin_step is alias-safe), in two phases: stepA advances the state,
    // stepB checks the new state and commits.  (x, y, pre, dfw) carry the new mean, its fp32 bounds pre-check and its
    // clearance-field word from one phase to the other.
    double x = 0, y = 0;
    bool pre = false;
    uint32_t dfw = 0u;
    auto stepA = [&](Lin& A, bool use_dense) {
        // stash the last valid state
        stash[0] = make_double2(A.mx, A.my);
        stash[1] = make_double2(A.S[0], A.S[1]); stash[2] = make_double2(A.S[2], A.S[3]);
        stash[3] = make_double2(A.L[0], A.L[1]); stash[4] = make_double2(A.L[2], A.L[3]);
        // the new mean first (UncertainLinearSP.cpp:92-93; the step functions below form the same sums): its clearance-field
        // cell is looked up NOW so that the load is in flight during the covariance recursion
        x = A.mx + dux; y = A.my + duy;
        // satisfiesBounds (NaN passes): the mean rounded to fp32 strictly inside the inward-rounded bounds proves it
        // (RN is monotone: xf < bhx <= high implies x < high); otherwise the exact fp64 test decides
        const float xf = __double2float_rn(x), yf = __double2float_rn(y);
        pre = (xf > g.blx) & (xf < g.bhx) & (yf > g.bly) & (yf < g.bhy);
        dfw = df_lookup(g, pre, xf, yf);
        if (EXPAND && use_tab) {   // the covariances of absolute step pstep + t + 1, as k_cov_table_lin computed them
            const double2* e = reinterpret_cast<const double2*>(a.cov_tab + (size_t)(pstep + t + 1) * 8);
            const double2 s01 = __ldg(e), s23 = __ldg(e + 1), l01 = __ldg(e + 2), l23 = __ldg(e + 3);
            A.mx = x; A.my = y;
            A.S[0] = s01.x; A.S[1] = s01.y; A.S[2] = s23.x; A.S[3] = s23.y;
            A.L[0] = l01.x; A.L[1] = l01.y; A.L[2] = l23.x; A.L[3] = l23.y;
            return;
        }
        if (use_dense) lin_step(A, dux, duy, p, A); else lin_step_diag(A, dux, duy, p, A);
    };
    // commit a valid step / end the rollout
    auto commit = [&](Lin& B, bool ok) {
        if (ok) {
            if (EXPAND && cons) {   // Sigma + Lambda formed here, not carried across the validity check
                double s0 = B.S[0], s1 = B.S[1], s2 = B.S[2], s3 = B.S[3];
                CCK_OPAQUE4(s0, s1, s2, s3);
                cons_ok = cons_ok && constraints_ok_at(cons, pstep + t + 1, x, y, s0 + B.L[0], s1 + B.L[1], s2 + B.L[2], s3 + B.L[3]);
            }
            if (TRAJ && t < a.traj_T) {
                double* tp = a.traj + (int64_t)t * 10 * a.traj_ld + i;
                tp[0] = B.mx; tp[a.traj_ld] = B.my;
#pragma unroll
                for (int f = 0; f < 4; ++f) { tp[(2 + f) * a.traj_ld] = B.S[f]; tp[(6 + f) * a.traj_ld] = B.L[f]; }
            }
            ++t;
        }
        // the rollout ends with t valid steps: died (t < steps; the result is the last valid state, still in the stash) or
        // complete (t == steps).  The result is stored by flush(), together with the other finished lanes of the warp.
        if (!ok || t == steps) { have = false; pending = true; }
    };
    auto stepB = [&](Lin& B) {
        bool small;
        const bool ok = grid_check_b<true>(
            g, gt, a.tab, a.exB, a.exCls, c, pre,
            [&] { return !(x - eps > sp.high[0] || x + eps < sp.low[0] || y - eps > sp.high[1] || y + eps < sp.low[1]); }, a.ftab_bytes > 0, x, y,
            B.S[0], B.S[1], B.S[2], B.S[3], B.L[0], B.L[1], B.L[2], B.L[3], &small, dfw);
        dense = !p.diag || !small;   // the new covariance is finite: the next step may skip the exact-zero terms
        commit(B, ok);
    };
    // Lanes are refilled as soon as their rollout ends, but in batches: the take / prefetch / result-store code is ~150
    // instructions that a single finished lane would make the whole warp issue.  Idle lanes therefore wait until
    // CCK_REFILL_MIN lanes of the warp are idle, or at most CCK_REFILL_WAIT iterations (ncu, realistic sweep: the refill
    // path was 24 % of the issue slots at 2.1 active lanes on average).
    int it = 0;
    bool checked = true;           // false: stepA has run for this lane, stepB is still due (set by the uniform section)
    while (true) {
        // ---- uniform section: every lane of the warp has work and may take the diagonal-constant step.  The warp then runs
        // min(steps - t) steps without the per-lane bookkeeping of the general loop (idle ballot, have / dense branches,
        // per-lane end test): per step ONE vote on "every lane passed the clearance-field test".  The first step some lane
        // does not pass falls through to the general section, which finishes exactly that step (stepB) for every lane.
        // An FP64 instruction costs two issue cycles and everything else one (cck_measure_fp64_rates), so the ~50 warp
        // instructions this saves per step are worth ~20 % on the all-survive sweep.
        if (!EXPAND && g.df != nullptr && __all_sync(0xffffffffu, have && !dense)) {
            int nf = (int)__reduce_min_sync(0xffffffffu, (unsigned)(steps - t));
            while (nf > 0) {
                stepA(cur, false);
                const double P0 = cur.S[0] + cur.L[0], P1 = cur.S[1] + cur.L[1], P2 = cur.S[2] + cur.L[2], P3 = cur.S[3] + cur.L[3];
                const float q1 = (float)(dfw & 255u);
                const bool fast = cov_small(P0, P1, P2, P3) & (__fmul_rn(fmaxf(__double2float_rn(P0), __double2float_rn(P3)), g.df_k2) < q1 * q1);
                if (!__all_sync(0xffffffffu, fast)) { checked = false; break; }
                if (TRAJ && t < a.traj_T) {
                    double* tp = a.traj + (int64_t)t * 10 * a.traj_ld + i;
                    tp[0] = cur.mx; tp[a.traj_ld] = cur.my;
#pragma unroll
                    for (int f = 0; f < 4; ++f) { tp[(2 + f) * a.traj_ld] = cur.S[f]; tp[(6 + f) * a.traj_ld] = cur.L[f]; }
                }
                ++t; --nf;
            }
            if (checked && t == steps) { have = false; pending = true; }
        }
        if (!checked) {            // warp-uniform: finish the step the uniform section started
            stepB(cur);
            checked = true;
        }
        const unsigned idle = __ballot_sync(0xffffffffu, !have);
        if (idle) {
            if (__popc(idle) >= CCK_REFILL_MIN || (++it & (CCK_REFILL_WAIT - 1)) == 0) {
                take(cur);
                if (__all_sync(0xffffffffu, !have && nxt < 0)) break;   // nothing running and nothing prefetched
            }
        }
        if (have) { stepA(cur, dense); stepB(cur); }
    }
}

}  // namespace cck
