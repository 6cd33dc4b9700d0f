// cck_rollout_bm.cuh — the fused propagateWhileValid kernel for the default configuration
// (2DUncertainLinear + PCCBlackmoreSVC), restructured around what ncu showed on the first
// versions (profiles/r1a_*, r1b_*): the per-obstacle loop was co-limited by LDS broadcast
// bandwidth (2 x LDS.128 per obstacle per warp) and FP64 compare/add chains, and the per-class
// sqrt work (4 sqrt x 6..16 classes per belief-step) had become a third of the FP64 time.
//
//  * CONSERVATIVE FP32 FILTER, EXACT FP64 DECISION.  For a super-class of (near-)identical
//    normals the four inequalities  a_i.mu >= fl(B_oi + bb_i)  become four thresholds
//    tau_i = (lhs_i - bb_i) - margin, rounded DOWN to fp32; the table holds B_oi rounded UP to
//    fp32.  Bf_oi <= tauf_i  PROVES the reference's fp64 inequality (rounding is monotone and
//    the margin covers the ulp-level spread of normals inside the super-class and all rounding),
//    so the common case costs one LDS.128 and four FSETP per obstacle and touches neither the
//    FP64 pipe nor a sqrt.  Whatever is not proven (a real collision, the ~1e-5-wide fp32 band,
//    cancellation in a^T P a, NaN/inf operands) is decided by the reference's exact fp64
//    expression on the obstacle's own normals (bm_exact_obstacle).  Verdicts are bit-identical.
//  * BPT beliefs per thread share every table row loaded from shared memory.
//  * only the current state lives in registers; a belief that dies at step r is re-played
//    (r propagate steps, no checks) to produce its last valid state.
#pragma once
#include "cck_kernels.cuh"

namespace cck {

// exact reference test of one obstacle (slow path; PCCBlackmoreSVC.cpp:63-75)
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

// ok[b] &= "belief b is safe w.r.t. every obstacle" (PCCBlackmoreSVC.cpp:54-58), bit-exact.
template <int BPT>
__device__ __forceinline__ void bm_filter_check(const unsigned char* ftab, const unsigned char* gtab, const double* exB,
                                                const int32_t* exCls, double c, const double (&x)[BPT], const double (&y)[BPT],
                                                const double (&P)[BPT][4], const bool (&act)[BPT], bool (&ok)[BPT]) {
    const FHeader* fh = reinterpret_cast<const FHeader*>(ftab);
    const SuperRec* sr = reinterpret_cast<const SuperRec*>(ftab + fh->off_super);
    const float4* rows = reinterpret_cast<const float4*>(ftab + fh->off_rows);
    const int ns = fh->n_super;
    for (int s = 0; s < ns; ++s) {
        float tf[BPT][4];
#pragma unroll
        for (int b = 0; b < BPT; ++b) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const double a0 = sr[s].A[2 * i], a1 = sr[s].A[2 * i + 1];
                const double t0 = a0 * P[b][0] + a1 * P[b][2];
                const double t1 = a0 * P[b][1] + a1 * P[b][3];
                const double q = t0 * a0 + t1 * a1;
                // cancellation guard: |terms| of a^T P a; with q >= 2^-10 * qabs a 1e-14 spread of the normals
                // perturbs sqrt(q) by < 1e-10 relative
                const double qabs = fabs(a0 * a0 * P[b][0]) + fabs(a0 * a1) * (fabs(P[b][1]) + fabs(P[b][2])) + fabs(a1 * a1 * P[b][3]);
                const double bb = 1.4142135623730951 * sqrt(q) * c;
                const double xa = x[b] * a0, ya = y[b] * a1;
                const double tau = ((xa + ya) - bb) - (fabs(xa) + fabs(ya) + fabs(bb)) * 3.725290298461914e-09;   // 2^-28
                const bool good = q >= qabs * 9.765625e-4;                                                        // false for NaN
                tf[b][i] = act[b] ? (good ? __double2float_rd(tau) : __int_as_float(0x7fc00000)) : __int_as_float(0x7f800000);
            }
        }
        // level 1 (warp-uniform, broadcast loads): group rows; level 2 (per lane): members of unproven groups
        const float4* grp = reinterpret_cast<const float4*>(ftab + fh->off_groups) + sr[s].first_group;
        const int ng = sr[s].n_groups, first = sr[s].first_row;
        for (int g0 = 0; g0 < ng; g0 += 32) {
            unsigned need[BPT];
#pragma unroll
            for (int b = 0; b < BPT; ++b) need[b] = 0u;
            const int gn = min(32, ng - g0);
#pragma unroll 4
            for (int j = 0; j < gn; ++j) {
                const float4 r = grp[g0 + j];
#pragma unroll
                for (int b = 0; b < BPT; ++b) {
                    const bool pv = (r.x <= tf[b][0]) | (r.y <= tf[b][1]) | (r.z <= tf[b][2]) | (r.w <= tf[b][3]);
                    need[b] |= (pv ? 0u : 1u) << j;
                }
            }
#pragma unroll
            for (int b = 0; b < BPT; ++b) {
                unsigned m = ok[b] ? need[b] : 0u;
                while (m) {
                    const int j = __ffs(m) - 1;
                    m &= m - 1;
                    const int row0 = first + (g0 + j);          // member q of this group sits at row0 + q*ng (transposed)
                    unsigned fail = 0u;                          // members the fp32 filter could not prove safe
#pragma unroll
                    for (int q = 0; q < CCK_GROUP; ++q) {
                        const float4 r = rows[row0 + q * ng];
                        const bool pv = (r.x <= tf[b][0]) | (r.y <= tf[b][1]) | (r.z <= tf[b][2]) | (r.w <= tf[b][3]);
                        fail |= (pv ? 0u : 1u) << q;
                    }
                    if (fail) {   // rare: exact fp64 decision on those obstacles' own normals
                        const TabHeader* gh = reinterpret_cast<const TabHeader*>(gtab);
                        const ClassRec* cls = reinterpret_cast<const ClassRec*>(gtab + gh->off_classes);
                        while (fail && ok[b]) {
                            const int q = __ffs(fail) - 1;
                            fail &= fail - 1;
                            const int row = row0 + q * ng, k = exCls[row];
                            if (k >= 0) ok[b] = bm_exact_obstacle(cls[k].A, exB + (size_t)row * 4, x[b], y[b], P[b][0], P[b][1], P[b][2], P[b][3], c);
                        }
                        if (!ok[b]) m = 0;
                    }
                }
            }
        }
    }
}

template <int BPT>
__device__ __forceinline__ void load_planes(const double* __restrict__ p, int64_t i, bool vec, double (&v)[BPT]) {
    if (BPT == 2 && vec) {
        const double2 t = *reinterpret_cast<const double2*>(p + i);
        v[0] = t.x; v[1] = t.y;
    } else if (BPT == 4 && vec) {
        const double2 t = *reinterpret_cast<const double2*>(p + i), u = *reinterpret_cast<const double2*>(p + i + 2);
        v[0] = t.x; v[1] = t.y; v[2] = u.x; v[3] = u.y;
    } else {
#pragma unroll
        for (int b = 0; b < BPT; ++b) v[b] = p[i + b];
    }
}
template <int BPT>
__device__ __forceinline__ void store_planes(double* __restrict__ p, int64_t i, bool vec, const double (&v)[BPT]) {
    if (BPT == 2 && vec) {
        *reinterpret_cast<double2*>(p + i) = make_double2(v[0], v[1]);
    } else if (BPT == 4 && vec) {
        *reinterpret_cast<double2*>(p + i) = make_double2(v[0], v[1]);
        *reinterpret_cast<double2*>(p + i + 2) = make_double2(v[2], v[3]);
    } else {
#pragma unroll
        for (int b = 0; b < BPT; ++b) p[i + b] = v[b];
    }
}

template <int BPT, bool TRAJ, bool EXPAND>
__global__ void __launch_bounds__(128, (BPT == 1 ? 4 : 1)) k_rollout_lin_bm(const __grid_constant__ RolloutArgs a, const __grid_constant__ LinDev p,
                                                        const __grid_constant__ SvcDev sp) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    unsigned char* ftab = smem_raw + 16;
    stage_begin(ftab, a.ftab, (uint32_t)a.ftab_bytes, bar);
    bool waited = false;
    // vector (16-byte) plane accesses need even leading dimensions and 16-byte aligned bases
    const bool vec = BPT > 1 && ((a.ld | a.ldc | a.ldo) & 1) == 0 && (((uintptr_t)a.bel | (uintptr_t)a.ctrl | (uintptr_t)a.out) & 15) == 0;
    const double eps = DBL_EPSILON;
    const double c = sp.erf_inv_result;

    for (int64_t base = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * BPT; base < a.n;
         base += (int64_t)gridDim.x * blockDim.x * BPT) {
        const bool full = base + BPT <= a.n;            // all BPT slots are real beliefs
        Lin cur[BPT];
        double dux[BPT], duy[BPT];
        int steps[BPT], r[BPT];
        bool alive[BPT];
        if (full) {
            double t[BPT];
            load_planes<BPT>(a.bel, base, vec, t);
#pragma unroll
            for (int b = 0; b < BPT; ++b) cur[b].mx = t[b];
            load_planes<BPT>(a.bel + a.ld, base, vec, t);
#pragma unroll
            for (int b = 0; b < BPT; ++b) cur[b].my = t[b];
#pragma unroll
            for (int f = 0; f < 4; ++f) {
                load_planes<BPT>(a.bel + (2 + f) * a.ld, base, vec, t);
#pragma unroll
                for (int b = 0; b < BPT; ++b) cur[b].S[f] = t[b];
                load_planes<BPT>(a.bel + (6 + f) * a.ld, base, vec, t);
#pragma unroll
                for (int b = 0; b < BPT; ++b) cur[b].L[f] = t[b];
            }
            load_planes<BPT>(a.ctrl, base, vec, t);
#pragma unroll
            for (int b = 0; b < BPT; ++b) dux[b] = p.dt * t[b];
            load_planes<BPT>(a.ctrl + a.ldc, base, vec, t);
#pragma unroll
            for (int b = 0; b < BPT; ++b) duy[b] = p.dt * t[b];
#pragma unroll
            for (int b = 0; b < BPT; ++b) steps[b] = a.steps[base + b];
        } else {
#pragma unroll
            for (int b = 0; b < BPT; ++b) {
                const bool in = base + b < a.n;
                const int64_t i = in ? base + b : a.n - 1;
                cur[b].mx = a.bel[i]; cur[b].my = a.bel[a.ld + i];
#pragma unroll
                for (int f = 0; f < 4; ++f) { cur[b].S[f] = a.bel[(2 + f) * a.ld + i]; cur[b].L[f] = a.bel[(6 + f) * a.ld + i]; }
                dux[b] = p.dt * a.ctrl[i]; duy[b] = p.dt * a.ctrl[a.ldc + i];
                steps[b] = in ? a.steps[i] : 0;
            }
        }
        int tmax = 0;
        double x0[BPT], y0[BPT];
        int pstep[BPT];
        bool cons_ok[BPT];
#pragma unroll
        for (int b = 0; b < BPT; ++b) {
            r[b] = steps[b] > 0 ? steps[b] : 0;
            alive[b] = steps[b] > 0;
            tmax = max(tmax, steps[b]);
            x0[b] = cur[b].mx; y0[b] = cur[b].my;
            cons_ok[b] = true;
            pstep[b] = 0;
            if (EXPAND) pstep[b] = a.parent_step[min(base + b, a.n - 1)];
        }
        if (!waited) { stage_wait((uint32_t)a.ftab_bytes, bar); waited = true; }

        for (int t = 0; t < tmax; ++t) {
            bool any = false;
#pragma unroll
            for (int b = 0; b < BPT; ++b) { if (alive[b] && t >= steps[b]) alive[b] = false; any = any || alive[b]; }
            if (!any) break;
            bool ok[BPT];
            double x[BPT], y[BPT], P[BPT][4];
#pragma unroll
            for (int b = 0; b < BPT; ++b) {
                if (alive[b]) lin_step(cur[b], dux[b], duy[b], p, cur[b]);
                x[b] = cur[b].mx; y[b] = cur[b].my;
#pragma unroll
                for (int f = 0; f < 4; ++f) P[b][f] = cur[b].S[f] + cur[b].L[f];
                // satisfiesBounds (NaN passes)
                ok[b] = !(x[b] - eps > sp.high[0] || x[b] + eps < sp.low[0] || y[b] - eps > sp.high[1] || y[b] + eps < sp.low[1]);
            }
            if (a.ftab_bytes > 0) bm_filter_check<BPT>(ftab, a.tab, a.exB, a.exCls, c, x, y, P, alive, ok);
#pragma unroll
            for (int b = 0; b < BPT; ++b) {
                if (!alive[b]) continue;
                if (!ok[b]) { alive[b] = false; r[b] = t; continue; }
                if (EXPAND && a.cons) cons_ok[b] = cons_ok[b] && constraints_ok_at(a.cons, pstep[b] + t + 1, x[b], y[b], P[b][0], P[b][1], P[b][2], P[b][3]);
                if (TRAJ && base + b < a.n) {
                    double* tp = a.traj + (int64_t)t * 10 * a.traj_ld + base + b;
                    tp[0] = cur[b].mx; tp[a.traj_ld] = cur[b].my;
#pragma unroll
                    for (int f = 0; f < 4; ++f) { tp[(2 + f) * a.traj_ld] = cur[b].S[f]; tp[(6 + f) * a.traj_ld] = cur[b].L[f]; }
                }
            }
        }
        // beliefs that died: replay r steps from the stored start state (bit-identical recursion)
#pragma unroll
        for (int b = 0; b < BPT; ++b) {
            if (r[b] < steps[b]) {
                const int64_t i = base + b;
                cur[b].mx = a.bel[i]; cur[b].my = a.bel[a.ld + i];
#pragma unroll
                for (int f = 0; f < 4; ++f) { cur[b].S[f] = a.bel[(2 + f) * a.ld + i]; cur[b].L[f] = a.bel[(6 + f) * a.ld + i]; }
                for (int t = 0; t < r[b]; ++t) lin_step(cur[b], dux[b], duy[b], p, cur[b]);
            }
        }
        if (full) {
            double t[BPT];
#pragma unroll
            for (int b = 0; b < BPT; ++b) t[b] = cur[b].mx;
            store_planes<BPT>(a.out, base, vec, t);
#pragma unroll
            for (int b = 0; b < BPT; ++b) t[b] = cur[b].my;
            store_planes<BPT>(a.out + a.ldo, base, vec, t);
#pragma unroll
            for (int f = 0; f < 4; ++f) {
#pragma unroll
                for (int b = 0; b < BPT; ++b) t[b] = cur[b].S[f];
                store_planes<BPT>(a.out + (2 + f) * a.ldo, base, vec, t);
#pragma unroll
                for (int b = 0; b < BPT; ++b) t[b] = cur[b].L[f];
                store_planes<BPT>(a.out + (6 + f) * a.ldo, base, vec, t);
            }
        }
#pragma unroll
        for (int b = 0; b < BPT; ++b) {
            const int64_t i = base + b;
            if (i >= a.n) continue;
            if (!full) {
                a.out[i] = cur[b].mx; a.out[a.ldo + i] = cur[b].my;
#pragma unroll
                for (int f = 0; f < 4; ++f) { a.out[(2 + f) * a.ldo + i] = cur[b].S[f]; a.out[(6 + f) * a.ldo + i] = cur[b].L[f]; }
            }
            a.valid[i] = r[b];
            if (EXPAND) {
                const double d0 = x0[b] - cur[b].mx, d1 = y0[b] - cur[b].my;
                a.cost[i] = a.parent_cost[i] + sqrt(d0 * d0 + d1 * d1);
                double gd;
                const bool g = goal_ok(a.goal, sp, cur[b].mx, cur[b].my, cur[b].S[0] + cur[b].L[0], cur[b].S[1] + cur[b].L[1],
                                       cur[b].S[2] + cur[b].L[2], cur[b].S[3] + cur[b].L[3], &gd);
                a.goal_dist[i] = gd;
                a.flags[i] = (uint8_t)((r[b] == steps[b] ? 1 : 0) | (cons_ok[b] ? 2 : 0) | (g ? 4 : 0));
            }
        }
    }
    if (!waited) stage_wait((uint32_t)a.ftab_bytes, bar);
}

// ---------------------------------------------------------------------------
// Lane-refill variant (one belief per lane): a lane whose belief died or finished immediately
// writes its result and pulls the next belief of the CTA's range from a shared-memory counter,
// so warps stay full when rollouts end at different steps (the realistic workload: ~2/3 of the
// beliefs die early).  Per-step work is identical to k_rollout_lin_bm<1,...>.
// ---------------------------------------------------------------------------
template <bool TRAJ, bool EXPAND>
__global__ void __launch_bounds__(128, 4) k_rollout_lin_refill(const __grid_constant__ RolloutArgs a, const __grid_constant__ LinDev p,
                                                               const __grid_constant__ SvcDev sp, int64_t per_cta) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ int next_rel;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    unsigned char* ftab = smem_raw + 16;
    stage_begin(ftab, a.ftab, (uint32_t)a.ftab_bytes, bar);
    stage_wait((uint32_t)a.ftab_bytes, bar);
    const double eps = DBL_EPSILON;
    const double c = sp.erf_inv_result;

    for (int64_t cta_lo = (int64_t)blockIdx.x * per_cta; cta_lo < a.n; cta_lo += (int64_t)gridDim.x * per_cta) {
        const int64_t cta_hi = min(a.n, cta_lo + per_cta);
        __syncthreads();
        if (threadIdx.x == 0) next_rel = 0;
        __syncthreads();
        bool have = false, done = false;
        Lin cur;
        double dux = 0, duy = 0, x0 = 0, y0 = 0;
        int steps = 0, t = 0, pstep = 0;
        int64_t i = 0;
        bool cons_ok = true;
        cur.mx = cur.my = 0;
#pragma unroll
        for (int f = 0; f < 4; ++f) cur.S[f] = cur.L[f] = 0;
        while (true) {
            while (!have && !done) {                       // fetch the next belief of this CTA's range
                i = cta_lo + atomicAdd(&next_rel, 1);
                if (i >= cta_hi) { done = true; break; }
                steps = a.steps[i];
                if (steps <= 0) {                          // propagateWhileValid(.., 0, ..): copy, return 0
#pragma unroll
                    for (int f = 0; f < 10; ++f) a.out[f * a.ldo + i] = a.bel[f * a.ld + i];
                    a.valid[i] = 0;
                    if (EXPAND) {
                        const double bx = a.bel[i], by = a.bel[a.ld + i];
                        double gd;
                        const bool g = goal_ok(a.goal, sp, bx, by, a.bel[2 * a.ld + i] + a.bel[6 * a.ld + i], a.bel[3 * a.ld + i] + a.bel[7 * a.ld + i],
                                               a.bel[4 * a.ld + i] + a.bel[8 * a.ld + i], a.bel[5 * a.ld + i] + a.bel[9 * a.ld + i], &gd);
                        a.cost[i] = a.parent_cost[i] + 0.0;
                        a.goal_dist[i] = gd;
                        a.flags[i] = (uint8_t)(1 | 2 | (g ? 4 : 0));
                    }
                    continue;
                }
                cur.mx = a.bel[i]; cur.my = a.bel[a.ld + i];
#pragma unroll
                for (int f = 0; f < 4; ++f) { cur.S[f] = a.bel[(2 + f) * a.ld + i]; cur.L[f] = a.bel[(6 + f) * a.ld + i]; }
                dux = p.dt * a.ctrl[i]; duy = p.dt * a.ctrl[a.ldc + i];
                x0 = cur.mx; y0 = cur.my;
                t = 0; cons_ok = true;
                if (EXPAND) pstep = a.parent_step[i];
                have = true;
            }
            if (__all_sync(0xffffffffu, !have)) break;
            bool ok[1];
            double x[1], y[1], P[1][4];
            const bool act[1] = {have};
            const Lin prev = cur;                          // last valid state (restored if this step is invalid)
            if (have) lin_step(cur, dux, duy, p, cur);
            x[0] = cur.mx; y[0] = cur.my;
#pragma unroll
            for (int f = 0; f < 4; ++f) P[0][f] = cur.S[f] + cur.L[f];
            ok[0] = !(x[0] - eps > sp.high[0] || x[0] + eps < sp.low[0] || y[0] - eps > sp.high[1] || y[0] + eps < sp.low[1]);
            if (a.ftab_bytes > 0) bm_filter_check<1>(ftab, a.tab, a.exB, a.exCls, c, x, y, P, act, ok);
            if (!have) continue;
            int r = -1;                                    // >= 0: this rollout ends now with r valid steps
            if (!ok[0]) {
                r = t;                                     // died: the result is the last valid state
                cur = prev;
            } else {
                if (EXPAND && a.cons) cons_ok = cons_ok && constraints_ok_at(a.cons, pstep + t + 1, x[0], y[0], P[0][0], P[0][1], P[0][2], P[0][3]);
                if (TRAJ) {
                    double* tp = a.traj + (int64_t)t * 10 * a.traj_ld + i;
                    tp[0] = cur.mx; tp[a.traj_ld] = cur.my;
#pragma unroll
                    for (int f = 0; f < 4; ++f) { tp[(2 + f) * a.traj_ld] = cur.S[f]; tp[(6 + f) * a.traj_ld] = cur.L[f]; }
                }
                ++t;
                if (t == steps) r = steps;
            }
            if (r >= 0) {
                a.out[i] = cur.mx; a.out[a.ldo + i] = cur.my;
#pragma unroll
                for (int f = 0; f < 4; ++f) { a.out[(2 + f) * a.ldo + i] = cur.S[f]; a.out[(6 + f) * a.ldo + i] = cur.L[f]; }
                a.valid[i] = r;
                if (EXPAND) {
                    const double d0 = x0 - cur.mx, d1 = y0 - cur.my;
                    a.cost[i] = a.parent_cost[i] + sqrt(d0 * d0 + d1 * d1);
                    double gd;
                    const bool g = goal_ok(a.goal, sp, cur.mx, cur.my, cur.S[0] + cur.L[0], cur.S[1] + cur.L[1], cur.S[2] + cur.L[2],
                                           cur.S[3] + cur.L[3], &gd);
                    a.goal_dist[i] = gd;
                    a.flags[i] = (uint8_t)((r == steps ? 1 : 0) | (cons_ok ? 2 : 0) | (g ? 4 : 0));
                }
                have = false;
            }
        }
    }
}

}  // namespace cck
