// cck_rollout_bm.cuh — the fused propagateWhileValid kernel for the default configuration
// (2DUncertainLinear + PCCBlackmoreSVC), restructured around what ncu showed on the first
// version (profiles/r1a_*): the per-obstacle loop was co-limited by LDS broadcast bandwidth
// (2 x LDS.128 per obstacle per warp) and by DADD->DSETP dependency chains.
//
//  * per class of identical normals the four inequalities  a_i.mu >= fl(B_oi + bb_i)  are
//    turned into four thresholds tau_i with a conservative margin: B_oi <= tau_i  PROVES the
//    reference's inequality holds (fl is monotone), so the common case is 4 DSETP per obstacle
//    and no DADD; anything not proven safe (a real collision, or the ~1e-12-wide ambiguous
//    band) is re-evaluated with the reference's exact expression.  Verdicts stay bit-identical.
//  * BPT beliefs per thread share every table row loaded from shared memory, dividing the
//    LDS traffic per belief by BPT and giving BPT x 4 independent compare chains.
//  * only the current state lives in registers; a belief that dies at step r is re-played
//    (r propagate steps, no checks: ~60 flop each) to produce its last valid state, instead of
//    carrying a second copy of (mean, Sigma, Lambda) through the whole rollout.
#pragma once
#include "cck_kernels.cuh"

namespace cck {

// exact reference test of one obstacle (slow path; PCCBlackmoreSVC.cpp:63-75)
__device__ __noinline__ bool bm_exact_obstacle(const double* A8, const double2* Brow, double x, double y, double P0, double P1,
                                               double P2, double P3, double c) {
    const double B[4] = {Brow[0].x, Brow[0].y, Brow[1].x, Brow[1].y};
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
__global__ void __launch_bounds__(128) k_rollout_lin_bm(const __grid_constant__ RolloutArgs a, const __grid_constant__ LinDev p,
                                                        const __grid_constant__ SvcDev sp) {
    CCK_SMEM_PROLOGUE(a)
    const TabHeader* h = reinterpret_cast<const TabHeader*>(tab);
    bool waited = false;
    // vector (16-byte) plane accesses need even leading dimensions and 16-byte aligned bases
    const bool vec = BPT > 1 && ((a.ld | a.ldc | a.ldo | a.traj_ld) & 1) == 0 &&
                     (((uintptr_t)a.bel | (uintptr_t)a.ctrl | (uintptr_t)a.out | (uintptr_t)a.traj) & 15) == 0;
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
        if (!waited) { stage_wait((uint32_t)a.tab_bytes, bar); waited = true; }

        for (int t = 0; t < tmax; ++t) {
            bool any = false;
#pragma unroll
            for (int b = 0; b < BPT; ++b) any = any || alive[b];
            if (!any) break;
            bool ok[BPT];
#pragma unroll
            for (int b = 0; b < BPT; ++b) {
                if (alive[b] && t >= steps[b]) alive[b] = false;      // finished all of its steps
                if (alive[b]) lin_step(cur[b], dux[b], duy[b], p, cur[b]);
                // satisfiesBounds (NaN passes)
                ok[b] = !(cur[b].mx - eps > sp.high[0] || cur[b].mx + eps < sp.low[0] || cur[b].my - eps > sp.high[1] ||
                          cur[b].my + eps < sp.low[1]);
            }
            const ClassRec* cls = reinterpret_cast<const ClassRec*>(tab + h->off_classes);
            const double2* Bt = reinterpret_cast<const double2*>(tab + h->off_B);
            const int nc = h->n_classes;
            for (int k = 0; k < nc; ++k) {
                double tau[BPT][4];
#pragma unroll
                for (int b = 0; b < BPT; ++b) {
                    const double P0 = cur[b].S[0] + cur[b].L[0], P1 = cur[b].S[1] + cur[b].L[1];
                    const double P2 = cur[b].S[2] + cur[b].L[2], P3 = cur[b].S[3] + cur[b].L[3];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const double a0 = cls[k].A[2 * i], a1 = cls[k].A[2 * i + 1];
                        const double t0 = a0 * P0 + a1 * P2;
                        const double t1 = a0 * P1 + a1 * P3;
                        const double tmp = t0 * a0 + t1 * a1;
                        const double PV = sqrt(tmp);
                        const double bb = 1.4142135623730951 * PV * c;
                        const double lhs = cur[b].mx * a0 + cur[b].my * a1;
                        // B <= tau  =>  fl(B + bb) <= lhs : margin 2^-40 * (|lhs|+|bb|) >> rounding (2^-52 relative);
                        // NaN/inf operands give tau = NaN -> nothing is "proven", the exact path decides.
                        tau[b][i] = alive[b] ? (lhs - bb) - (fabs(lhs) + fabs(bb)) * 9.094947017729282e-13 : HUGE_VAL;
                    }
                }
                const int first = cls[k].first, cnt = cls[k].count, cntp = cls[k].count_pad;
                // branch-free over chunks of 4 table rows (padding rows are -inf: always proven)
                for (int o = 0; o < cntp; o += 4) {
                    bool pr[BPT];
#pragma unroll
                    for (int b = 0; b < BPT; ++b) pr[b] = true;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const double2 b01 = Bt[(first + o + j) * 2], b23 = Bt[(first + o + j) * 2 + 1];
#pragma unroll
                        for (int b = 0; b < BPT; ++b)
                            pr[b] = pr[b] & ((b01.x <= tau[b][0]) | (b01.y <= tau[b][1]) | (b23.x <= tau[b][2]) | (b23.y <= tau[b][3]));
                    }
#pragma unroll
                    for (int b = 0; b < BPT; ++b) {
                        if (!pr[b] && ok[b]) {   // rare: a real collision or the ambiguous band -> exact reference test
                            for (int j = 0; j < 4 && o + j < cnt; ++j)
                                ok[b] = ok[b] && bm_exact_obstacle(cls[k].A, &Bt[(first + o + j) * 2], cur[b].mx, cur[b].my,
                                                                   cur[b].S[0] + cur[b].L[0], cur[b].S[1] + cur[b].L[1],
                                                                   cur[b].S[2] + cur[b].L[2], cur[b].S[3] + cur[b].L[3], c);
                        }
                    }
                }
            }
#pragma unroll
            for (int b = 0; b < BPT; ++b) {
                if (!alive[b]) continue;
                if (!ok[b]) { alive[b] = false; r[b] = t; continue; }
                if (EXPAND && a.cons)
                    cons_ok[b] = cons_ok[b] && constraints_ok_at(a.cons, pstep[b] + t + 1, cur[b].mx, cur[b].my, cur[b].S[0] + cur[b].L[0],
                                                                 cur[b].S[1] + cur[b].L[1], cur[b].S[2] + cur[b].L[2], cur[b].S[3] + cur[b].L[3]);
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
    if (!waited) stage_wait((uint32_t)a.tab_bytes, bar);
}

}  // namespace cck
