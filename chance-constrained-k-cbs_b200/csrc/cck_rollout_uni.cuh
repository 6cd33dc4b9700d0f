// cck_rollout_uni.cuh — second generation of the fused propagateWhileValid kernel for the
// Uncertain-Unicycle model (UncertainUnicycleSP.cpp:146-227).
//
// The first version (one thread per belief, static assignment, generic per-obstacle test, dense 4x4 algebra: git
// history) spent ~2000 + ~1100 FP64 instructions per belief-step at 255 registers with spills.  This one
//  * checks PCCBlackmore obstacles with the conservative fp32 box/cell-grid broad phase of
//    cck_rollout_grid.cuh (exact fp64 decision for whatever is not proven: verdicts unchanged),
//  * exploits the EXACT zeros of the model matrices (F = I + dt(e0e1^T + e2e3^T), A_cl_d block
//    diagonal, Q/R diagonal).  A skipped term is 0*x, which is an exact zero for finite x, and
//    adding an exact zero changes nothing: results are bit-identical to the dense evaluation in the
//    reference's accumulation order.  Beliefs with non-finite covariance entries (where 0*inf = NaN
//    would matter) take the dense step,
//  * keeps the last valid state in shared memory instead of a second 64-register copy, and refills
//    a lane from a CTA-wide work counter as soon as its rollout ends (ragged step counts, deaths).
#pragma once
#include "cck_rollout_grid.cuh"

namespace cck {

// Sigma' , Lambda' in place; bit-identical to uni_cov_step for finite inputs when the parameter
// matrices have the reference's sparsity (checked on the host, cck_set_model).
__device__ __forceinline__ void uni_cov_step_sparse(double (&S)[16], double (&L)[16], const UniDev& p) {
    double lp[16];
    {   // Lambda_pred = A_cl_d * Lambda * A_cl_d^T   (:217)
        double t[16];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            t[0 * 4 + j] = p.Acl[0] * L[0 * 4 + j] + p.Acl[1] * L[1 * 4 + j];
            t[1 * 4 + j] = p.Acl[4] * L[0 * 4 + j] + p.Acl[5] * L[1 * 4 + j];
            t[2 * 4 + j] = p.Acl[10] * L[2 * 4 + j] + p.Acl[11] * L[3 * 4 + j];
            t[3 * 4 + j] = p.Acl[14] * L[2 * 4 + j] + p.Acl[15] * L[3 * 4 + j];
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            lp[i * 4 + 0] = t[i * 4 + 0] * p.Acl[0] + t[i * 4 + 1] * p.Acl[1];
            lp[i * 4 + 1] = t[i * 4 + 0] * p.Acl[4] + t[i * 4 + 1] * p.Acl[5];
            lp[i * 4 + 2] = t[i * 4 + 2] * p.Acl[10] + t[i * 4 + 3] * p.Acl[11];
            lp[i * 4 + 3] = t[i * 4 + 2] * p.Acl[14] + t[i * 4 + 3] * p.Acl[15];
        }
    }
    double sp[16];
    {   // Sigma_pred = F * Sigma * F + Q   (:213, no transpose: quirk 2)
        const double f01 = p.F[1], f23 = p.F[11];
        double t[16];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            t[0 * 4 + j] = S[0 * 4 + j] + f01 * S[1 * 4 + j];
            t[1 * 4 + j] = S[1 * 4 + j];
            t[2 * 4 + j] = S[2 * 4 + j] + f23 * S[3 * 4 + j];
            t[3 * 4 + j] = S[3 * 4 + j];
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            sp[i * 4 + 0] = t[i * 4 + 0];
            sp[i * 4 + 1] = t[i * 4 + 0] * f01 + t[i * 4 + 1];
            sp[i * 4 + 2] = t[i * 4 + 2];
            sp[i * 4 + 3] = t[i * 4 + 2] * f23 + t[i * 4 + 3];
        }
#pragma unroll
        for (int d = 0; d < 4; ++d) sp[d * 5] = sp[d * 5] + p.Q[d * 5];
    }
    double K[16];
    {
        double Sf[16], Si[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) Sf[i] = sp[i];
#pragma unroll
        for (int d = 0; d < 4; ++d) Sf[d * 5] = sp[d * 5] + p.R[d * 5];   // :215
        inv4(Sf, Si);
        mm4(sp, Si, K);                                                     // :216
    }
    {
        double KP[16];
        mm4(K, sp, KP);                                                     // :220
#pragma unroll
        for (int i = 0; i < 16; ++i) L[i] = lp[i] + KP[i];
    }
    double ImK[16];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) ImK[i * 4 + j] = (i == j ? 1.0 : 0.0) - K[i * 4 + j];
    mm4(ImK, sp, S);                                                        // :219
}

// dense step through local memory (cold path: non-finite covariances)
__device__ __noinline__ void uni_cov_step_dense(double* SL, const UniDev& p) {
    double So[16], Lo[16];
    uni_cov_step(SL, SL + 16, p, So, Lo);
    for (int i = 0; i < 16; ++i) { SL[i] = So[i]; SL[16 + i] = Lo[i]; }
}

#define CCK_UNI_THREADS 128
// shared memory: [mbarrier 16][table][work counter 16][stash 36 planes x 128 doubles]
__host__ __device__ inline int uni_smem_bytes(int tab_bytes) { return 16 + tab_bytes + 16 + 36 * CCK_UNI_THREADS * 8; }

template <int KIND, bool TRAJ, bool EXPAND>
__global__ void __launch_bounds__(CCK_UNI_THREADS, 2) k_rollout_uni2(const __grid_constant__ RolloutArgs a, const __grid_constant__ UniDev p,
                                                                    const __grid_constant__ SvcDev sp, const __grid_constant__ GridConst g,
                                                                    const int range_log2) {
    // KIND == Blackmore: a.ftab is the grid table (staged), a.tab the class table (global, exact path only);
    // other kinds: a.ftab == a.tab is the class table (staged)
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    unsigned char* tab = smem_raw + 16;
    int* next_k = reinterpret_cast<int*>(smem_raw + 16 + a.ftab_bytes);
    double* stash = reinterpret_cast<double*>(smem_raw + 16 + a.ftab_bytes + 16) + threadIdx.x;   // plane f at stash[f*128]
    if (threadIdx.x == 0) *next_k = 0;
    const unsigned char* cons = a.cons;
    if (EXPAND && a.cons_smem_bytes > 0) {   // small constraint table: shared memory copy (see k_rollout_lin_grid)
        unsigned char* sc = smem_raw + uni_smem_bytes(a.ftab_bytes);
        for (int o = threadIdx.x * 16; o < a.cons_smem_bytes; o += CCK_UNI_THREADS * 16)
            *reinterpret_cast<uint4*>(sc + o) = *reinterpret_cast<const uint4*>(a.cons + o);
        cons = sc;
    }
    stage_begin(tab, a.ftab, (uint32_t)a.ftab_bytes, bar);
    stage_wait((uint32_t)a.ftab_bytes, bar);
    const double c = sp.erf_inv_result;
    const bool has_obstacles = reinterpret_cast<const TabHeader*>(a.tab)->M > 0;

    bool have = false, done = false, dense = false, pending = false;
    bool use_tab = false;          // EXPAND: covariances from the tree's covariance table (RolloutArgs::cov_tab)
    double v[4], u[4], S[16], L[16], v0[4];
    double sr = 0, cr = 1, pcost = 0;
    int steps = 0, t = 0, pstep = 0;
    int64_t i = 0;
    bool cons_ok = true;
#pragma unroll
    for (int f = 0; f < 4; ++f) v[f] = u[f] = v0[f] = 0;
#pragma unroll
    for (int f = 0; f < 16; ++f) S[f] = L[f] = 0;

    auto finish = [&](int r) {
#pragma unroll
        for (int f = 0; f < 4; ++f) a.out[f * a.ldo + i] = v[f];
#pragma unroll
        for (int f = 0; f < 16; ++f) { a.out[(4 + f) * a.ldo + i] = S[f]; a.out[(20 + f) * a.ldo + i] = L[f]; }
        a.valid[i] = r;
        if (EXPAND) {
            double acc = 0;
#pragma unroll
            for (int f = 0; f < 4; ++f) { const double d = v0[f] - v[f]; acc = acc + d * d; }
            a.cost[i] = pcost + sqrt(acc);
            double gd;
            const bool g = goal_ok(a.goal, sp, v[0], v[1], S[0] + L[0], S[1] + L[1], S[4] + L[4], S[5] + L[5], &gd);
            a.goal_dist[i] = gd;
            a.flags[i] = (uint8_t)((r == steps ? 1 : 0) | (cons_ok ? 2 : 0) | (g ? 4 : 0) | (use_tab && steps > 0 ? 8 : 0));
        }
    };

    // store the result of a finished rollout: a lane that is idle keeps its registers and its stash until then
    auto flush = [&]() {
        if (!pending) return;
        pending = false;
        if (t < steps) {           // died: the result is the last valid state
#pragma unroll
            for (int f = 0; f < 4; ++f) v[f] = stash[f * CCK_UNI_THREADS];
#pragma unroll
            for (int f = 0; f < 16; ++f) { S[f] = stash[(4 + f) * CCK_UNI_THREADS]; L[f] = stash[(20 + f) * CCK_UNI_THREADS]; }
        }
        finish(t);
    };
    // refill: the CTA owns ranges blockIdx, blockIdx+gridDim, ... of 2^range_log2 beliefs
    auto take = [&]() {
        flush();
        while (!have && !done) {
            const int kk = atomicAdd(next_k, 1);
            const int rg = kk >> range_log2;
            const int64_t base = ((int64_t)blockIdx.x + (int64_t)rg * gridDim.x) << range_log2;
            if (base >= a.n) { done = true; break; }
            i = base + (kk & ((1 << range_log2) - 1));
            if (i >= a.n) continue;
#pragma unroll
            for (int f = 0; f < 4; ++f) { v[f] = a.bel[f * a.ld + i]; u[f] = a.ctrl[f * a.ldc + i]; }
#pragma unroll
            for (int f = 0; f < 16; ++f) { S[f] = a.bel[(4 + f) * a.ld + i]; L[f] = a.bel[(20 + f) * a.ld + i]; }
            steps = a.steps[i];
            if (EXPAND) { pcost = a.parent_cost[i]; pstep = a.parent_step[i]; }
#pragma unroll
            for (int f = 0; f < 4; ++f) v0[f] = v[f];
            cons_ok = true;
            if (steps <= 0) { steps = 0; finish(0); continue; }
            sincos(u[2], &sr, &cr);
            double mag = 0;
#pragma unroll
            for (int f = 0; f < 16; ++f) mag = mag + (fabs(S[f]) + fabs(L[f]));
            dense = !p.sparse || !(mag < 1e300);      // non-reference sparsity, or a non-finite (or absurd) covariance: 0*x is not an exact zero any more
            if (EXPAND) {          // parent covariance bit-identical to the table's entry for its depth, and the segment inside the table
                use_tab = false;
                if (a.cov_tab && pstep >= 0 && pstep + steps < a.cov_tab_n) {
                    const double* e = a.cov_tab + (size_t)pstep * 32;
                    bool same = true;
#pragma unroll
                    for (int f = 0; f < 16; ++f)
                        same = same && __double_as_longlong(S[f]) == __double_as_longlong(__ldg(e + f)) &&
                               __double_as_longlong(L[f]) == __double_as_longlong(__ldg(e + 16 + f));
                    use_tab = same;
                }
            }
            t = 0;
            have = true;
        }
    };
    // Finished lanes are served in batches (see k_rollout_lin_grid): the result store (36 planes) and the reload are
    // ~400 instructions that one finished lane would make the whole warp issue.
    int it = 0;
    while (true) {
        const unsigned idle = __ballot_sync(0xffffffffu, !have);
        if (idle) {
            if (__popc(idle) >= CCK_REFILL_MIN || (++it & (CCK_REFILL_WAIT - 1)) == 0) {
                take();
                if (__all_sync(0xffffffffu, !have)) break;
            }
        }
        if (!have) continue;
        // stash the last valid state, advance in place
#pragma unroll
        for (int f = 0; f < 4; ++f) stash[f * CCK_UNI_THREADS] = v[f];
#pragma unroll
        for (int f = 0; f < 16; ++f) { stash[(4 + f) * CCK_UNI_THREADS] = S[f]; stash[(20 + f) * CCK_UNI_THREADS] = L[f]; }
        {
            double vn[4];
            uni_mean_step(v, u, cr, sr, a.duration, p, vn);
#pragma unroll
            for (int f = 0; f < 4; ++f) v[f] = vn[f];
        }
        if (EXPAND && use_tab) {   // the covariances of absolute step pstep + t + 1, as k_cov_table_uni computed them
            const double2* e = reinterpret_cast<const double2*>(a.cov_tab + (size_t)(pstep + t + 1) * 32);
#pragma unroll
            for (int f = 0; f < 8; ++f) {
                const double2 sv = __ldg(e + f), lv = __ldg(e + 8 + f);
                S[2 * f] = sv.x; S[2 * f + 1] = sv.y; L[2 * f] = lv.x; L[2 * f + 1] = lv.y;
            }
        } else if (!dense) {
            uni_cov_step_sparse(S, L, p);
        } else {
            double SL[32];
#pragma unroll
            for (int f = 0; f < 16; ++f) { SL[f] = S[f]; SL[16 + f] = L[f]; }
            uni_cov_step_dense(SL, p);
#pragma unroll
            for (int f = 0; f < 16; ++f) { S[f] = SL[f]; L[f] = SL[16 + f]; }
        }
        const double P0 = S[0] + L[0], P1 = S[1] + L[1], P2 = S[4] + L[4], P3 = S[5] + L[5];   // (x, xdot) block, quirk 2
        bool ok = bounds_ok(sp, v);
        if (ok && has_obstacles) {
            if (KIND == CCK_SVC_BLACKMORE) ok = grid_check(g, tab, a.tab, a.exB, a.exCls, c, v[0], v[1], P0, P1, P2, P3);
            else if (KIND == CCK_SVC_ADAPTIVE) ok = halfplane_check<true>(tab, v[0], v[1], P0, P1, P2, P3, adaptive_c(sp, tab, v[0], v[1]));
            else ok = chisq_check(sp, tab, v[0], v[1], P0, P1, P2, P3);
        }
        if (!ok) {
            have = false; pending = true;      // died with t valid steps; flush() restores the last valid state from the stash
        } else {
            if (EXPAND && cons) {
                // getDistribution_ picks entries (0,0),(0,2),(2,0),(2,2) (BeliefPVC.cpp:94-97)
                cons_ok = cons_ok && constraints_ok_at(cons, pstep + t + 1, v[0], v[1], S[0] + L[0], S[2] + L[2], S[8] + L[8], S[10] + L[10]);
            }
            if (TRAJ && t < a.traj_T) {
                double* tp = a.traj + (int64_t)t * 36 * a.traj_ld + i;
#pragma unroll
                for (int f = 0; f < 4; ++f) tp[f * a.traj_ld] = v[f];
#pragma unroll
                for (int f = 0; f < 16; ++f) { tp[(4 + f) * a.traj_ld] = S[f]; tp[(20 + f) * a.traj_ld] = L[f]; }
            }
            ++t;
            if (t == steps) { have = false; pending = true; }
        }
    }
}

}  // namespace cck
