// cck_kernels.cuh — the CUDA kernels behind the C ABI (sm_100a).
//
// Data layout: beliefs are SoA planes in HBM (plane f of belief i at base[f*ld+i]) so a
// warp's loads/stores of one plane are 256 contiguous bytes.  One thread owns one belief
// and keeps (mean, Sigma, Lambda) in registers across all steps of a rollout; the obstacle
// table (classes of identical half-plane normals + per-obstacle offsets B) is staged once
// per CTA into shared memory with a TMA bulk copy that overlaps the belief loads.
#pragma once
#include "cck_device.cuh"

namespace cck {

#define CCK_CONS_SMEM_MAX 16384
struct RolloutArgs {
    int64_t n;
    const double* bel; int64_t ld;
    const double* ctrl; int64_t ldc;
    const int32_t* steps;
    double* out; int64_t ldo;
    int32_t* valid;
    double* traj; int64_t traj_ld; int32_t traj_T;
    const unsigned char* tab; int32_t tab_bytes;
    double duration;   // step size handed to propagate() (= dt inside propagateWhileValid)
    // batched-expansion extras (EXPAND kernels only)
    const int32_t* parent_step; const double* parent_cost;
    double* cost; double* goal_dist; uint8_t* flags;
    cck_goal_params goal;
    // Covariance table of the planner's tree (EXPAND kernels): the covariance recursion does not depend on the mean or the
    // control (UncertainLinearSP.cpp:98-104, UncertainUnicycleSP.cpp:210-220), and every node of a low-level tree descends from
    // the same start covariance, so the node at absolute step s carries cov_tab[s] = (Sigma_s, Lambda_s) - computed once by
    // k_cov_table with the very instruction sequence of the step functions.  A candidate whose parent covariance equals
    // cov_tab[parent_step] BIT FOR BIT reads its next covariances from the table instead of recomputing them; any other
    // candidate takes the normal step.  Entry s at cov_tab + s * 2 * dim * dim (Sigma row-major, then Lambda).
    const double* cov_tab; int32_t cov_tab_n;
    const unsigned char* cons;   // constraint table (global memory)
    int32_t cons_smem_bytes;     // > 0: the table is small, every CTA copies it into shared memory first (multiple of 16)
    // staged table: the fp32 box/cell-grid table (Blackmore; exact rows exB/exCls stay in global memory for the rare
    // exact decision) or the class table (other checkers, unicycle kernel)
    const unsigned char* ftab; int32_t ftab_bytes;
    const double* exB; const int32_t* exCls;
};

// ---------------------------------------------------------------------------
// Constraint table (BeliefConstraint entries flattened, sorted by absolute step)
// ---------------------------------------------------------------------------
struct ConsHeader {
    int32_t n_entries, max_step;       // entries cover steps [0, max_step]
    int32_t off_first, off_entries;    // byte offsets: int32 first[max_step+2]; ConsEntry entries[n]
    int32_t pvc_kind, dim, k, grid_n;  // grid_n: CDFGridPVC's disk_steps_
    double c_pair, p_coll_agnts;
    double sc, pad2;                   // ChiSquaredBoundaryPVC: chi2_2 quantile (ChiSquaredBoundaryPVC.cpp:19)
};
struct ConsEntry {
    double mu[2], cov[4];              // the constraining robot's getDistribution_ at this step
    double A[8], B[4];                 // half-planes of the unordered robot pair
    double rad[2];                     // bounding radii: [0] constrained (planning) robot, [1] constraining robot
};
static_assert(sizeof(ConsHeader) == 64 && sizeof(ConsEntry) == 160, "constraint table records");

__host__ __device__ __forceinline__ bool pair_safe_halfplanes(const double* A, const double* B, double mx, double my, double S0,
                                                     double S1, double S2, double S3, double c) {
    bool safe = false;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const double a0 = A[2 * i], a1 = A[2 * i + 1];
        const double t0 = a0 * S0 + a1 * S2;
        const double t1 = a0 * S1 + a1 * S3;
        const double tmp = t0 * a0 + t1 * a1;
        const double Pv = sqrt(tmp);
        const double vbar = 1.4142135623730951 * Pv * c;
        safe = safe | (a0 * mx + a1 * my - B[i] >= vbar);   // MinkowskiSumBlackmorePVC.cpp:176-181
    }
    return safe;
}

// Blackmore2PVC::isSafe_ (Blackmore2PVC.cpp:169-192): smallest half-plane collision probability against p_coll_agnts_
__host__ __device__ __forceinline__ bool pair_safe_blackmore2(const double* A, const double* B, double mx, double my, double S0, double S1, double S2,
                                                     double S3, double p_coll_agnts) {
    double p_obs = DBL_MAX;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const double a0 = A[2 * i], a1 = A[2 * i + 1];
        const double t0 = a0 * S0 + a1 * S2;
        const double t1 = a0 * S1 + a1 * S3;
        const double pv2 = t0 * a0 + t1 * a1;
        const double PV = sqrt(2 * pv2);
        const double pc = 0.5 + 0.5 * erf((B[i] - (mx * a0 + my * a1)) / PV);
        if (pc < p_obs) p_obs = pc;
    }
    return !(p_obs > p_coll_agnts);
}

// Eigen::EigenSolver<Matrix2d> as CDFGridPVC uses it, in closed form (see oracle/cck_oracle.hpp::eig2_general for the contract)
__host__ __device__ __forceinline__ void eig2_general(const double* M, double* lam, double* V) {
    const double a = M[0], b = M[1], c = M[2], d = M[3];
    V[0] = 1; V[1] = 0; V[2] = 0; V[3] = 1;
    if (b == 0 && c == 0) { lam[0] = a; lam[1] = d; return; }
    const double h = (a - d) / 2, disc = h * h + b * c, m = (a + d) / 2;
    if (!(disc >= 0)) { lam[0] = lam[1] = m; return; }
    const double sq = sqrt(disc);
    lam[0] = m + sq; lam[1] = m - sq;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        double vx, vy;
        if (fabs(b) >= fabs(c)) { vx = b; vy = lam[j] - a; } else { vx = lam[j] - d; vy = c; }
        const double n = sqrt(vx * vx + vy * vy);
        V[0 + j] = vx / n; V[2 + j] = vy / n;
    }
}
__host__ __device__ __forceinline__ double cdf_rect(double x1, double y1, double x2, double y2) {   // CDFGridPVC.cpp:293-306
    const double r2 = 1.4142135623730951;   // sqrt(2)
    const double ex1 = 1 + erf(x1 / r2), ex2 = 1 + erf(x2 / r2), ey1 = 1 + erf(y1 / r2), ey2 = 1 + erf(y2 / r2);
    const double p1 = (1.0 / 2.0) * ex1 * (1.0 / 2.0) * ey1, p2 = (1.0 / 2.0) * ex2 * (1.0 / 2.0) * ey1;
    const double p3 = (1.0 / 2.0) * ex2 * (1.0 / 2.0) * ey2, p4 = (1.0 / 2.0) * ex1 * (1.0 / 2.0) * ey2;
    return p3 - p2 - p4 + p1;
}
// CDFGridPVC::isSafe_ (CDFGridPVC.cpp:182-291); belief a = (mua, ca), belief b = (mub, cb), box half-side r_sum, grid n
__host__ __device__ __forceinline__ bool pair_safe_cdfgrid(const double* mua, const double* ca, const double* mub, const double* cb, double r_sum, int n,
                                                  double p_coll_agnts) {
    const double T0 = mua[0] - mub[0], T1 = mua[1] - mub[1];
    double S[4], lam[2], V[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) S[i] = ca[i] + cb[i];
    eig2_general(S, lam, V);
    const double l0 = sqrt(lam[0]), l1 = sqrt(lam[1]);
    const double invdet = 1.0 / (l0 * l1 - 0.0 * 0.0);
    const double Li00 = l1 * invdet, Li11 = l0 * invdet;
    const double R0 = V[0] * Li00 + V[1] * 0.0, R1 = V[0] * 0.0 + V[1] * Li11, R2 = V[2] * Li00 + V[3] * 0.0, R3 = V[2] * 0.0 + V[3] * Li11;
    double px[4], py[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const double vx = (k == 0 || k == 3) ? -r_sum : r_sum, vy = (k < 2) ? -r_sum : r_sum;
        const double ex = -vx + T0, ey = -vy + T1;
        px[k] = R0 * ex + R2 * ey;
        py[k] = R1 * ex + R3 * ey;
    }
    double x_low = px[0], x_high = px[0], y_low = py[0], y_high = py[0];
#pragma unroll
    for (int k = 1; k < 4; ++k) {   // std::min / std::max semantics (a NaN candidate never replaces the running value)
        x_low = px[k] < x_low ? px[k] : x_low; x_high = x_high < px[k] ? px[k] : x_high;
        y_low = py[k] < y_low ? py[k] : y_low; y_high = y_high < py[k] ? py[k] : y_high;
    }
    double prob = 0;
    if (n == 2) {
        prob += cdf_rect(x_low, y_low, x_high, y_high);
    } else if (n > 2) {
        double area2 = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) { const int k1 = (k + 1) & 3; area2 += px[k] * py[k1] - px[k1] * py[k]; }
        const double sg = area2 >= 0 ? 1.0 : -1.0;
        auto inside = [&](double x, double y) {
            bool in = true;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int k1 = (k + 1) & 3;
                const double cr = (px[k1] - px[k]) * (y - py[k]) - (py[k1] - py[k]) * (x - px[k]);
                if (sg * cr < 0) in = false;
            }
            return in;
        };
        const double dx = (x_high - x_low) / (n - 1), dy = (y_high - y_low) / (n - 1);
        for (int ix = 0; ix != n - 1; ++ix)
            for (int iy = 0; iy != n - 1; ++iy) {
                const double x1 = x_low + dx * ix, x2 = (ix + 1 == n - 1) ? x_high : x_low + dx * (ix + 1);
                const double y1 = y_low + dy * iy, y2 = (iy + 1 == n - 1) ? y_high : y_low + dy * (iy + 1);
                if (inside(x1, y1) || inside(x2, y1) || inside(x2, y2) || inside(x1, y2)) prob += cdf_rect(x1, y1, x2, y2);
            }
    }
    return prob < p_coll_agnts;
}

// All constraint entries that fall on absolute step `s` against the belief (mu, cov).
__device__ __forceinline__ bool constraints_ok_at(const unsigned char* cons, int s, double mx, double my, double C0,
                                                  double C1, double C2, double C3) {
    const ConsHeader* h = reinterpret_cast<const ConsHeader*>(cons);
    if (s < 0 || s > h->max_step) return true;
    const int32_t* first = reinterpret_cast<const int32_t*>(cons + h->off_first);
    const ConsEntry* ent = reinterpret_cast<const ConsEntry*>(cons + h->off_entries);
    bool ok = true;
    for (int e = first[s]; e < first[s + 1] && ok; ++e) {
        const double dx = mx - ent[e].mu[0], dy = my - ent[e].mu[1];
        const double S0 = C0 + ent[e].cov[0], S1 = C1 + ent[e].cov[1], S2 = C2 + ent[e].cov[2], S3 = C3 + ent[e].cov[3];
        if (h->pvc_kind == CCK_PVC_CHISQUARED) {
            // ChiSquaredBoundaryPVC::satisfiesConstraints (ChiSquaredBoundaryPVC.cpp:54-87) -> isSafe_ (:139-168):
            // belief_a = the planning robot's state, belief_b = the stored state of the constraining robot
            const double la = max_eig_real_2x2(C0, C1, C2, C3), lb = max_eig_real_2x2(ent[e].cov[0], ent[e].cov[1], ent[e].cov[2], ent[e].cov[3]);
            const double dist = sqrt(dx * dx + dy * dy);
            const double boundary = (la * h->sc + ent[e].rad[0]) + (lb * h->sc + ent[e].rad[1]);
            ok = dist > boundary;
            continue;
        }
        if (h->pvc_kind == CCK_PVC_BLACKMORE2) {      // Blackmore2PVC.cpp:64-104
            ok = pair_safe_blackmore2(ent[e].A, ent[e].B, dx, dy, S0, S1, S2, S3, h->p_coll_agnts);
            continue;
        }
        if (h->pvc_kind == CCK_PVC_CDFGRID) {         // CDFGridPVC.cpp:72-109: the planning robot's belief first
            const double mua[2] = {mx, my}, ca[4] = {C0, C1, C2, C3};
            ok = pair_safe_cdfgrid(mua, ca, ent[e].mu, ent[e].cov, ent[e].rad[0] + ent[e].rad[1], h->grid_n, h->p_coll_agnts);
            continue;
        }
        double c = h->c_pair;
        if (h->pvc_kind == CCK_PVC_ADAPTIVE_BLACKMORE || h->pvc_kind == CCK_PVC_ADAPTIVE_BOUNDINGBOX) {
            // states_map holds only the two robots (AdaptiveRiskBlackmorePVC.cpp:92-101)
            const double di = sqrt(dx * dx + dy * dy);
            double sum = 0;
            sum += (1 / di);
            const double alpha = 1 / sum;
            c = erfinv(1 - (2 * ((alpha / di) * h->p_coll_agnts)));
        }
        ok = pair_safe_halfplanes(ent[e].A, ent[e].B, dx, dy, S0, S1, S2, S3, c);
    }
    return ok;
}

// ChanceConstrainedGoal::isSafe_ (BeliefSpaceGoals.cpp:90-147): decagon of radius
// lambda_max*sc inside the goal decagon of radius `threshold`.
__device__ __forceinline__ bool goal_ok(const cck_goal_params& g, const SvcDev& sp, double x, double y, double P0, double P1,
                                        double P2, double P3, double* dist) {
    const double dx = x - g.gx, dy = y - g.gy;
    *dist = (dx * dx + dy * dy);
    const double max_lambda = max_eig_real_2x2(P0, P1, P2, P3);
    const double boundary = (max_lambda * g.sc + 0.0);
    double gx[11], gy[11];
#pragma unroll
    for (int i = 0; i < 10; ++i) { gx[i] = g.gx + g.threshold * sp.dec_cos[i]; gy[i] = g.gy + g.threshold * sp.dec_sin[i]; }
    gx[10] = gx[0]; gy[10] = gy[0];
    double area2 = 0;
#pragma unroll
    for (int i = 0; i < 10; ++i) area2 += gx[i] * gy[i + 1] - gx[i + 1] * gy[i];
    const double sgn = area2 >= 0 ? 1.0 : -1.0;
    bool inside = true;
    for (int j = 0; j < 10 && inside; ++j) {
        const double px = x + boundary * sp.dec_cos[j], py = y + boundary * sp.dec_sin[j];
#pragma unroll
        for (int i = 0; i < 10; ++i) {
            const double cr = (gx[i + 1] - gx[i]) * (py - gy[i]) - (gy[i + 1] - gy[i]) * (px - gx[i]);
            if (sgn * cr < 0) inside = false;
        }
    }
    return inside;
}

// ---------------------------------------------------------------------------
// shared-memory carve-up: [mbarrier (16 B)] [obstacle table]
// ---------------------------------------------------------------------------
#define CCK_SMEM_PROLOGUE(a)                                                        \
    extern __shared__ __align__(128) unsigned char smem_raw[];                      \
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);                          \
    unsigned char* tab = smem_raw + 16;                                             \
    stage_begin(tab, (a).tab, (uint32_t)(a).tab_bytes, bar);

// ---------------------------------------------------------------------------
// StatePropagator::propagate, batched, one step
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_propagate_lin(int64_t n, const double* __restrict__ bel, int64_t ld,
                                                       const double* __restrict__ ctrl, int64_t ldc,
                                                       double* __restrict__ out, int64_t ldo,
                                                       const __grid_constant__ LinDev p) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        Lin s, o;
        s.mx = bel[i]; s.my = bel[ld + i];
#pragma unroll
        for (int f = 0; f < 4; ++f) { s.S[f] = bel[(2 + f) * ld + i]; s.L[f] = bel[(6 + f) * ld + i]; }
        const double dux = p.dt * ctrl[i], duy = p.dt * ctrl[ldc + i];   // quirk 1: member duration_, not the argument
        lin_step(s, dux, duy, p, o);
        out[i] = o.mx; out[ldo + i] = o.my;
#pragma unroll
        for (int f = 0; f < 4; ++f) { out[(2 + f) * ldo + i] = o.S[f]; out[(6 + f) * ldo + i] = o.L[f]; }
    }
}

__global__ void __launch_bounds__(128) k_propagate_uni(int64_t n, const double* __restrict__ bel, int64_t ld,
                                                       const double* __restrict__ ctrl, int64_t ldc, double duration,
                                                       double* __restrict__ out, int64_t ldo,
                                                       const __grid_constant__ UniDev p) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double v[4], u[4], S[16], L[16], vo[4], So[16], Lo[16];
#pragma unroll
        for (int f = 0; f < 4; ++f) { v[f] = bel[f * ld + i]; u[f] = ctrl[f * ldc + i]; }
#pragma unroll
        for (int f = 0; f < 16; ++f) { S[f] = bel[(4 + f) * ld + i]; L[f] = bel[(20 + f) * ld + i]; }
        double sr, cr;
        sincos(u[2], &sr, &cr);
        uni_mean_step(v, u, cr, sr, duration, p, vo);
        uni_cov_step(S, L, p, So, Lo);
#pragma unroll
        for (int f = 0; f < 4; ++f) out[f * ldo + i] = vo[f];
#pragma unroll
        for (int f = 0; f < 16; ++f) { out[(4 + f) * ldo + i] = So[f]; out[(20 + f) * ldo + i] = Lo[f]; }
    }
}

// ---------------------------------------------------------------------------
// StateValidityChecker::isValid, batched
// ---------------------------------------------------------------------------
template <int KIND>
__global__ void __launch_bounds__(256) k_is_valid(int64_t n, int dim, const double* __restrict__ bel, int64_t ld,
                                                  uint8_t* __restrict__ valid, const unsigned char* __restrict__ gtab,
                                                  int32_t tab_bytes, const __grid_constant__ SvcDev sp) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    unsigned char* tab = smem_raw + 16;
    stage_begin(tab, gtab, (uint32_t)tab_bytes, bar);
    bool waited = false;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double v[4] = {0, 0, 0, 0};
        const int dd = dim * dim;
        for (int f = 0; f < dim; ++f) v[f] = bel[f * ld + i];
        // (Sigma+Lambda).block<2,2>(0,0)
        const double P0 = bel[(dim + 0) * ld + i] + bel[(dim + dd + 0) * ld + i];
        const double P1 = bel[(dim + 1) * ld + i] + bel[(dim + dd + 1) * ld + i];
        const double P2 = bel[(dim + dim) * ld + i] + bel[(dim + dd + dim) * ld + i];
        const double P3 = bel[(dim + dim + 1) * ld + i] + bel[(dim + dd + dim + 1) * ld + i];
        if (!waited) { stage_wait((uint32_t)tab_bytes, bar); waited = true; }
        valid[i] = svc_valid<KIND>(sp, tab, v, P0, P1, P2, P3) ? 1 : 0;
    }
    if (!waited) stage_wait((uint32_t)tab_bytes, bar);   // never exit with a copy in flight
}

// ---------------------------------------------------------------------------
// propagateWhileValid, batched: the fused rollout kernel (linear model)
//   EXPAND adds the constraint check on the new segment, the path-length cost and the
//   goal test (ConstraintRespectingBSST.cpp:250-326) to the same launch.
// ---------------------------------------------------------------------------
template <int KIND, bool TRAJ, bool EXPAND>
__global__ void __launch_bounds__(128) k_rollout_lin(const __grid_constant__ RolloutArgs a, const __grid_constant__ LinDev p,
                                                     const __grid_constant__ SvcDev sp) {
    CCK_SMEM_PROLOGUE(a)
    bool waited = false;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * blockDim.x) {
        Lin cur, nxt;
        cur.mx = a.bel[i]; cur.my = a.bel[a.ld + i];
#pragma unroll
        for (int f = 0; f < 4; ++f) { cur.S[f] = a.bel[(2 + f) * a.ld + i]; cur.L[f] = a.bel[(6 + f) * a.ld + i]; }
        const double dux = p.dt * a.ctrl[i], duy = p.dt * a.ctrl[a.ldc + i];
        const int steps = a.steps[i];
        const double x0 = cur.mx, y0 = cur.my;
        int pstep = 0;
        bool cons_ok = true;
        if (EXPAND) pstep = a.parent_step[i];
        if (!waited) { stage_wait((uint32_t)a.tab_bytes, bar); waited = true; }
        int r = steps > 0 ? steps : 0;
        for (int t = 0; t < steps; ++t) {
            lin_step(cur, dux, duy, p, nxt);
            const double v[4] = {nxt.mx, nxt.my, 0, 0};
            const double P0 = nxt.S[0] + nxt.L[0], P1 = nxt.S[1] + nxt.L[1], P2 = nxt.S[2] + nxt.L[2], P3 = nxt.S[3] + nxt.L[3];
            if (!svc_valid<KIND>(sp, tab, v, P0, P1, P2, P3)) { r = t; break; }
            cur = nxt;
            if (EXPAND && a.cons) cons_ok = cons_ok && constraints_ok_at(a.cons, pstep + t + 1, nxt.mx, nxt.my, P0, P1, P2, P3);
            if (TRAJ && t < a.traj_T) {   // device-pointer callers: steps beyond the buffer are not recorded
                double* tp = a.traj + (int64_t)t * 10 * a.traj_ld + i;
                tp[0] = nxt.mx; tp[a.traj_ld] = nxt.my;
#pragma unroll
                for (int f = 0; f < 4; ++f) { tp[(2 + f) * a.traj_ld] = nxt.S[f]; tp[(6 + f) * a.traj_ld] = nxt.L[f]; }
            }
        }
        a.out[i] = cur.mx; a.out[a.ldo + i] = cur.my;
#pragma unroll
        for (int f = 0; f < 4; ++f) { a.out[(2 + f) * a.ldo + i] = cur.S[f]; a.out[(6 + f) * a.ldo + i] = cur.L[f]; }
        a.valid[i] = r;
        if (EXPAND) {
            // EuclideanPathLengthObjective::motionCost + combineCosts (StateCostObjectives.h:17-26)
            const double d0 = x0 - cur.mx, d1 = y0 - cur.my;
            a.cost[i] = a.parent_cost[i] + sqrt(d0 * d0 + d1 * d1);
            double gd;
            const bool g = goal_ok(a.goal, sp, cur.mx, cur.my, cur.S[0] + cur.L[0], cur.S[1] + cur.L[1], cur.S[2] + cur.L[2],
                                   cur.S[3] + cur.L[3], &gd);
            a.goal_dist[i] = gd;
            a.flags[i] = (uint8_t)((r == steps ? 1 : 0) | (cons_ok ? 2 : 0) | (g ? 4 : 0));
        }
    }
    if (!waited) stage_wait((uint32_t)a.tab_bytes, bar);
}

// ---------------------------------------------------------------------------
// covariance table of the planner's tree (see RolloutArgs::cov_tab): one thread iterates the model's own step
// ---------------------------------------------------------------------------
__global__ void k_cov_table_lin(double* __restrict__ tab, int n, const __grid_constant__ LinDev p, const double* __restrict__ root) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    Lin s, o;
    s.mx = s.my = 0;
    for (int f = 0; f < 4; ++f) { s.S[f] = root[f]; s.L[f] = root[4 + f]; }
    for (int k = 0; k < n; ++k) {
        for (int f = 0; f < 4; ++f) { tab[(size_t)k * 8 + f] = s.S[f]; tab[(size_t)k * 8 + 4 + f] = s.L[f]; }
        // the diagonal-constant step where the rollout kernels would take it (finite covariance), else the dense one: the
        // two agree except for the sign of an exact zero (cck_device.cuh), and the table must hold what the kernels compute
        const bool diag = p.diag && cov_small(s.S[0] + s.L[0], s.S[1] + s.L[1], s.S[2] + s.L[2], s.S[3] + s.L[3]);
        if (diag) lin_step_diag(s, 0.0, 0.0, p, o); else lin_step(s, 0.0, 0.0, p, o);
        s = o;
    }
}
__global__ void k_cov_table_uni(double* __restrict__ tab, int n, const __grid_constant__ UniDev p, const double* __restrict__ root) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    double S[16], L[16], So[16], Lo[16];
    for (int f = 0; f < 16; ++f) { S[f] = root[f]; L[f] = root[16 + f]; }
    for (int k = 0; k < n; ++k) {
        for (int f = 0; f < 16; ++f) { tab[(size_t)k * 32 + f] = S[f]; tab[(size_t)k * 32 + 16 + f] = L[f]; }
        uni_cov_step(S, L, p, So, Lo);     // the dense step: bit-identical to the sparse one for finite inputs (cck_rollout_uni.cuh)
        for (int f = 0; f < 16; ++f) { S[f] = So[f]; L[f] = Lo[f]; }
    }
}

// ---------------------------------------------------------------------------
// PlanValidityChecker::validatePlan
// ---------------------------------------------------------------------------
struct PvcDev {
    int32_t kind, k, dim, n_pairs, grid_n, pad;
    double c_pair, p_coll_agnts, sc;
    const double* radii;       // [k]
    const double* pair_A;      // [k*k*8]
    const double* pair_B;      // [k*k*4]
    const int32_t* map_order;  // [k]
};

// getDistribution_ of robot r at (time-aligned) step t (BeliefPVC.cpp:42-113)
__device__ __forceinline__ void pvc_dist(const PvcDev& pv, const double* traj, int64_t ld, const int32_t* lens, int r, int t,
                                         double* mu, double* cov) {
    const int W = pv.dim + 2 * pv.dim * pv.dim;
    const int len = lens[r];
    const int tt = t < len ? t : len - 1;                     // robots past their end hold their last state
    const double* b = traj + (int64_t)r * W * ld + tt;
    mu[0] = b[0]; mu[1] = b[ld];
    if (pv.dim == 4) {
        cov[0] = b[(4 + 0) * ld] + b[(20 + 0) * ld];   cov[1] = b[(4 + 2) * ld] + b[(20 + 2) * ld];
        cov[2] = b[(4 + 8) * ld] + b[(20 + 8) * ld];   cov[3] = b[(4 + 10) * ld] + b[(20 + 10) * ld];
    } else {
#pragma unroll
        for (int f = 0; f < 4; ++f) cov[f] = b[(2 + f) * ld] + b[(6 + f) * ld];
    }
}

// is pair (a,b) (a before b in map order) safe at step t?  `pair_only` selects the eta map
// of the follow-up loop (only the two robots are in states_map).
__device__ __forceinline__ bool pvc_pair_safe(const PvcDev& pv, const double* traj, int64_t ld, const int32_t* lens, int a, int b,
                                              int t, bool pair_only) {
    double mua[2], ca[4], mub[2], cb[4];
    pvc_dist(pv, traj, ld, lens, a, t, mua, ca);
    pvc_dist(pv, traj, ld, lens, b, t, mub, cb);
    if (pv.kind == CCK_PVC_CHISQUARED) {                      // ChiSquaredBoundaryPVC.cpp:139-168
        const double la = max_eig_real_2x2(ca[0], ca[1], ca[2], ca[3]), lb = max_eig_real_2x2(cb[0], cb[1], cb[2], cb[3]);
        const double dx = mua[0] - mub[0], dy = mua[1] - mub[1];
        const double dist = sqrt(dx * dx + dy * dy);
        const double boundary = (la * pv.sc + pv.radii[a]) + (lb * pv.sc + pv.radii[b]);
        return dist > boundary;
    }
    if (pv.kind == CCK_PVC_CDFGRID) {                         // CDFGridPVC.cpp:111-151: r = r_a + r_b (getBoundingBox_, :330-351)
        const int lo = a < b ? a : b, hi = a < b ? b : a;
        return pair_safe_cdfgrid(mua, ca, mub, cb, pv.radii[lo] + pv.radii[hi], pv.grid_n, pv.p_coll_agnts);
    }
    const double dx = mua[0] - mub[0], dy = mua[1] - mub[1];
    if (pv.kind == CCK_PVC_BLACKMORE2) {                      // Blackmore2PVC.cpp:125-167
        const int lo = a < b ? a : b, hi = a < b ? b : a;
        return pair_safe_blackmore2(pv.pair_A + ((int64_t)lo * pv.k + hi) * 8, pv.pair_B + ((int64_t)lo * pv.k + hi) * 4, dx, dy, ca[0] + cb[0],
                                    ca[1] + cb[1], ca[2] + cb[2], ca[3] + cb[3], pv.p_coll_agnts);
    }
    double c = pv.c_pair;
    if (pv.kind == CCK_PVC_ADAPTIVE_BLACKMORE || pv.kind == CCK_PVC_ADAPTIVE_BOUNDINGBOX) {
        // createEtaMap_ (AdaptiveRiskBlackmorePVC.cpp:179-206): sum over the active robots in map order
        double sum = 0, d_b = 0;
        for (int jj = 0; jj < pv.k; ++jj) {
            const int j = pv.map_order[jj];
            if (j == a) continue;
            if (pair_only && j != b) continue;
            double muj[2], cj[4];
            pvc_dist(pv, traj, ld, lens, j, t, muj, cj);
            const double ex = mua[0] - muj[0], ey = mua[1] - muj[1];
            const double di = sqrt(ex * ex + ey * ey);
            if (j == b) d_b = di;
            sum += (1 / di);
        }
        const double alpha = 1 / sum;
        c = erfinv(1 - (2 * ((alpha / d_b) * pv.p_coll_agnts)));
    }
    const int lo = a < b ? a : b, hi = a < b ? b : a;
    const double* A = pv.pair_A + ((int64_t)lo * pv.k + hi) * 8;
    const double* B = pv.pair_B + ((int64_t)lo * pv.k + hi) * 4;
    return pair_safe_halfplanes(A, B, dx, dy, ca[0] + cb[0], ca[1] + cb[1], ca[2] + cb[2], ca[3] + cb[3], c);
}

// map-order pair rank -> (ia, ib) positions
__device__ __forceinline__ void rank_to_pair(int rank, int k, int& ia, int& ib) {
    int row = 0, rem = rank;
    while (rem >= k - 1 - row) { rem -= (k - 1 - row); ++row; }
    ia = row; ib = row + 1 + rem;
}

// Phase 1: one warp per time step; lanes sweep the pairs in map order, a ballot finds the
// first unsafe pair of that step (checkForConflicts_, MinkowskiSumBlackmorePVC.cpp:125-167).
// first_rank[t] = rank of that pair or INT_MAX.  An atomicMin on (t<<32 | rank) yields the
// earliest conflict of the plan.
// [t_begin, t_end): the steps this launch scans (the whole plan, or one rank's share of it: cck_validate_plan_range).
__global__ void __launch_bounds__(128) k_validate_steps(const __grid_constant__ PvcDev pv, const double* __restrict__ traj,
                                                        int64_t ld, const int32_t* __restrict__ lens, int t_begin, int t_end,
                                                        unsigned long long* __restrict__ best_key) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int t = t_begin + warp; t < t_end; t += nwarps) {
        // steps later than an already-found conflict cannot win
        unsigned long long cur = 0;
        if (lane == 0) cur = *reinterpret_cast<volatile unsigned long long*>(best_key);
        cur = __shfl_sync(0xffffffffu, cur, 0);
        if ((cur >> 32) < (unsigned long long)t) continue;
        int found = -1;
        for (int base = 0; base < pv.n_pairs && found < 0; base += 32) {
            const int rank = base + lane;
            bool unsafe = false;
            if (rank < pv.n_pairs) {
                int ia, ib;
                rank_to_pair(rank, pv.k, ia, ib);
                unsafe = !pvc_pair_safe(pv, traj, ld, lens, pv.map_order[ia], pv.map_order[ib], t, false);
            }
            const unsigned m = __ballot_sync(0xffffffffu, unsafe);
            if (m) found = base + (__ffs(m) - 1);
        }
        if (found >= 0 && lane == 0)
            atomicMin(best_key, ((unsigned long long)(unsigned)t << 32) | (unsigned long long)(unsigned)found);
    }
}

// Phase 2 (one CTA): follow the winning pair while it stays in conflict
// (validatePlan's inner while loop, :52-57, quirk 7) and emit the run.
__global__ void __launch_bounds__(256) k_validate_finalize(const __grid_constant__ PvcDev pv, const double* __restrict__ traj,
                                                           int64_t ld, const int32_t* __restrict__ lens, int max_states,
                                                           const unsigned long long* __restrict__ best_key,
                                                           cck_conflict_run* __restrict__ out) {
    __shared__ int first_safe;
    const unsigned long long key = *best_key;
    if (key == ~0ull) {
        if (threadIdx.x == 0) { out->a = -1; out->b = -1; out->first_step = -1; out->run_len = 0; }
        return;
    }
    const int t0 = (int)(key >> 32), rank = (int)(key & 0xffffffffu);
    int ia, ib;
    rank_to_pair(rank, pv.k, ia, ib);
    const int a = pv.map_order[ia], b = pv.map_order[ib];
    if (threadIdx.x == 0) first_safe = max_states;
    __syncthreads();
    for (int base = t0 + 1; base < max_states; base += blockDim.x) {
        const int t = base + threadIdx.x;
        bool safe = false;
        if (t < max_states) safe = pvc_pair_safe(pv, traj, ld, lens, a, b, t, true);
        const unsigned m = __ballot_sync(0xffffffffu, safe);
        if (m && (threadIdx.x & 31) == 0) atomicMin(&first_safe, base + (int)(threadIdx.x & ~31u) + (__ffs(m) - 1));
        __syncthreads();
        const int fs = first_safe;
        __syncthreads();
        if (fs < max_states) break;
    }
    __syncthreads();
    if (threadIdx.x == 0) { out->a = a; out->b = b; out->first_step = t0; out->run_len = first_safe - t0; }
}

// ---------------------------------------------------------------------------
// PlanValidityChecker::satisfiesConstraints, batched over candidate paths
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_constraints(int64_t n_paths, int dim, const double* __restrict__ path, int64_t ld,
                                                     const int32_t* __restrict__ path_len, const int32_t* __restrict__ step0,
                                                     const unsigned char* __restrict__ cons, uint8_t* __restrict__ ok) {
    const ConsHeader* h = reinterpret_cast<const ConsHeader*>(cons);
    const int W = dim + 2 * dim * dim;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_paths; i += (int64_t)gridDim.x * blockDim.x) {
        bool good = true;
        const int len = path_len[i], s0 = step0[i];
        for (int t = 0; t < len && good; ++t) {
            const int s = s0 + t;
            if (s > h->max_step) break;
            const int32_t* first = reinterpret_cast<const int32_t*>(cons + h->off_first);
            if (s < 0 || first[s] == first[s + 1]) continue;
            const double* b = path + (int64_t)t * W * ld + i;
            double C0, C1, C2, C3;
            if (dim == 4) {
                C0 = b[4 * ld] + b[20 * ld]; C1 = b[6 * ld] + b[22 * ld]; C2 = b[12 * ld] + b[28 * ld]; C3 = b[14 * ld] + b[30 * ld];
            } else {
                C0 = b[2 * ld] + b[6 * ld]; C1 = b[3 * ld] + b[7 * ld]; C2 = b[4 * ld] + b[8 * ld]; C3 = b[5 * ld] + b[9 * ld];
            }
            good = constraints_ok_at(cons, s, b[0], b[ld], C0, C1, C2, C3);
        }
        ok[i] = good ? 1 : 0;
    }
}

// ---------------------------------------------------------------------------
// CentralizedChiSquaredBoundarySVC's agent-agent test (CentralizedChiSquaredBoundarySVC.cpp:70-99), batched: the 10-gon of
// radius lambda_max(cov_a) * sc + rad_a around mu_a against the one of agent b; out = 1 iff bg::disjoint (touching is NOT
// disjoint).  Separating axes of two convex polygons: the edge normals of both.
// ---------------------------------------------------------------------------
// one pair; dec_cos / dec_sin: the ten directions of point_circle(10)
__host__ __device__ __forceinline__ bool chisq_pair_disjoint(double mxa, double mya, double c0a, double c1a, double c2a, double c3a, double rada, double mxb,
                                                             double myb, double c0b, double c1b, double c2b, double c3b, double radb, double sc,
                                                             const double* dec_cos, const double* dec_sin) {
    const double ra = (max_eig_real_2x2(c0a, c1a, c2a, c3a) * sc + rada);
    const double rb = (max_eig_real_2x2(c0b, c1b, c2b, c3b) * sc + radb);
    double ax[11], ay[11], bx[11], by[11];
    for (int k = 0; k < 10; ++k) {
        ax[k] = mxa + ra * dec_cos[k]; ay[k] = mya + ra * dec_sin[k];
        bx[k] = mxb + rb * dec_cos[k]; by[k] = myb + rb * dec_sin[k];
    }
    ax[10] = ax[0]; ay[10] = ay[0]; bx[10] = bx[0]; by[10] = by[0];
    bool sep = false;
    for (int pass = 0; pass < 2 && !sep; ++pass) {
        const double* ex = pass ? bx : ax; const double* ey = pass ? by : ay;
        for (int e = 0; e < 10 && !sep; ++e) {
            const double nx = ey[e + 1] - ey[e], ny = ex[e] - ex[e + 1];
            if (nx == 0 && ny == 0) continue;
            // running extrema start at +-inf and move only on a TRUE comparison (the oracle's std::min / std::max): NaN projections
            // are never taken, so a decagon with NaN vertices is "separated" from everything
            double amin = HUGE_VAL, amax = -HUGE_VAL, bmin = HUGE_VAL, bmax = -HUGE_VAL;
            for (int j = 0; j < 10; ++j) {
                const double da = nx * ax[j] + ny * ay[j], db = nx * bx[j] + ny * by[j];
                amin = da < amin ? da : amin; amax = amax < da ? da : amax; bmin = db < bmin ? db : bmin; bmax = bmax < db ? db : bmax;
            }
            sep = (amax < bmin) || (bmax < amin);
        }
    }
    return sep;
}
__global__ void __launch_bounds__(128) k_chisq_pairs(int64_t n, const double* __restrict__ mu_a, const double* __restrict__ cov_a,
                                                     const double* __restrict__ rad_a, const double* __restrict__ mu_b,
                                                     const double* __restrict__ cov_b, const double* __restrict__ rad_b, double sc,
                                                     const __grid_constant__ SvcDev sp, uint8_t* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = chisq_pair_disjoint(mu_a[i], mu_a[n + i], cov_a[i], cov_a[n + i], cov_a[2 * n + i], cov_a[3 * n + i], rad_a[i], mu_b[i], mu_b[n + i],
                                     cov_b[i], cov_b[n + i], cov_b[2 * n + i], cov_b[3 * n + i], rad_b[i], sc, sp.dec_cos, sp.dec_sin) ? 1 : 0;
}

// ---------------------------------------------------------------------------
// per-rank result summary (the record gathered over NCCL)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_summary(int64_t n, const int32_t* __restrict__ steps, const int32_t* __restrict__ valid,
                                                 unsigned long long* __restrict__ out) {
    unsigned long long full = 0, exec = 0, vsum = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int s = steps[i], v = valid[i];
        full += (v == s) ? 1u : 0u;
        exec += (unsigned long long)(s <= 0 ? 0 : (v + 1 < s ? v + 1 : s));
        vsum += (unsigned long long)v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        full += __shfl_xor_sync(0xffffffffu, full, o);
        exec += __shfl_xor_sync(0xffffffffu, exec, o);
        vsum += __shfl_xor_sync(0xffffffffu, vsum, o);
    }
    if ((threadIdx.x & 31) == 0) { atomicAdd(out, full); atomicAdd(out + 1, exec); atomicAdd(out + 2, vsum); }
}

// winner record: lexicographic min of (cost, index) over admissible candidates, one CTA
__global__ void __launch_bounds__(1024) k_winner(int64_t n, const double* __restrict__ cost, const uint8_t* __restrict__ flags,
                                                 int mask, int64_t offset, double* __restrict__ out) {
    __shared__ double sc[32];
    __shared__ long long si[32];
    double bc = HUGE_VAL;
    long long bi = -1;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        if ((flags[i] & mask) != mask) continue;
        const double c = cost[i];
        if (c < bc || (c == bc && (bi < 0 || i < bi))) { bc = c; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double oc = __shfl_xor_sync(0xffffffffu, bc, o);
        const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (oi >= 0 && (bi < 0 || oc < bc || (oc == bc && oi < bi))) { bc = oc; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { sc[threadIdx.x >> 5] = bc; si[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x < 32) {
        bc = threadIdx.x < (blockDim.x >> 5) ? sc[threadIdx.x] : HUGE_VAL;
        bi = threadIdx.x < (blockDim.x >> 5) ? si[threadIdx.x] : -1;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double oc = __shfl_xor_sync(0xffffffffu, bc, o);
            const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (oi >= 0 && (bi < 0 || oc < bc || (oc == bc && oi < bi))) { bc = oc; bi = oi; }
        }
        if (threadIdx.x == 0) { out[0] = bi >= 0 ? bc : HUGE_VAL; out[1] = bi >= 0 ? (double)(offset + bi) : -1.0; }
    }
}

// ---------------------------------------------------------------------------
// FP64 pipe peak: 8 independent DFMA chains per thread (measurement helper)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_fp64_peak(double* out, int iters, double seed) {
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 0.999999, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = __fma_rn(a0, m, c); a1 = __fma_rn(a1, m, c); a2 = __fma_rn(a2, m, c); a3 = __fma_rn(a3, m, c);
        a4 = __fma_rn(a4, m, c); a5 = __fma_rn(a5, m, c); a6 = __fma_rn(a6, m, c); a7 = __fma_rn(a7, m, c);
    }
    const double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (s == 123.456) out[0] = s;   // keep the chains alive
}

// Non-fused FP64 instruction rate: MODE 0 alternating DMUL / DADD, 1 DADD only, 2 DMUL only (8 independent chains per
// thread, __dmul_rn / __dadd_rn are never contracted).  The decision arithmetic of this library may not form FMAs, so
// this - not the DFMA flop peak - is the attainable rate of its FP64 work.
template <int MODE>
__global__ void __launch_bounds__(256) k_fp64_rate(double* out, int iters, double seed) {
    double a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = seed + threadIdx.x + j;
    const double m = 0.999999, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (MODE == 0) a[j] = (i & 1) ? __dmul_rn(a[j], m) : __dadd_rn(a[j], c);
            else if (MODE == 1) a[j] = __dadd_rn(a[j], c);
            else a[j] = __dmul_rn(a[j], m);
        }
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += a[j];
    if (s == 123.456) out[0] = s;
}

// Does an FP64 instruction keep the ISSUE port busy for more than one cycle?  NF fp32 FMULs (own chains) per non-fused FP64
// instruction: if the FP64 rate is unchanged, fp32/int work issues in the shadow of the FP64 pipe; if it drops, it does not.
template <int NF>
__global__ void __launch_bounds__(256) k_fp64_mix(double* out, int iters, double seed) {
    double a[8];
    float f[8 * (NF > 0 ? NF : 1)];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = seed + threadIdx.x + j;
#pragma unroll
    for (int j = 0; j < 8 * NF; ++j) f[j] = (float)seed + threadIdx.x + j;
    const double m = 0.999999, c = 1e-9;
    const float fm = 0.99999f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            a[j] = (i & 1) ? __dmul_rn(a[j], m) : __dadd_rn(a[j], c);
#pragma unroll
            for (int q = 0; q < NF; ++q) f[j * NF + q] = __fmul_rn(f[j * NF + q], fm);
        }
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += a[j];
#pragma unroll
    for (int j = 0; j < 8 * NF; ++j) s += (double)f[j];
    if (s == 123.456) out[0] = s;
}

}  // namespace cck
