// cck_oracle_capi.cpp — C entry points over the CPU ORACLE for ctypes
// (test infrastructure; see the header of cck_oracle.hpp).  Beliefs are AoS
// records of doubles: [values(D) | sigma(DxD row-major) | lambda(DxD row-major)],
// D=2 -> 10 doubles, D=4 -> 36 doubles, mirroring the reference's per-state
// objects (include/Spaces/RealVectorBeliefSpace.h:21-96).
#include <thread>

#include "cck_oracle.hpp"

using namespace cck_oracle;

namespace {
struct World {
    Instance inst;
};
struct SvcH {
    Svc svc;
};
struct PvcH {
    Pvc pvc;
};

template <int D, class Model>
void rollout_range(const Model& m, const Svc& svc, double dt, const double* bel, const double* ctrl, const int* steps,
                   double* out, int* valid, double* traj, int traj_stride, uint64_t* cnt3, long lo, long hi) {
    constexpr int W = D + 2 * D * D;
    Counters c;
    std::vector<Belief<D>> tbuf;
    for (long i = lo; i < hi; ++i) {
        Belief<D> s, r;
        std::memcpy(&s, bel + i * W, sizeof(s));
        Belief<D>* tp = nullptr;
        if (traj) { tbuf.assign(steps[i] > 0 ? steps[i] : 1, Belief<D>{}); tp = tbuf.data(); }
        valid[i] = propagate_while_valid<D>(m, svc, dt, s, ctrl + i * D, steps[i], r, &c, tp);
        std::memcpy(out + i * W, &r, sizeof(r));
        if (traj)
            for (int t = 0; t < valid[i]; ++t) std::memcpy(traj + ((size_t)i * traj_stride + t) * W, &tbuf[t], sizeof(Belief<D>));
    }
    cnt3[0] = c.n_steps; cnt3[1] = c.n_halfplane; cnt3[2] = c.n_obstacle;
}

template <int D, class Model>
void rollout_mt(const Model& m, const Svc& svc, double dt, long n, const double* bel, const double* ctrl, const int* steps,
                double* out, int* valid, double* traj, int traj_stride, uint64_t* counters, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    std::vector<std::thread> th;
    std::vector<uint64_t> cnt((size_t)nthreads * 3, 0);
    const long chunk = (n + nthreads - 1) / nthreads;
    for (int t = 0; t < nthreads; ++t) {
        const long lo = t * chunk, hi = std::min(n, lo + chunk);
        if (lo >= hi) break;
        if (nthreads == 1) rollout_range<D>(m, svc, dt, bel, ctrl, steps, out, valid, traj, traj_stride, &cnt[0], lo, hi);
        else th.emplace_back([=, &m, &svc, &cnt] { rollout_range<D>(m, svc, dt, bel, ctrl, steps, out, valid, traj, traj_stride, &cnt[(size_t)t * 3], lo, hi); });
    }
    for (auto& x : th) x.join();
    if (counters) {
        counters[0] = counters[1] = counters[2] = 0;
        for (int t = 0; t < nthreads; ++t) for (int j = 0; j < 3; ++j) counters[j] += cnt[(size_t)t * 3 + j];
    }
}
}  // namespace

extern "C" {

double orc_erf_inv(double z) { return erf_inv(z); }
double orc_chi2_quantile_2dof(double p) { return chi2_quantile_2dof(p); }

// out: dt, Q[4], R[4], Acl[4]  (13 doubles)
void orc_linear_params(double dt, double* out) {
    LinearParams p = make_linear_params(dt);
    out[0] = p.dt;
    for (int i = 0; i < 4; ++i) { out[1 + i] = p.Q[i]; out[5 + i] = p.R[i]; out[9 + i] = p.Acl[i]; }
}
// out: F[16], Acl[16], Q[16], R[16], gains[8], acc_lo, acc_hi, turn_lo, turn_hi  (76 doubles)
void orc_unicycle_params(double dt, double* out) {
    UnicycleParams p = make_unicycle_params(dt);
    std::memcpy(out, p.F, sizeof(p.F)); std::memcpy(out + 16, p.Acl, sizeof(p.Acl));
    std::memcpy(out + 32, p.Q, sizeof(p.Q)); std::memcpy(out + 48, p.R, sizeof(p.R));
    std::memcpy(out + 64, p.gains, sizeof(p.gains));
    out[72] = p.acc_lo; out[73] = p.acc_hi; out[74] = p.turn_lo; out[75] = p.turn_hi;
}

// ---- world ----
void* orc_world_from_text(const char* map_text, const char* scen_text, int num_agents, double p_safe) {
    World* w = new World();
    if (!load_map_text(map_text, w->inst) || !load_scen_text(scen_text, num_agents, w->inst)) { delete w; return nullptr; }
    split_risk(w->inst, p_safe);
    return w;
}
// synthetic world: unit-square obstacles at the given centres, k identical 0.25x0.25 robots
void* orc_world_synthetic(int n_obs, const double* centres, double x_max, double y_max, int k_agents, double p_safe) {
    World* w = new World();
    w->inst.x_max = x_max; w->inst.y_max = y_max;
    for (int i = 0; i < n_obs; ++i)
        w->inst.obstacles.push_back(Obstacle{centres[2 * i], centres[2 * i + 1], rect_ring(centres[2 * i], centres[2 * i + 1], 1, 1)});
    for (int i = 0; i < k_agents; ++i) {
        RobotSpec r;
        r.name = "Robot " + std::to_string(i); r.model = "synthetic";
        r.sx = r.sy = r.gx = r.gy = 0;
        r.shape = rectangular_robot(0, 0, 0.25, 0.25);
        w->inst.robots.push_back(r);
    }
    split_risk(w->inst, p_safe);
    return w;
}
void orc_world_free(void* w) { delete (World*)w; }
int orc_world_n_obstacles(void* w) { return (int)((World*)w)->inst.obstacles.size(); }
int orc_world_n_robots(void* w) { return (int)((World*)w)->inst.robots.size(); }
int orc_world_map_overrun(void* w) { return ((World*)w)->inst.map_overrun ? 1 : 0; }
// out: x_max, y_max, p_safe_obs, p_safe_agents
void orc_world_info(void* w, double* out) {
    const Instance& i = ((World*)w)->inst;
    out[0] = i.x_max; out[1] = i.y_max; out[2] = i.p_safe_obs; out[3] = i.p_safe_agents;
}
void orc_world_obstacle_centres(void* w, double* out) {
    const Instance& i = ((World*)w)->inst;
    for (size_t o = 0; o < i.obstacles.size(); ++o) { out[2 * o] = i.obstacles[o].x; out[2 * o + 1] = i.obstacles[o].y; }
}
// out per robot: sx, sy, gx, gy, bounding_rad
void orc_world_robots(void* w, double* out) {
    const Instance& i = ((World*)w)->inst;
    for (size_t r = 0; r < i.robots.size(); ++r) {
        out[5 * r] = i.robots[r].sx; out[5 * r + 1] = i.robots[r].sy; out[5 * r + 2] = i.robots[r].gx; out[5 * r + 3] = i.robots[r].gy;
        out[5 * r + 4] = i.robots[r].shape.bounding_rad;
    }
}
int orc_world_robot_model(void* w, int r, char* buf, int cap) {
    const std::string& m = ((World*)w)->inst.robots[r].model;
    std::snprintf(buf, cap, "%s", m.c_str());
    return (int)m.size();
}

// ---- SVC ----
// accep_prob < 0 => the reference's wiring (OmplSetUp.cpp:213-222): p_safe_obs for
// Blackmore/Adaptive, the hard-coded 0.9 for ChiSquared.
void* orc_svc_create(void* w, int kind, int dim, int robot, const double* low, const double* high, double accep_prob) {
    const Instance& inst = ((World*)w)->inst;
    if (accep_prob < 0) accep_prob = (kind == SVC_CHISQUARED) ? 0.9 : inst.p_safe_obs;
    SvcH* h = new SvcH();
    h->svc = make_svc(kind, dim, low, high, inst.obstacles, inst.robots[robot].shape, accep_prob);
    return h;
}
void orc_svc_free(void* s) { delete (SvcH*)s; }
// out: erf_inv_result, p_collision, sc, max_rad
void orc_svc_info(void* s, double* out) {
    const Svc& v = ((SvcH*)s)->svc;
    out[0] = v.erf_inv_result; out[1] = v.p_collision; out[2] = v.sc; out[3] = v.max_rad;
}
// test hook: the robot bounding radius the chi-squared checker adds to lambda_max * sc (band tests shift it by +-1e-6)
void orc_svc_set_max_rad(void* s, double r) { ((SvcH*)s)->svc.max_rad = r; }
// A[M][4][2], B[M][4], centres[M][2], rects[M][4] = (xlo, ylo, xhi, yhi) of the inflated ring
void orc_svc_tables(void* s, double* A, double* B, double* centres, double* rects) {
    const Svc& v = ((SvcH*)s)->svc;
    for (size_t o = 0; o < v.tab.hp.size(); ++o) {
        for (int i = 0; i < 4; ++i) { A[o * 8 + i * 2] = v.tab.hp[o].A[i][0]; A[o * 8 + i * 2 + 1] = v.tab.hp[o].A[i][1]; B[o * 4 + i] = v.tab.hp[o].B[i]; }
        centres[o * 2] = v.tab.centre[o].x; centres[o * 2 + 1] = v.tab.centre[o].y;
        const Ring& m = v.tab.minkowski[o];
        rects[o * 4] = m[0].x; rects[o * 4 + 1] = m[0].y; rects[o * 4 + 2] = m[2].x; rects[o * 4 + 3] = m[2].y;
    }
}

// Test hook: replace the half-plane tables (and the erf_inv constant) of a Blackmore checker with
// arbitrary ones, so parity can be checked on tables the map parser never produces (rotated or
// duplicated planes, non-finite offsets).  Only HyperplaneCCValidityChecker_ reads them.
void orc_svc_override_tables(void* s, int M, const double* A, const double* B, double erf_inv_result) {
    Svc& v = ((SvcH*)s)->svc;
    v.tab.hp.assign(M, HalfPlanes{});
    v.tab.centre.assign(M, Pt{0, 0});
    v.tab.minkowski.assign(M, Ring(5, Pt{0, 0}));
    for (int o = 0; o < M; ++o)
        for (int i = 0; i < 4; ++i) { v.tab.hp[o].A[i][0] = A[o * 8 + i * 2]; v.tab.hp[o].A[i][1] = A[o * 8 + i * 2 + 1]; v.tab.hp[o].B[i] = B[o * 4 + i]; }
    v.erf_inv_result = erf_inv_result;
}

// ---- propagation / validity (batched) ----
void orc_propagate(int dim, double dt, double duration, long n, const double* bel, const double* ctrl, double* out) {
    if (dim == 2) {
        Model2 m{make_linear_params(dt)};
        for (long i = 0; i < n; ++i) { Belief2 s, r; std::memcpy(&s, bel + i * 10, sizeof(s)); m.step(s, ctrl + i * 2, duration, r); std::memcpy(out + i * 10, &r, sizeof(r)); }
    } else {
        Model4 m{make_unicycle_params(dt)};
        for (long i = 0; i < n; ++i) { Belief4 s, r; std::memcpy(&s, bel + i * 36, sizeof(s)); m.step(s, ctrl + i * 4, duration, r); std::memcpy(out + i * 36, &r, sizeof(r)); }
    }
}

void orc_is_valid(void* s, int dim, long n, const double* bel, unsigned char* out, uint64_t* counters) {
    const Svc& v = ((SvcH*)s)->svc;
    Counters c;
    const int W = dim + 2 * dim * dim;
    for (long i = 0; i < n; ++i) {
        const double* b = bel + i * W;
        double c2[4];
        const double* S = b + dim; const double* L = b + dim + dim * dim;
        c2[0] = S[0] + L[0]; c2[1] = S[1] + L[1]; c2[2] = S[dim] + L[dim]; c2[3] = S[dim + 1] + L[dim + 1];
        out[i] = svc_is_valid(v, b, c2, &c) ? 1 : 0;
    }
    if (counters) { counters[0] = c.n_steps; counters[1] = c.n_halfplane; counters[2] = c.n_obstacle; }
}

// propagateWhileValid over a batch.  traj (optional): [n][traj_stride][W] receives
// the valid intermediate states.  counters: n_steps, n_halfplane, n_obstacle.
void orc_rollout(void* s, int dim, double dt, long n, const double* bel, const double* ctrl, const int* steps,
                 double* out, int* valid, double* traj, int traj_stride, uint64_t* counters, int nthreads) {
    const Svc& v = ((SvcH*)s)->svc;
    if (dim == 2) { Model2 m{make_linear_params(dt)}; rollout_mt<2>(m, v, dt, n, bel, ctrl, steps, out, valid, traj, traj_stride, counters, nthreads); }
    else { Model4 m{make_unicycle_params(dt)}; rollout_mt<4>(m, v, dt, n, bel, ctrl, steps, out, valid, traj, traj_stride, counters, nthreads); }
}

// propagate n steps without validity checks (PathControl::interpolate building block)
void orc_propagate_n(int dim, double dt, const double* bel, const double* ctrl, int steps, double* out) {
    if (dim == 2) { Model2 m{make_linear_params(dt)}; Belief2 s; std::memcpy(&s, bel, sizeof(s)); propagate_n<2>(m, dt, s, ctrl, steps, (Belief2*)out); }
    else { Model4 m{make_unicycle_params(dt)}; Belief4 s; std::memcpy(&s, bel, sizeof(s)); propagate_n<4>(m, dt, s, ctrl, steps, (Belief4*)out); }
}

// ---- PVC ----
void* orc_pvc_create(void* w, int kind, double p_safe_arg) {
    const Instance& inst = ((World*)w)->inst;
    std::vector<RobotShape> shapes;
    for (const auto& r : inst.robots) shapes.push_back(r.shape);
    // reference wiring (demos/main.cpp:105-138): Blackmore-type PVCs get p_safe_agents;
    // ChiSquared gets a hard-coded 0.9 (quirk 3)
    if (p_safe_arg < 0) p_safe_arg = (kind == PVC_CHISQUARED) ? 0.9 : inst.p_safe_agents;
    PvcH* h = new PvcH();
    h->pvc = make_pvc(kind, shapes, p_safe_arg);
    return h;
}
void orc_pvc_set_grid_n(void* p, int n) { ((PvcH*)p)->pvc.grid_n = n; }     // CDFGridPVC's disk_steps_
void orc_pvc_free(void* p) { delete (PvcH*)p; }
// CDFGridPVC::isSafe_ in isolation: collision mass of n pairs (mu[2], cov[4] each), radius sum r, grid n
void orc_cdfgrid_prob(long n, const double* mu_a, const double* cov_a, const double* mu_b, const double* cov_b, double r_sum, int grid_n, double* prob) {
    for (long i = 0; i < n; ++i) pair_is_safe_cdfgrid(mu_a + 2 * i, cov_a + 4 * i, mu_b + 2 * i, cov_b + 4 * i, r_sum, grid_n, 1.0, prob + i);
}
// out: p_coll_agnts, p_coll_dist, c_pair, sc
void orc_pvc_info(void* p, double* out) {
    const Pvc& v = ((PvcH*)p)->pvc;
    out[0] = v.p_coll_agnts; out[1] = v.p_coll_dist; out[2] = v.c_pair; out[3] = v.sc;
}
void orc_pvc_map_order(void* p, int* out) {
    const Pvc& v = ((PvcH*)p)->pvc;
    for (int i = 0; i < v.k; ++i) out[i] = v.map_order[i];
}
// pair half-planes for (a<b): A[8], B[4]
void orc_pvc_pair_halfplanes(void* p, int a, int b, double* A, double* B) {
    const Pvc& v = ((PvcH*)p)->pvc;
    const HalfPlanes& hp = v.pair_hp[(size_t)std::min(a, b) * v.k + std::max(a, b)];
    for (int i = 0; i < 4; ++i) { A[i * 2] = hp.A[i][0]; A[i * 2 + 1] = hp.A[i][1]; B[i] = hp.B[i]; }
}

}  // extern "C"
template <int D>
static std::vector<std::vector<Belief<D>>> unpack_trajs(int k, const int* lens, const double* states) {
    constexpr int W = D + 2 * D * D;
    std::vector<std::vector<Belief<D>>> trajs(k);
    size_t off = 0;
    for (int r = 0; r < k; ++r) {
        trajs[r].resize(lens[r]);
        for (int t = 0; t < lens[r]; ++t) std::memcpy(&trajs[r][t], states + (off + t) * W, sizeof(Belief<D>));
        off += lens[r];
    }
    return trajs;
}
extern "C" {

// validatePlan: trajectories concatenated robot after robot.  Returns #conflicts;
// out_conf[i] = (a, b, step), at most cap rows written.
int orc_validate_plan(void* p, int dim, int k, const int* lens, const double* states, int* out_conf, int cap) {
    const Pvc& v = ((PvcH*)p)->pvc;
    std::vector<Conflict> c;
    if (dim == 2) c = validate_plan<2>(v, unpack_trajs<2>(k, lens, states));
    else c = validate_plan<4>(v, unpack_trajs<4>(k, lens, states));
    for (size_t i = 0; i < c.size() && (int)i < cap; ++i) { out_conf[3 * i] = c[i].a; out_conf[3 * i + 1] = c[i].b; out_conf[3 * i + 2] = c[i].step; }
    return (int)c.size();
}

// satisfiesConstraints for one path.  Constraints are flattened: for constraint j,
// constraining[j], n_times[j], then steps / states concatenated.
int orc_satisfies_constraints(void* p, int dim, int path_len, const double* path, double step_size, int constrained,
                              int n_constraints, const int* constraining, const int* n_times, const int* steps, const double* states) {
    const Pvc& v = ((PvcH*)p)->pvc;
    auto run = [&](auto tag) {
        constexpr int D = decltype(tag)::value;
        constexpr int W = D + 2 * D * D;
        std::vector<Belief<D>> pth(path_len);
        for (int t = 0; t < path_len; ++t) std::memcpy(&pth[t], path + (size_t)t * W, sizeof(Belief<D>));
        std::vector<BeliefConstraint<D>> cons(n_constraints);
        size_t off = 0;
        for (int j = 0; j < n_constraints; ++j) {
            cons[j].constrained = constrained; cons[j].constraining = constraining[j];
            for (int i = 0; i < n_times[j]; ++i) {
                cons[j].steps.push_back(steps[off + i]);
                cons[j].times.push_back(steps[off + i] * step_size);
                Belief<D> b; std::memcpy(&b, states + (off + i) * W, sizeof(b));
                cons[j].states.push_back(b);
            }
            off += n_times[j];
        }
        return satisfies_constraints<D>(v, pth, step_size, cons) ? 1 : 0;
    };
    if (dim == 2) return run(std::integral_constant<int, 2>{});
    return run(std::integral_constant<int, 4>{});
}

// CentralizedChiSquaredBoundarySVC agent-agent test for n pairs (AoS: mu[n][2], cov[n][4], rad[n])
void orc_chisq_pairs_disjoint(long n, const double* mu_a, const double* cov_a, const double* rad_a, const double* mu_b, const double* cov_b,
                              const double* rad_b, double sc, unsigned char* out) {
    for (long i = 0; i < n; ++i) out[i] = chisq_pair_disjoint(mu_a + 2 * i, cov_a + 4 * i, rad_a[i], mu_b + 2 * i, cov_b + 4 * i, rad_b[i], sc) ? 1 : 0;
}

// ChanceConstrainedGoal::isSatisfied over a batch: out_ok[n], out_dist[n]
void orc_goal(int dim, long n, const double* bel, double gx, double gy, double threshold, double p_safe, unsigned char* out_ok, double* out_dist) {
    const int W = dim + 2 * dim * dim;
    for (long i = 0; i < n; ++i) {
        const double* b = bel + i * W;
        const double* S = b + dim; const double* L = b + dim + dim * dim;
        double c2[4] = {S[0] + L[0], S[1] + L[1], S[dim] + L[dim], S[dim + 1] + L[dim + 1]};
        out_ok[i] = goal_is_satisfied(b, c2, gx, gy, threshold, p_safe, out_dist + i) ? 1 : 0;
    }
}

// ---- f2: belief-space distance and the planner's nearest-neighbour rules (AoS records) ----
void orc_belief_distance(int dim, long n, const double* a, const double* b, double* out) {
    const int W = dim + 2 * dim * dim;
    for (long i = 0; i < n; ++i) out[i] = belief_distance(a + i * W, b + i * W, dim);
}
void orc_select_nodes(int dim, long n_nodes, const double* nodes, const double* cost, const unsigned char* active, long K, const double* samples,
                      double radius, int* out_idx, double* out_dist) {
    const int W = dim + 2 * dim * dim;
    for (long q = 0; q < K; ++q) out_idx[q] = select_node(nodes, cost, active, n_nodes, samples + q * W, dim, radius, out_dist + q);
}
void orc_nearest_witness(int dim, long n_wit, const double* wit, long K, const double* queries, int* out_idx, double* out_dist) {
    const int W = dim + 2 * dim * dim;
    for (long q = 0; q < K; ++q) out_idx[q] = nearest_witness(wit, n_wit, queries + q * W, dim, out_dist + q);
}

}  // extern "C"
