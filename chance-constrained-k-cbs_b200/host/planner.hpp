// planner.hpp — the low-level kinodynamic belief-SST of the reference with its expansion loop
// re-designed for the GPU: instead of one (node, control, duration) candidate per iteration
// (ConstraintRespectingBSST.cpp:223-537) the loop samples K candidates, ships them to ONE
// cck_expand_batch launch (propagateWhileValid + constraint check on the new segment + path
// length cost + goal test) and then commits the survivors to the tree in sample order with
// the reference's witness / pruning rules.  A thin K-CBS driver (KCBS.cpp:192-495) sits on top.
//
// The random stream necessarily differs from the reference's (which is never seeded), so
// planner-level results are comparable only statistically; parity is asserted at the
// propagate/check level (tests/).  SST parameters: goal bias 0.05, selection radius 0.2,
// pruning radius 0.1 (ConstraintRespectingBSST.h:232-238); control duration uniform in
// [1,10] steps (OmplSetUp.cpp:225); controls uniform in the control bounds.
#pragma once
#include <chrono>
#include <cmath>
#include <limits>
#include <queue>
#include <random>
#include <unordered_map>

#include "adapters.hpp"

namespace cck_b200 {

// ---------------------------------------------------------------------------
// RealVectorBeliefSpace::distance (src/Spaces/RealVectorBeliefSpace.cpp:16-62): squared mean
// distance + trace(C1 + C2 - 2 (C2^1/2 C1 C2^1/2)^1/2), |.| of the total.  The covariance of a
// tree node depends only on its depth (the recursion is control independent), so the trace term
// is memoised per (depth, depth) / (sample diag, depth) instead of three eigen-solves per call.
// ---------------------------------------------------------------------------
struct SymEig {
    // Jacobi eigen-decomposition of the symmetric part (lower triangle, as SelfAdjointEigenSolver reads it)
    static void sqrt_sym(const double* Ain, int n, double* out) {
        double A[16], V[16];
        for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) { A[i * n + j] = Ain[std::max(i, j) * n + std::min(i, j)]; V[i * n + j] = i == j; }
        for (int sweep = 0; sweep < 60; ++sweep) {
            double off = 0;
            for (int i = 0; i < n; ++i) for (int j = 0; j < i; ++j) off += A[i * n + j] * A[i * n + j];
            if (off < 1e-300) break;
            for (int p = 0; p < n; ++p)
                for (int q = p + 1; q < n; ++q) {
                    if (std::fabs(A[p * n + q]) < 1e-300) continue;
                    const double th = (A[q * n + q] - A[p * n + p]) / (2 * A[p * n + q]);
                    const double t = (th >= 0 ? 1.0 : -1.0) / (std::fabs(th) + std::sqrt(th * th + 1));
                    const double c = 1 / std::sqrt(t * t + 1), s = t * c;
                    for (int k = 0; k < n; ++k) { const double akp = A[k * n + p], akq = A[k * n + q]; A[k * n + p] = c * akp - s * akq; A[k * n + q] = s * akp + c * akq; }
                    for (int k = 0; k < n; ++k) { const double apk = A[p * n + k], aqk = A[q * n + k]; A[p * n + k] = c * apk - s * aqk; A[q * n + k] = s * apk + c * aqk; }
                    for (int k = 0; k < n; ++k) { const double vkp = V[k * n + p], vkq = V[k * n + q]; V[k * n + p] = c * vkp - s * vkq; V[k * n + q] = s * vkp + c * vkq; }
                }
        }
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) {
                double acc = 0;
                for (int k = 0; k < n; ++k) acc += V[i * n + k] * std::sqrt(std::max(A[k * n + k], 0.0)) * V[j * n + k];
                out[i * n + j] = acc;
            }
    }
    static double cov_term(const double* C1, const double* C2, int n) {
        double s2[16], t[16], m[16], s3[16];
        sqrt_sym(C2, n, s2);
        for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) { double a = 0; for (int k = 0; k < n; ++k) a += s2[i * n + k] * C1[k * n + j]; t[i * n + j] = a; }
        for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) { double a = 0; for (int k = 0; k < n; ++k) a += t[i * n + k] * s2[k * n + j]; m[i * n + j] = a; }
        sqrt_sym(m, n, s3);
        double tr = 0;
        for (int i = 0; i < n; ++i) tr += C1[i * n + i] + C2[i * n + i] - 2 * s3[i * n + i];
        return tr;
    }
};

struct PlannerParams {
    int K = 1024;                 // candidates per launch
    double goal_bias = 0.05, selection_radius = 0.2, pruning_radius = 0.1;
    double goal_threshold = 2.0, goal_p_safe = 0.95;   // ChanceConstrainedGoal(si, goal, 2.0, 0.95)  (OmplSetUp.cpp:239)
    int min_steps = 1, max_steps = 10;
    uint64_t seed = 1;
};

struct PlannerStats { int64_t launches = 0, candidates = 0, accepted = 0, nodes = 0; double seconds = 0; };

// One robot's low-level planning problem (what set_up_ConstraintBSST_MP_Problems wires per robot,
// OmplSetUp.cpp:183-353): belief space + bounds, control bounds, SVC, propagator, start, goal.
class BatchedBSST {
public:
    struct Motion {
        std::vector<double> state;    // [W] belief record
        double ctrl[4] = {0, 0, 0, 0};
        int steps = 0, parent = -1, depth = 0;   // depth = absolute step index of this state
        double accCost = 0;
        bool inactive = false;
        int numChildren = 0;
    };
    struct Witness { std::vector<double> state; int depth; int rep; };

    BatchedBSST(InstancePtr inst, int robot, const std::string& svc_name, const std::string& pvc_name, PlannerParams prm, int device = 0)
        : inst_(std::move(inst)), robot_(robot), prm_(prm), rng_(prm.seed + 7919 * (uint64_t)robot) {
        const Robot& r = inst_->getRobots()[robot_];
        dim_ = r.getDynamicsModel() == "Uncertain-Unicycle-Model" ? 4 : 2;
        W_ = belief_width(dim_);
        gpu_ = std::make_shared<GpuContext>(device);
        // spaces + bounds (OmplSetUp.cpp:194-207 / 263-289)
        auto space = std::make_shared<RealVectorBeliefSpace>(dim_);
        ob::RealVectorBounds b(dim_);
        b.setLow(0, -1); b.setHigh(0, inst_->getDimensions()[0]); b.setLow(1, -1); b.setHigh(1, inst_->getDimensions()[1]);
        ob::RealVectorBounds cb(dim_);
        if (dim_ == 4) { b.setLow(2, -M_PI); b.setHigh(2, M_PI); b.setLow(3, 0.01); b.setHigh(3, 10.0); cb = b; }
        else { cb.setLow(-1.0); cb.setHigh(1.0); }
        space->setBounds(b);
        sbounds_ = b; cbounds_ = cb;
        auto cspace = std::make_shared<oc::RealVectorControlSpace>(space, dim_);
        cspace->setBounds(cb);
        si_ = std::make_shared<oc::SpaceInformation>(space, cspace);
        si_->setPropagationStepSize(0.2);                                   // :186,224
        si_->setMinMaxControlDuration(prm_.min_steps, prm_.max_steps);      // :225
        // propagator + validity checker (the adapters install the device tables)
        if (dim_ == 2) si_->setStatePropagator(std::make_shared<R2_UncertainLinearStatePropagator>(si_, gpu_));
        else si_->setStatePropagator(std::make_shared<DynUnicycleControlSpace>(si_, gpu_));
        if (svc_name == "Blackmore") si_->setStateValidityChecker(std::make_shared<PCCBlackmoreSVC>(si_, gpu_, inst_, &r, inst_->getPsafeObs()));
        else if (svc_name == "AdaptiveBlackmore") si_->setStateValidityChecker(std::make_shared<AdaptiveRiskBlackmoreSVC>(si_, gpu_, inst_, &r, inst_->getPsafeObs()));
        else if (svc_name == "ChiSquared") si_->setStateValidityChecker(std::make_shared<ChiSquaredBoundarySVC>(si_, gpu_, inst_, &r, 0.9));   // :221
        else throw std::runtime_error("unknown --svc " + svc_name);
        pvc_ = make_pvc(pvc_name, gpu_, inst_, dim_, 0.2);
        // start state: Sigma0 = Lambda0 = 0.01 I, yaw 0, surge 0.1  (:232-236, 315-320)
        start_.assign(W_, 0.0);
        start_[0] = r.getStartLocation().x_; start_[1] = r.getStartLocation().y_;
        if (dim_ == 4) { start_[2] = 0.0; start_[3] = 0.1; }
        for (int i = 0; i < dim_; ++i) { start_[dim_ + i * dim_ + i] = 0.01; start_[dim_ + dim_ * dim_ + i * dim_ + i] = 0.01; }
        goal_.gx = r.getGoalLocation().x_; goal_.gy = r.getGoalLocation().y_;
        goal_.threshold = prm_.goal_threshold; goal_.sc = cck_chi2_quantile_2dof(prm_.goal_p_safe);
    }

    const oc::SpaceInformationPtr& getSpaceInformation() const { return si_; }
    const PlanValidityCheckerPtr& getPlanValidator() const { return pvc_; }
    const GpuContextPtr& gpu() const { return gpu_; }
    int dim() const { return dim_; }
    const PlannerStats& stats() const { return stats_; }

    // ConstraintRespectingPlanner::updateConstraints + clear
    void updateConstraints(const std::vector<ConstraintPtr>& c) { constraints_ = c; clear(); }
    void clear() { motions_.clear(); witnesses_.clear(); cov_cache_.clear(); }

    // solve(seconds): returns the interpolated trajectory (one state per dt) or empty
    Trajectory solve(double seconds, double* cost_out = nullptr) {
        const auto t0 = std::chrono::steady_clock::now();
        pvc_->installConstraints(constraints_);
        if (motions_.empty()) {
            Motion m; m.state = start_;
            motions_.push_back(m);
            find_closest_witness((int)motions_.size() - 1, true);
        }
        const int K = prm_.K;
        std::vector<double> bel((size_t)W_ * K), ctrl((size_t)dim_ * K), out((size_t)W_ * K), pcost(K), cost(K), gdist(K);
        std::vector<int32_t> steps(K), pstep(K), valid(K);
        std::vector<uint8_t> flags(K);
        std::vector<int> parent(K);
        std::uniform_real_distribution<double> U01(0.0, 1.0);
        int solution = -1;
        while (solution < 0) {
            if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > seconds) break;
            // ---- sample K candidates (ConstraintRespectingBSST.cpp:226-247) ----
            for (int j = 0; j < K; ++j) {
                double rs[4], diag[2];
                if (U01(rng_) < prm_.goal_bias) { rs[0] = goal_.gx; rs[1] = goal_.gy; for (int d = 2; d < dim_; ++d) rs[d] = uni(sbounds_.low[d], sbounds_.high[d]); }
                else for (int d = 0; d < dim_; ++d) rs[d] = uni(sbounds_.low[d], sbounds_.high[d]);
                diag[0] = U01(rng_) * max_eig_; diag[1] = U01(rng_) * max_eig_;           // :235-236
                const int n = select_node(rs, diag);
                parent[j] = n;
                const Motion& m = motions_[n];
                for (int f = 0; f < W_; ++f) bel[(size_t)f * K + j] = m.state[f];
                for (int d = 0; d < dim_; ++d) ctrl[(size_t)d * K + j] = uni(cbounds_.low[d], cbounds_.high[d]);
                steps[j] = prm_.min_steps + (int)(rng_() % (uint64_t)(prm_.max_steps - prm_.min_steps + 1));
                pstep[j] = m.depth; pcost[j] = m.accCost;
            }
            // ---- ONE launch: K x (<=10 steps) propagate + check (+ constraints, cost, goal) ----
            gpu_->check(cck_expand_batch(gpu_->get(), K, bel.data(), K, ctrl.data(), K, steps.data(), pstep.data(), pcost.data(), &goal_, out.data(), K,
                                         valid.data(), cost.data(), gdist.data(), flags.data()), "cck_expand_batch");
            stats_.launches++; stats_.candidates += K;
            // ---- commit in sample order with the reference's witness rules (:250-326) ----
            for (int j = 0; j < K && solution < 0; ++j) {
                if (!(flags[j] & 1)) continue;                               // propCd != cd
                Motion m;
                m.state.resize(W_);
                for (int f = 0; f < W_; ++f) m.state[f] = out[(size_t)f * K + j];
                for (int d = 0; d < dim_; ++d) m.ctrl[d] = ctrl[(size_t)d * K + j];
                m.steps = steps[j]; m.parent = parent[j]; m.depth = motions_[parent[j]].depth + steps[j]; m.accCost = cost[j];
                motions_.push_back(m);
                const int idx = (int)motions_.size() - 1;
                const int w = find_closest_witness(idx, false);
                Witness& wit = witnesses_[w];
                const bool is_new = wit.rep == idx;
                if (!(is_new || m.accCost < motions_[wit.rep].accCost)) { motions_.pop_back(); continue; }
                if (!constraints_.empty() && !(flags[j] & 2)) {             // satisfiesConstraints failed (:300)
                    if (is_new) witnesses_.pop_back();
                    motions_.pop_back();
                    continue;
                }
                if (!is_new) { prune(wit.rep); wit.rep = idx; }
                motions_[parent[j]].numChildren++;
                stats_.accepted++;
                const double c00 = m.state[dim_] + m.state[dim_ + dim_ * dim_], c11 = m.state[dim_ + dim_ + 1] + m.state[dim_ + dim_ * dim_ + dim_ + 1];
                if (c00 > max_eig_) max_eig_ = c00; else if (c11 > max_eig_) max_eig_ = c11;     // :313-320
                if (flags[j] & 4) solution = idx;                            // goal->isSatisfied (:326)
            }
        }
        stats_.nodes = (int64_t)motions_.size();
        stats_.seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (solution < 0) return {};
        if (cost_out) *cost_out = motions_[solution].accCost;
        return extract(solution);
    }

private:
    double uni(double lo, double hi) { return lo + (hi - lo) * std::uniform_real_distribution<double>(0.0, 1.0)(rng_); }

    // covariance (Sigma + Lambda) of a tree state
    void cov_of(const std::vector<double>& s, double* C) const { for (int f = 0; f < dim_ * dim_; ++f) C[f] = s[dim_ + f] + s[dim_ + dim_ * dim_ + f]; }

    double cov_term_nodes(int d1, const std::vector<double>& s1, int d2, const std::vector<double>& s2) {
        const uint64_t key = ((uint64_t)(uint32_t)d1 << 32) | (uint32_t)d2;
        auto it = cov_cache_.find(key);
        if (it != cov_cache_.end()) return it->second;
        double C1[16], C2[16];
        cov_of(s1, C1); cov_of(s2, C2);
        const double v = SymEig::cov_term(C1, C2, dim_);
        cov_cache_[key] = v;
        return v;
    }

    // selectNode (ConstraintRespectingBSST.cpp:119-148): best-cost active node within the selection
    // radius of the sample, else the nearest active node.  The sample carries sigma = diag(d0, d1, 0.01..),
    // lambda = 0.01 I (allocState default).
    int select_node(const double* rs, const double* diag) {
        double Cs[16] = {0};
        for (int i = 0; i < dim_; ++i) Cs[i * dim_ + i] = 0.02;
        Cs[0] = diag[0] + 0.01; Cs[dim_ + 1] = diag[1] + 0.01;
        std::unordered_map<int, double> term;   // per-depth covariance term for this sample
        int best = -1, nearest = -1;
        double best_cost = std::numeric_limits<double>::infinity(), nd = std::numeric_limits<double>::infinity();
        for (int i = 0; i < (int)motions_.size(); ++i) {
            const Motion& m = motions_[i];
            if (m.inactive) continue;
            double d2 = 0;
            for (int d = 0; d < dim_; ++d) { const double e = rs[d] - m.state[d]; d2 += e * e; }
            double dist = 0;
            if (d2 != 0) {
                auto it = term.find(m.depth);
                if (it == term.end()) { double Cm[16]; cov_of(m.state, Cm); it = term.emplace(m.depth, SymEig::cov_term(Cs, Cm, dim_)).first; }
                dist = std::fabs(d2 + it->second);
            }
            if (dist <= prm_.selection_radius && m.accCost < best_cost) { best_cost = m.accCost; best = i; }
            if (dist < nd) { nd = dist; nearest = i; }
        }
        return best >= 0 ? best : nearest;
    }

    // findClosestWitness (:150-178)
    int find_closest_witness(int node, bool) {
        const Motion& m = motions_[node];
        int closest = -1;
        double cd = std::numeric_limits<double>::infinity();
        for (int w = 0; w < (int)witnesses_.size(); ++w) {
            double d2 = 0;
            for (int d = 0; d < dim_; ++d) { const double e = witnesses_[w].state[d] - m.state[d]; d2 += e * e; }
            const double dist = d2 == 0 ? 0.0 : std::fabs(d2 + cov_term_nodes(witnesses_[w].depth, witnesses_[w].state, m.depth, m.state));
            if (dist < cd) { cd = dist; closest = w; }
        }
        if (closest < 0 || cd > prm_.pruning_radius) {
            witnesses_.push_back(Witness{m.state, m.depth, node});
            return (int)witnesses_.size() - 1;
        }
        return closest;
    }

    // the old representative becomes inactive; childless inactive nodes are dropped up the tree (:357-381)
    void prune(int old_rep) {
        motions_[old_rep].inactive = true;
        int cur = old_rep;
        while (cur >= 0 && motions_[cur].inactive && motions_[cur].numChildren == 0) {
            const int par = motions_[cur].parent;
            motions_[cur].numChildren = -1;          // tombstone (indices stay stable)
            if (par >= 0) motions_[par].numChildren--;
            cur = par;
        }
    }

    // root -> solution path, interpolated to one state per dt (PathControl::interpolate):
    // the intermediate states are re-propagated on the GPU in trajectory mode.
    Trajectory extract(int sol) {
        std::vector<int> chain;
        for (int c = sol; c >= 0; c = motions_[c].parent) chain.push_back(c);
        std::reverse(chain.begin(), chain.end());
        Trajectory tr;
        tr.push_back(BeliefState{dim_, motions_[chain[0]].state});
        for (size_t s = 1; s < chain.size(); ++s) {
            const Motion& m = motions_[chain[s]];
            const Motion& p = motions_[chain[s - 1]];
            std::vector<double> traj((size_t)m.steps * W_), fin(W_);
            int32_t st = m.steps, valid = 0;
            gpu_->check(cck_propagate_while_valid(gpu_->get(), 1, p.state.data(), 1, m.ctrl, 1, &st, fin.data(), 1, &valid, traj.data(), 1, m.steps), "interpolate");
            for (int t = 0; t < m.steps; ++t) tr.push_back(BeliefState{dim_, std::vector<double>(traj.begin() + (size_t)t * W_, traj.begin() + (size_t)(t + 1) * W_)});
        }
        return tr;
    }

    InstancePtr inst_;
    int robot_, dim_ = 2, W_ = 10;
    PlannerParams prm_;
    std::mt19937_64 rng_;
    GpuContextPtr gpu_;
    oc::SpaceInformationPtr si_;
    PlanValidityCheckerPtr pvc_;
    ob::RealVectorBounds sbounds_{0}, cbounds_{0};
    std::vector<double> start_;
    cck_goal_params goal_{};
    std::vector<ConstraintPtr> constraints_;
    std::vector<Motion> motions_;
    std::vector<Witness> witnesses_;
    std::unordered_map<uint64_t, double> cov_cache_;
    double max_eig_ = 10.0;                                            // :221
    PlannerStats stats_;
};

// ---------------------------------------------------------------------------
// K-CBS (src/Planners/KCBS.cpp:192-495): best-first search over constraint-tree nodes; each
// conflict spawns two children, one per involved agent, each re-planned under the union of that
// agent's constraints along the branch.  (The merge block is unimplemented in the reference.)
// ---------------------------------------------------------------------------
struct KCBSResult { bool solved = false; Plan plan; double cost = 0, seconds = 0; int nodes_expanded = 0, replans = 0; };

class KCBS {
public:
    KCBS(InstancePtr inst, const std::string& svc, const std::string& pvc, PlannerParams prm, double low_level_seconds = 5.0, int n_devices = 1)
        : inst_(std::move(inst)), ll_seconds_(low_level_seconds) {
        const int k = (int)inst_->getRobots().size();
        for (int r = 0; r < k; ++r) planners_.push_back(std::make_unique<BatchedBSST>(inst_, r, svc, pvc, prm, n_devices > 1 ? r % n_devices : 0));
    }
    BatchedBSST& planner(int r) { return *planners_[r]; }

    KCBSResult solve(double seconds) {
        struct Node { Plan plan; double cost; std::vector<ConstraintPtr> cons; };
        auto cmp = [](const std::shared_ptr<Node>& a, const std::shared_ptr<Node>& b) { return a->cost > b->cost; };
        std::priority_queue<std::shared_ptr<Node>, std::vector<std::shared_ptr<Node>>, decltype(cmp)> pq(cmp);
        const auto t0 = std::chrono::steady_clock::now();
        auto elapsed = [&] { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); };
        KCBSResult res;
        const int k = (int)planners_.size();
        auto root = std::make_shared<Node>();
        root->cost = 0;
        for (int r = 0; r < k; ++r) {                                   // :213-228
            planners_[r]->updateConstraints({});
            Trajectory t;
            while (t.empty() && elapsed() < seconds) t = planners_[r]->solve(ll_seconds_);
            if (t.empty()) { res.seconds = elapsed(); return res; }
            root->cost += 0.2 * (double)(t.size() - 1);                 // sum of control durations (KCBSNode::updatePlanAndCost)
            root->plan.push_back(std::move(t));
        }
        pq.push(root);
        const PlanValidityCheckerPtr& pv = planners_[0]->getPlanValidator();
        while (!pq.empty() && elapsed() < seconds) {
            auto cur = pq.top(); pq.pop();
            res.nodes_expanded++;
            const std::vector<ConflictPtr> confs = pv->validatePlan(cur->plan);      // :289
            if (confs.empty()) { res.solved = true; res.plan = cur->plan; res.cost = cur->cost; break; }
            const int agents[2] = {confs.front()->agent1Idx_, confs.front()->agent2Idx_};
            for (int a : agents) {                                       // :389-467 (two independent replans)
                auto child = std::make_shared<Node>(*cur);
                child->cons.push_back(pv->createConstraint(cur->plan, confs, a));
                std::vector<ConstraintPtr> mine;
                for (const auto& c : child->cons) if (c->getConstrainedAgent() == a) mine.push_back(c);
                planners_[a]->updateConstraints(mine);
                res.replans++;
                Trajectory t = planners_[a]->solve(ll_seconds_);
                if (t.empty()) continue;                                 // failed replans are dropped (the reference re-queues them at infinite cost)
                child->cost += 0.2 * ((double)t.size() - (double)child->plan[a].size());
                child->plan[a] = std::move(t);
                pq.push(child);
            }
        }
        res.seconds = elapsed();
        return res;
    }
private:
    InstancePtr inst_;
    double ll_seconds_;
    std::vector<std::unique_ptr<BatchedBSST>> planners_;
};

}  // namespace cck_b200
