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
#include <fstream>
#include <future>
#include <cmath>
#include <limits>
#include <queue>
#include <random>
#include <unordered_map>

#include "adapters.hpp"

namespace cck_b200 {

// ---------------------------------------------------------------------------
// RealVectorBeliefSpace::distance (src/Spaces/RealVectorBeliefSpace.cpp:16-62) drives both
// nearest-neighbour structures of the planner (nn_ for selectNode, witnesses_ for
// findClosestWitness).  Both live on the GPU as cck_tree objects: one brute-force batched launch
// answers the K selectNode queries of an iteration, one more the K nearest-witness queries, and a
// lower-triangular distance matrix of the batch lets the host replay the reference's SEQUENTIAL
// witness rules exactly (a candidate must see the witnesses created by earlier candidates of the
// same launch).
// ---------------------------------------------------------------------------
class DeviceTree {
public:
    DeviceTree(const GpuContextPtr& gpu, int dim) : gpu_(gpu) { gpu_->check(cck_tree_create(gpu_->get(), dim, &t_), "cck_tree_create"); }
    ~DeviceTree() { cck_tree_destroy(t_); }
    DeviceTree(const DeviceTree&) = delete;
    DeviceTree& operator=(const DeviceTree&) = delete;
    void clear() { gpu_->check(cck_tree_clear(t_), "cck_tree_clear"); }
    cck_tree* get() const { return t_; }
    int64_t size() const { return cck_tree_size(t_); }
    void append(int64_t n, const double* bel_soa, int64_t ld, const double* cost) { gpu_->check(cck_tree_append(t_, n, bel_soa, ld, cost, nullptr), "cck_tree_append"); }
    void deactivate(const std::vector<int32_t>& idx) {
        if (idx.empty()) return;
        std::vector<uint8_t> z(idx.size(), 0);
        gpu_->check(cck_tree_update(t_, (int64_t)idx.size(), idx.data(), nullptr, z.data()), "cck_tree_update");
    }
    void select(int64_t K, const double* samples, int64_t lds, double radius, int32_t* out) { gpu_->check(cck_tree_select(t_, K, samples, lds, radius, out, nullptr), "cck_tree_select"); }
    void nearest(int64_t K, const double* q, int64_t ldq, int32_t* idx, double* dist) { gpu_->check(cck_tree_nearest(t_, K, q, ldq, idx, dist), "cck_tree_nearest"); }
private:
    GpuContextPtr gpu_;
    cck_tree* t_ = nullptr;
};

struct PlannerParams {
    int K = 1024;                 // candidates per launch
    double goal_bias = 0.05, selection_radius = 0.2, pruning_radius = 0.1;
    double goal_threshold = 2.0, goal_p_safe = 0.95;   // ChanceConstrainedGoal(si, goal, 2.0, 0.95)  (OmplSetUp.cpp:239)
    int min_steps = 1, max_steps = 10;
    uint64_t seed = 1;
    // Stock SST deactivates the old representative of a witness and prunes childless inactive nodes.  The reference's file does
    // NOT: `inactive_` is only ever assigned inside `while (oldRep->inactive_ && ...)` (ConstraintRespectingBSST.cpp:388-404,
    // :517-531), so that loop is dead code, every node stays selectable and the tree is never pruned.  false = as written.
    bool prune = false;
};

// The solution as the reference's PathControl holds it (ConstraintRespectingBSST.cpp:327-349): tree-node states,
// the control held on each segment and its duration in seconds (steps * stepSize).
struct SegmentPath {
    int dim = 2;
    std::vector<std::vector<double>> states;     // n+1 belief records [W]
    std::vector<std::vector<double>> controls;   // n control vectors [dim]
    std::vector<double> durations;               // n durations (seconds)
};

struct PlannerStats { int64_t launches = 0, candidates = 0, accepted = 0, nodes = 0; double seconds = 0;
                      double t_sample = 0, t_select = 0, t_expand = 0, t_commit = 0, t_extract = 0;
                      double c_witness = 0, c_replay = 0, c_mirror = 0; int64_t c_surv = 0, c_fresh = 0; };   // where solve() spends its time

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
    struct Witness { int rep; };     // rep < 0: the witness a constraint-rejected candidate left behind (its rep_ is the reference's scratch motion)

    BatchedBSST(InstancePtr inst, int robot, const std::string& svc_name, const std::string& pvc_name, PlannerParams prm, int device = 0)
        : inst_(std::move(inst)), robot_(robot), prm_(prm), rng_(prm.seed + 7919 * (uint64_t)robot) {
        const Robot& r = inst_->getRobots()[robot_];
        dim_ = r.getDynamicsModel() == "Uncertain-Unicycle-Model" ? 4 : 2;
        W_ = belief_width(dim_);
        gpu_ = std::make_shared<GpuContext>(device);
        nodes_dev_ = std::make_unique<DeviceTree>(gpu_, dim_);
        wit_dev_ = std::make_unique<DeviceTree>(gpu_, dim_);
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
        // every node of this robot's trees descends from the start covariance: the expansion kernel reads (Sigma_s, Lambda_s)
        // of absolute step s from a table instead of re-running the Kalman step (cck_set_cov_table)
        gpu_->check(cck_set_cov_table(gpu_->get(), start_.data() + dim_, 4096), "cck_set_cov_table");
        goal_.gx = r.getGoalLocation().x_; goal_.gy = r.getGoalLocation().y_;
        goal_.threshold = prm_.goal_threshold; goal_.sc = cck_chi2_quantile_2dof(prm_.goal_p_safe);
        warmup();
    }

    // Part of the setup, like the reference's planner->setup(): one throw-away iteration at full batch size on constant
    // data (the random stream is untouched).  It loads the kernels (CUDA loads a module's functions lazily, ~ms each)
    // and sizes the device scratch and pinned staging buffers, so that solve() does not pay for them.
    void warmup() {
        const int K = prm_.K;
        std::vector<double> bel((size_t)W_ * K), ctrl((size_t)dim_ * K, 0.0), out((size_t)W_ * K), pcost(K, 0.0), cost(K), gdist(K), dist(K);
        for (int j = 0; j < K; ++j) for (int f = 0; f < W_; ++f) bel[(size_t)f * K + j] = start_[f];
        std::vector<int32_t> steps(K, 1), pstep(K, 0), valid(K), idx(K);
        std::vector<uint8_t> flags(K);
        const double zero = 0.0;
        pvc_->installConstraints({});
        nodes_dev_->append(1, start_.data(), 1, &zero);
        wit_dev_->append(1, start_.data(), 1, &zero);
        nodes_dev_->select(K, bel.data(), K, prm_.selection_radius, idx.data());
        gpu_->check(cck_expand_batch(gpu_->get(), K, bel.data(), K, ctrl.data(), K, steps.data(), pstep.data(), pcost.data(), &goal_, out.data(), K,
                                     valid.data(), cost.data(), gdist.data(), flags.data()), "cck_expand_batch (warm-up)");
        {
            const int S = std::min(K, 1024);
            std::vector<int32_t> near(S);
            std::vector<double> nd(S);
            std::vector<uint8_t> fr(S);
            gpu_->check(cck_tree_witness_batch(wit_dev_->get(), S, bel.data(), K, prm_.pruning_radius, nullptr, idx.data(), dist.data(), near.data(), nd.data(),
                                               fr.data()), "cck_tree_witness_batch (warm-up)");
        }
        nodes_dev_->deactivate({0});
        clear();
    }

    const oc::SpaceInformationPtr& getSpaceInformation() const { return si_; }
    const PlanValidityCheckerPtr& getPlanValidator() const { return pvc_; }
    const GpuContextPtr& gpu() const { return gpu_; }
    int dim() const { return dim_; }
    const PlannerStats& stats() const { return stats_; }
    bool startStateValid() const { return start_valid_ != 0; }         // false after solve(): OMPL's INVALID_START
    const SegmentPath& lastSolutionPath() const { return last_path_; }   // segments of the last trajectory solve() returned

    // ConstraintRespectingPlanner::updateConstraints + clear
    void updateConstraints(const std::vector<ConstraintPtr>& c) { constraints_ = c; clear(); }
    void clear() { motions_.clear(); witnesses_.clear(); nodes_dev_->clear(); wit_dev_->clear(); }

    // solve(seconds): returns the interpolated trajectory (one state per dt) or empty
    Trajectory solve(double seconds, double* cost_out = nullptr) {
        const auto t0 = std::chrono::steady_clock::now();
        pvc_->installConstraints(constraints_);
        // PlannerInputStates::nextStart() (ConstraintRespectingBSST.cpp:186): a start state the validity checker rejects is
        // skipped and the planner returns INVALID_START ("There are no valid initial states!")
        if (start_valid_ < 0) {
            uint8_t v = 0;
            gpu_->check(cck_is_valid(gpu_->get(), 1, start_.data(), 1, &v), "cck_is_valid (start state)");
            start_valid_ = v ? 1 : 0;
        }
        if (!start_valid_) return {};
        if (motions_.empty()) {
            Motion m; m.state = start_;
            motions_.push_back(m);
            const double zero = 0.0;
            nodes_dev_->append(1, start_.data(), 1, &zero);      // one record: SoA with ld = 1 is the record itself
            witnesses_.push_back(Witness{0});
            wit_dev_->append(1, start_.data(), 1, &zero);
        }
        // the reference's satisfiesConstraints walks the whole root-to-node path, start state (time 0) included: a constraint at
        // step 0 that the start state violates rejects every candidate, so the low-level search cannot succeed
        if (!constraints_.empty()) {
            Trajectory st{BeliefState{dim_, start_}};
            if (!pvc_->satisfiesConstraintsInstalled(st)) { stats_.seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); return {}; }
        }
        const int K = prm_.K;
        std::vector<double> bel((size_t)W_ * K), ctrl((size_t)dim_ * K), out((size_t)W_ * K), pcost(K), cost(K), gdist(K), samples((size_t)W_ * K);
        std::vector<int32_t> steps(K), pstep(K), valid(K), parent(K);
        std::vector<uint8_t> flags(K);
        std::uniform_real_distribution<double> U01(0.0, 1.0);
        int solution = -1;
        while (solution < 0) {
            if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > seconds) break;
            const auto ta = std::chrono::steady_clock::now();
            // ---- sample K (state, control, duration) triples (ConstraintRespectingBSST.cpp:226-247) ----
            std::fill(samples.begin(), samples.end(), 0.0);
            for (int j = 0; j < K; ++j) {
                double rs[4];
                if (U01(rng_) < prm_.goal_bias) { rs[0] = goal_.gx; rs[1] = goal_.gy; for (int d = 2; d < dim_; ++d) rs[d] = uni(sbounds_.low[d], sbounds_.high[d]); }
                else for (int d = 0; d < dim_; ++d) rs[d] = uni(sbounds_.low[d], sbounds_.high[d]);
                for (int d = 0; d < dim_; ++d) samples[(size_t)d * K + j] = rs[d];
                // the sample carries sigma = diag(d0, d1, 0.01..), lambda = 0.01 I (allocState default + :235-236)
                for (int d = 0; d < dim_; ++d) { samples[(size_t)(dim_ + d * dim_ + d) * K + j] = 0.01; samples[(size_t)(dim_ + dim_ * dim_ + d * dim_ + d) * K + j] = 0.01; }
                samples[(size_t)(dim_ + 0) * K + j] = U01(rng_) * max_eig_;
                samples[(size_t)(dim_ + dim_ + 1) * K + j] = U01(rng_) * max_eig_;
                for (int d = 0; d < dim_; ++d) ctrl[(size_t)d * K + j] = uni(cbounds_.low[d], cbounds_.high[d]);
                steps[j] = prm_.min_steps + (int)(rng_() % (uint64_t)(prm_.max_steps - prm_.min_steps + 1));
            }
            const auto tb = std::chrono::steady_clock::now();
            // ---- ONE launch: selectNode for all K samples (:119-148) ----
            nodes_dev_->select(K, samples.data(), K, prm_.selection_radius, parent.data());
            stats_.launches++;
            for (int j = 0; j < K; ++j) {
                const Motion& m = motions_[parent[j]];
                for (int f = 0; f < W_; ++f) bel[(size_t)f * K + j] = m.state[f];
                pstep[j] = m.depth; pcost[j] = m.accCost;
            }
            const auto tc = std::chrono::steady_clock::now();
            // ---- ONE launch: K x (<=10 steps) propagate + check (+ constraints, cost, goal) ----
            gpu_->check(cck_expand_batch(gpu_->get(), K, bel.data(), K, ctrl.data(), K, steps.data(), pstep.data(), pcost.data(), &goal_, out.data(), K,
                                         valid.data(), cost.data(), gdist.data(), flags.data()), "cck_expand_batch");
            stats_.launches++; stats_.candidates += K;
            const auto td = std::chrono::steady_clock::now();
            // ---- commit in sample order with the reference's witness rules (:250-326) ----
            solution = commit(K, out, ctrl, steps, parent, cost, flags);
            const auto te = std::chrono::steady_clock::now();
            auto sec = [](auto a, auto b) { return std::chrono::duration<double>(b - a).count(); };
            stats_.t_sample += sec(ta, tb); stats_.t_select += sec(tb, tc); stats_.t_expand += sec(tc, td); stats_.t_commit += sec(td, te);
        }
        stats_.nodes = (int64_t)motions_.size();
        stats_.seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (solution < 0) return {};
        if (cost_out) *cost_out = motions_[solution].accCost;
        const auto tx = std::chrono::steady_clock::now();
        Trajectory traj = extract(solution);
        stats_.t_extract += std::chrono::duration<double>(std::chrono::steady_clock::now() - tx).count();
        return traj;
    }

private:
    double uni(double lo, double hi) { return lo + (hi - lo) * std::uniform_real_distribution<double>(0.0, 1.0)(rng_); }

    // Deviations of the batched commit from the one-candidate-at-a-time reference (documented in DESIGN.md section 7): all K
    // parents of a launch are selected before any of its candidates is committed, so a candidate cannot have a node created by
    // the same launch as its parent (and, with prm.prune, a node deactivated by an earlier candidate of the launch can still
    // gain a child); constraints at absolute step 0 are checked once per solve() on the start state, not once per candidate.
    // Commit the survivors of one launch.  The reference handles one candidate at a time: findClosestWitness
    // (:150-178) sees every witness created so far, including those of earlier candidates of the same batch.
    // Here: cck_tree_witness_batch finds each survivor's nearest PRE-BATCH witness, runs the greedy new-witness scan on the
    // device and returns each survivor's nearest same-launch witness; the loop below replays the remaining bookkeeping
    // (representatives, pruning, tree growth) with O(1) work per survivor.
    int commit(int K, const std::vector<double>& out, const std::vector<double>& ctrl, const std::vector<int32_t>& steps,
               const std::vector<int32_t>& parent, const std::vector<double>& cost, const std::vector<uint8_t>& flags) {
        std::vector<int> surv;
        for (int j = 0; j < K; ++j) if (flags[j] & 1) surv.push_back(j);          // propCd == cd
        int solution = -1;
        const int kMaxChunk = 1024;      // cck_tree_witness_batch resolves up to 1024 survivors per call
        for (size_t c0 = 0; c0 < surv.size() && solution < 0; c0 += kMaxChunk) {
            const int S = (int)std::min<size_t>(kMaxChunk, surv.size() - c0);
            std::vector<double> q((size_t)W_ * S), qcost(S);
            for (int s = 0; s < S; ++s) { const int j = surv[c0 + s]; for (int f = 0; f < W_; ++f) q[(size_t)f * S + s] = out[(size_t)f * K + j]; qcost[s] = cost[j]; }
            auto now = [] { return std::chrono::steady_clock::now(); };
            auto secs = [](auto a, auto b) { return std::chrono::duration<double>(b - a).count(); };
            const auto t1 = now();
            // ONE call: nearest existing witness of every survivor, which survivors become new witnesses (greedy in sample
            // order, on the device) and each survivor's nearest new witness among its predecessors.  Every survivor is
            // eligible: findClosestWitness (:258) runs BEFORE satisfiesConstraints (:300), so a candidate the constraints
            // reject still leaves its new witness behind (with rep_ == rmotion, the scratch motion: rep -1 here).
            std::vector<int32_t> widx(S), near(S);
            std::vector<double> wdist(S), near_dist(S);
            std::vector<uint8_t> fresh_dev(S);
            gpu_->check(cck_tree_witness_batch(wit_dev_->get(), S, q.data(), S, prm_.pruning_radius, nullptr,
                                               widx.data(), wdist.data(), near.data(), near_dist.data(), fresh_dev.data()), "cck_tree_witness_batch");
            stats_.launches += 5;
            const auto t2 = now();
            std::vector<int> wit_of(S, -1);               // survivor -> index of the witness it created in this chunk
            std::vector<std::pair<int, int>> fresh;       // (survivor s, witness index) of witnesses created in this chunk
            std::vector<int32_t> deact;
            std::vector<int> new_nodes;                   // survivors appended to the tree, in order
            const size_t wit_before = witnesses_.size();
            for (int s = 0; s < S && solution < 0; ++s) {
                const int j = surv[c0 + s];
                int w = widx[s];
                double cd = w >= 0 ? wdist[s] : std::numeric_limits<double>::infinity();
                // the reference's scan over the witnesses created so far: the device found the first nearest one below
                // min(cd, radius), which is the only one that can change (w, cd) or the new-witness decision
                if (near[s] >= 0 && wit_of[near[s]] >= 0 && near_dist[s] < cd) { cd = near_dist[s]; w = wit_of[near[s]]; }
                const int idx = (int)motions_.size();
                bool is_new = false;
                if (w < 0 || cd > prm_.pruning_radius) { witnesses_.push_back(Witness{-1}); w = (int)witnesses_.size() - 1; fresh.emplace_back(s, w); wit_of[s] = w; is_new = true; }
                if (is_new != (fresh_dev[s] != 0)) throw std::logic_error("witness batch: device and host disagree on a new witness");
                Witness& wit = witnesses_[w];
                // :261 `rep_ == rmotion || cost better`: true for a fresh witness and for one whose creator was rejected below
                if (!(wit.rep < 0 || cost[j] < motions_[wit.rep].accCost)) continue;
                if (!constraints_.empty() && !(flags[j] & 2)) continue;         // satisfiesConstraints failed (:300): the witness stays, without a representative
                Motion m;
                m.state.resize(W_);
                for (int f = 0; f < W_; ++f) m.state[f] = out[(size_t)f * K + j];
                for (int d = 0; d < dim_; ++d) m.ctrl[d] = ctrl[(size_t)d * K + j];
                m.steps = steps[j]; m.parent = parent[j]; m.depth = motions_[parent[j]].depth + steps[j]; m.accCost = cost[j];
                motions_.push_back(m);
                new_nodes.push_back(s);
                if (wit.rep >= 0 && prm_.prune) prune(wit.rep, deact);
                wit.rep = idx;                                                   // linkRep (:301)
                motions_[parent[j]].numChildren++;
                stats_.accepted++;
                const double c00 = m.state[dim_] + m.state[dim_ + dim_ * dim_], c11 = m.state[dim_ + dim_ + 1] + m.state[dim_ + dim_ * dim_ + dim_ + 1];
                if (c00 > max_eig_) max_eig_ = c00; else if (c11 > max_eig_) max_eig_ = c11;     // :313-320
                if (flags[j] & 4) solution = idx;                                // goal->isSatisfied (:326)
            }
            const auto t4 = now();
            // mirror the chunk's changes on the device: new nodes, new witnesses, deactivated representatives
            if (!new_nodes.empty()) {
                const int n = (int)new_nodes.size();
                std::vector<double> nb((size_t)W_ * n), nc(n);
                for (int a = 0; a < n; ++a) { for (int f = 0; f < W_; ++f) nb[(size_t)f * n + a] = q[(size_t)f * S + new_nodes[a]]; nc[a] = qcost[new_nodes[a]]; }
                nodes_dev_->append(n, nb.data(), n, nc.data());
            }
            if (witnesses_.size() > wit_before) {
                const int n = (int)fresh.size();
                std::vector<double> wb((size_t)W_ * n), wc(n, 0.0);
                for (int a = 0; a < n; ++a) for (int f = 0; f < W_; ++f) wb[(size_t)f * n + a] = q[(size_t)f * S + fresh[a].first];
                wit_dev_->append(n, wb.data(), n, wc.data());
            }
            nodes_dev_->deactivate(deact);
            const auto t5 = now();
            stats_.c_witness += secs(t1, t2); stats_.c_replay += secs(t2, t4); stats_.c_mirror += secs(t4, t5);
            stats_.c_surv += S; stats_.c_fresh += (int64_t)fresh.size();
        }
        return solution;
    }

    // the old representative becomes inactive; childless inactive nodes are dropped up the tree (:357-381)
    void prune(int old_rep, std::vector<int32_t>& deact) {
        motions_[old_rep].inactive = true;
        deact.push_back(old_rep);          // mirrored on the device after this chunk's nodes are appended
        int cur = old_rep;
        while (cur >= 0 && motions_[cur].inactive && motions_[cur].numChildren == 0) {
            const int par = motions_[cur].parent;
            motions_[cur].numChildren = -1;          // tombstone (indices stay stable)
            if (par >= 0) motions_[par].numChildren--;
            cur = par;
        }
    }

    // root -> solution path, interpolated to one state per dt (PathControl::interpolate,
    // ConstraintRespectingBSST.cpp:283-298): ALL segments are re-propagated by ONE trajectory-mode launch
    // (segment s = belief s of the batch, its parent state, held control and step count).
    Trajectory extract(int sol) {
        std::vector<int> chain;
        for (int c = sol; c >= 0; c = motions_[c].parent) chain.push_back(c);
        std::reverse(chain.begin(), chain.end());
        Trajectory tr;
        tr.push_back(BeliefState{dim_, motions_[chain[0]].state});
        const int n = (int)chain.size() - 1;
        last_path_ = SegmentPath{};
        last_path_.dim = dim_;
        for (size_t s = 0; s < chain.size(); ++s) {
            const Motion& m = motions_[chain[s]];
            last_path_.states.push_back(m.state);
            if (s > 0) { last_path_.controls.emplace_back(m.ctrl, m.ctrl + dim_); last_path_.durations.push_back(0.2 * m.steps); }
        }
        if (n <= 0) return tr;
        int T = 0;
        for (int s = 1; s <= n; ++s) T = std::max(T, motions_[chain[s]].steps);
        std::vector<double> bel((size_t)W_ * n), ctrl((size_t)dim_ * n), fin((size_t)W_ * n), traj((size_t)T * W_ * n);
        std::vector<int32_t> st(n), valid(n);
        for (int s = 0; s < n; ++s) {
            const Motion& m = motions_[chain[s + 1]];
            const Motion& p = motions_[chain[s]];
            for (int f = 0; f < W_; ++f) bel[(size_t)f * n + s] = p.state[f];
            for (int d = 0; d < dim_; ++d) ctrl[(size_t)d * n + s] = m.ctrl[d];
            st[s] = m.steps;
        }
        gpu_->check(cck_propagate_while_valid(gpu_->get(), n, bel.data(), n, ctrl.data(), n, st.data(), fin.data(), n, valid.data(), traj.data(), n, T), "interpolate");
        stats_.launches++;
        for (int s = 0; s < n; ++s)
            for (int t = 0; t < st[s]; ++t) {
                std::vector<double> v(W_);
                for (int f = 0; f < W_; ++f) v[f] = traj[((size_t)t * W_ + f) * n + s];
                tr.push_back(BeliefState{dim_, std::move(v)});
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
    std::unique_ptr<DeviceTree> nodes_dev_, wit_dev_;
    int start_valid_ = -1;                                             // isValid(start), evaluated once (-1: not yet)
    double max_eig_ = 10.0;                                            // :221
    PlannerStats stats_;
    SegmentPath last_path_;
};

// ---------------------------------------------------------------------------
// K-CBS (src/Planners/KCBS.cpp:192-495): best-first search over constraint-tree nodes; each
// conflict spawns two children, one per involved agent, each re-planned under the union of that
// agent's constraints along the branch.  (The merge block is unimplemented in the reference.)
// ---------------------------------------------------------------------------
struct KCBSResult { bool solved = false; Plan plan; std::vector<SegmentPath> paths; double cost = 0, seconds = 0; int nodes_expanded = 0, replans = 0;
                    double t_root = 0, t_validate = 0; };

class KCBS {
public:
    KCBS(InstancePtr inst, const std::string& svc, const std::string& pvc, PlannerParams prm, double low_level_seconds = 5.0, int n_devices = 1)
        : inst_(std::move(inst)), ll_seconds_(low_level_seconds) {
        const int k = (int)inst_->getRobots().size();
        for (int r = 0; r < k; ++r) planners_.push_back(std::make_unique<BatchedBSST>(inst_, r, svc, pvc, prm, n_devices > 1 ? r % n_devices : 0));
    }
    BatchedBSST& planner(int r) { return *planners_[r]; }

    KCBSResult solve(double seconds) {
        struct Node { Plan plan; std::vector<SegmentPath> paths; double cost; std::vector<ConstraintPtr> cons; };
        auto cmp = [](const std::shared_ptr<Node>& a, const std::shared_ptr<Node>& b) { return a->cost > b->cost; };
        std::priority_queue<std::shared_ptr<Node>, std::vector<std::shared_ptr<Node>>, decltype(cmp)> pq(cmp);
        const auto t0 = std::chrono::steady_clock::now();
        auto elapsed = [&] { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); };
        KCBSResult res;
        const int k = (int)planners_.size();
        auto root = std::make_shared<Node>();
        root->cost = 0;
        // :213-228.  The k root plans are independent (one planner, context, stream and device tree per robot, own
        // random stream), so they are computed concurrently; results are consumed in robot order.
        std::vector<std::future<Trajectory>> roots;
        for (int r = 0; r < k; ++r) {
            planners_[r]->updateConstraints({});
            BatchedBSST* pl = planners_[r].get();
            const double secs = ll_seconds_;
            roots.push_back(std::async(k > 1 ? std::launch::async : std::launch::deferred, [pl, secs, seconds, &elapsed] {
                Trajectory t;
                while (t.empty() && elapsed() < seconds && pl->startStateValid()) t = pl->solve(secs);   // INVALID_START is final
                return t;
            }));
        }
        std::vector<Trajectory> root_traj;
        for (int r = 0; r < k; ++r) root_traj.push_back(roots[r].get());      // join every task before any early return
        res.t_root = elapsed();
        for (int r = 0; r < k; ++r) {
            Trajectory& t = root_traj[r];
            if (t.empty()) { res.seconds = elapsed(); return res; }
            root->cost += 0.2 * (double)(t.size() - 1);                 // sum of control durations (KCBSNode::updatePlanAndCost)
            root->plan.push_back(std::move(t));
            root->paths.push_back(planners_[r]->lastSolutionPath());
        }
        pq.push(root);
        const PlanValidityCheckerPtr& pv = planners_[0]->getPlanValidator();
        while (!pq.empty() && elapsed() < seconds) {
            auto cur = pq.top(); pq.pop();
            res.nodes_expanded++;
            const double tv = elapsed();
            const std::vector<ConflictPtr> confs = pv->validatePlan(cur->plan);      // :289
            res.t_validate += elapsed() - tv;
            if (confs.empty()) { res.solved = true; res.plan = cur->plan; res.paths = cur->paths; res.cost = cur->cost; break; }
            const int agents[2] = {confs.front()->agent1Idx_, confs.front()->agent2Idx_};
            // :389-467: the two child replans are independent (different agents, each planner owns its context,
            // stream and device trees), so they run concurrently — on different GPUs when n_devices > 1
            std::shared_ptr<Node> child[2];
            std::future<Trajectory> job[2];
            for (int i = 0; i < 2; ++i) {
                const int a = agents[i];
                child[i] = std::make_shared<Node>(*cur);
                child[i]->cons.push_back(pv->createConstraint(cur->plan, confs, a));
                std::vector<ConstraintPtr> mine;
                for (const auto& c : child[i]->cons) if (c->getConstrainedAgent() == a) mine.push_back(c);
                planners_[a]->updateConstraints(mine);
                res.replans++;
                BatchedBSST* pl = planners_[a].get();
                const double secs = ll_seconds_;
                if (agents[0] != agents[1]) job[i] = std::async(std::launch::async, [pl, secs] { return pl->solve(secs); });
                else { std::promise<Trajectory> pr; pr.set_value(pl->solve(secs)); job[i] = pr.get_future(); }
            }
            for (int i = 0; i < 2; ++i) {
                const int a = agents[i];
                Trajectory t = job[i].get();
                if (t.empty()) continue;                                 // failed replans are dropped (the reference re-queues them at infinite cost)
                child[i]->cost += 0.2 * ((double)t.size() - (double)child[i]->plan[a].size());
                child[i]->plan[a] = std::move(t);
                child[i]->paths[a] = planners_[a]->lastSolutionPath();
                pq.push(child[i]);
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

// exportBeliefPlan (src/utils/postProcess.cpp:381-417): solutions/<name>/agentNN.txt in OMPL's
// PathControl::printAsMatrix layout (first row: start state and zeros for control + duration; then one row per
// segment: node state, control, duration) and agentNN_covs.txt with Sigma + Lambda (row-major) of every node,
// both with the default ostream precision the reference writes.
inline void exportBeliefPlan(const std::vector<SegmentPath>& plan, const std::string& dir) {
    for (size_t i = 0; i < plan.size(); ++i) {
        const SegmentPath& p = plan[i];
        const std::string num = (i < 10 ? "0" : "") + std::to_string(i);
        std::ofstream file(dir + "/agent" + num + ".txt");
        const int d = p.dim;
        for (int f = 0; f < d; ++f) file << p.states[0][f] << " ";
        for (int f = 0; f < d + 1; ++f) file << "0 ";
        file << '\n';
        for (size_t s = 0; s < p.controls.size(); ++s) {
            for (int f = 0; f < d; ++f) file << p.states[s + 1][f] << " ";
            for (int f = 0; f < d; ++f) file << p.controls[s][f] << ' ';
            file << p.durations[s] << '\n';
        }
        std::ofstream covs(dir + "/agent" + num + "_covs.txt");
        for (const auto& st : p.states) {
            for (int f = 0; f < d * d; ++f) covs << st[d + f] + st[d + d * d + f] << " ";
            covs << std::endl;
        }
    }
}

}  // namespace cck_b200
