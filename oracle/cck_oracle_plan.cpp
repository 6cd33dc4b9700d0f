// cck_oracle_plan.cpp — TEST INFRASTRUCTURE / CPU BASELINE: the reference's planner loop on the CPU oracle, ONE candidate at
// a time, for the solve-time half of BASELINE.json's metric ("CPU-oracle planner vs GPU-batched planner").
//
//   low level   ConstraintRespectingBSST::solve  (src/Planners/ConstraintRespectingBSST.cpp:180-537), selectNode (:119-148),
//               findClosestWitness (:150-178), SST parameters of include/Planners/ConstraintRespectingBSST.h:232-238
//               (goal bias 0.05, selection radius 0.2, pruning radius 0.1), control duration uniform in [1, 10] steps and
//               controls uniform in the control bounds (OmplSetUp.cpp:203-207, 225, 275-289)
//   high level  KCBS::solve  (src/Planners/KCBS.cpp:192-495): best-first over constraint-tree nodes, one child per agent of the
//               first conflict, each re-planned under that agent's constraints along the branch
//
// Followed AS WRITTEN, including two quirks that change behaviour:
//   * old representatives are never deactivated: `inactive_` is only ever set inside `while (oldRep->inactive_ && ...)`
//     (:388-404, :517-531), so the loop body is dead code and the tree is never pruned (--prune restores stock SST);
//   * a candidate that creates a NEW witness but is then rejected by satisfiesConstraints (:300) leaves that witness behind
//     with rep_ == rmotion, the scratch motion; every later candidate that lands on it passes the `rep_ == rmotion` test
//     (:261) and is accepted whatever its cost.
// Deliberately FASTER than the reference where that cannot change a decision, so that it is a conservative baseline:
//   * no heap allocation per state, no re-propagation of the root-to-node path for the constraint check (:283-298: the
//     reference interpolates the whole path on every accepted candidate) - every node keeps the per-step states of its own
//     segment and a constraint can only fire on the new segment (the parent chain was checked when it was built);
//   * nearest-neighbour queries (OMPL's GNAT in the reference) are exact scans pruned by the lower bound
//     distance >= |mu1 - mu2|^2 over a uniform grid of the means: same answers as a linear scan.
// The random stream cannot match the reference's (it is never seeded there), so results are statistical, like the GPU planner's.
//
//   cck_oracle_plan -m x.map -s x.scen -k 2 [--svc Blackmore] [--pvc Blackmore] [-p 0.9] [-t 300] [--lltime 20] [--seed 1] [--prune] [-o plan.txt]
// Prints one JSON line in the format of host/cck_plan.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <memory>
#include <queue>
#include <random>
#include <sstream>
#include <unordered_map>

#include "cck_oracle.hpp"

using namespace cck_oracle;

namespace {

double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

struct Params {
    double goal_bias = 0.05, selection_radius = 0.2, pruning_radius = 0.1, goal_threshold = 2.0, goal_p_safe = 0.95;
    int min_steps = 1, max_steps = 10;
    bool prune = false;      // stock-SST deactivation of old representatives (NOT what the reference file does)
    uint64_t seed = 1;
};

// exact nearest-neighbour scans over belief records, pruned by distance >= squared mean distance
template <int D>
class MeanGrid {
public:
    explicit MeanGrid(double cell = 0.5) : cell_(cell) {}
    void clear() { cells_.clear(); pts_.clear(); }
    void add(const Belief<D>& b) {
        const int id = (int)pts_.size();
        pts_.push_back(b);
        cells_[key(cx(b.v[0]), cx(b.v[1]))].push_back(id);
    }
    size_t size() const { return pts_.size(); }
    const Belief<D>& at(int i) const { return pts_[i]; }
    // argmin over items with ok(i) of dist(i); ties to the lowest index (a linear scan in insertion order would do the same)
    template <class Dist, class Ok>
    int nearest(const double* q, const Dist& dist, const Ok& ok, double* dout) const {
        int best = -1;
        double bd = HUGE_VAL;
        const int qx = cx(q[0]), qy = cx(q[1]);
        for (int ring = 0;; ++ring) {
            // every item in a cell at Chebyshev ring distance `ring` is at least (ring - 1) * cell away in the mean
            const double lb = ring > 0 ? (ring - 1) * cell_ : 0.0;
            if (best >= 0 && lb * lb > bd) break;
            if (ring > max_ring_) break;
            for (int dx = -ring; dx <= ring; ++dx)
                for (int dy = -ring; dy <= ring; ++dy) {
                    if (std::max(std::abs(dx), std::abs(dy)) != ring) continue;
                    auto it = cells_.find(key(qx + dx, qy + dy));
                    if (it == cells_.end()) continue;
                    for (int i : it->second) {
                        if (!ok(i)) continue;
                        const double d = dist(i);
                        if (d < bd || (d == bd && i < best)) { bd = d; best = i; }
                    }
                }
        }
        if (dout) *dout = bd;
        return best;
    }
    // items with ok(i) and dist(i) <= radius
    template <class Dist, class Ok, class F>
    void within(const double* q, double radius, const Dist& dist, const Ok& ok, const F& f) const {
        const int reach = (int)std::ceil(std::sqrt(radius) / cell_) + 1;
        const int qx = cx(q[0]), qy = cx(q[1]);
        for (int dx = -reach; dx <= reach; ++dx)
            for (int dy = -reach; dy <= reach; ++dy) {
                auto it = cells_.find(key(qx + dx, qy + dy));
                if (it == cells_.end()) continue;
                for (int i : it->second) if (ok(i)) { const double d = dist(i); if (d <= radius) f(i, d); }
            }
    }
private:
    int cx(double v) const { return (int)std::floor(v / cell_); }
    static long long key(int a, int b) { return (long long)(((unsigned long long)(unsigned)a << 32) | (unsigned)b); }
    double cell_;
    int max_ring_ = 400;
    std::unordered_map<long long, std::vector<int>> cells_;
    std::vector<Belief<D>> pts_;
};

struct Stats { long long iterations = 0, accepted = 0, nodes = 0; double seconds = 0; };

template <int D, class Model>
class OracleBSST {
public:
    struct Motion { Belief<D> state; double ctrl[4] = {0, 0, 0, 0}; int steps = 0, parent = -1, depth = 0; double cost = 0; bool inactive = false; int children = 0;
                    std::vector<Belief<D>> seg; };
    OracleBSST(const Instance& inst, int robot, int svc_kind, const Pvc& pvc, Params prm)
        : inst_(inst), robot_(robot), pvc_(pvc), prm_(prm), rng_(prm.seed + 7919 * (uint64_t)robot) {
        for (int d = 0; d < 4; ++d) { slow_[d] = shigh_[d] = clow_[d] = chigh_[d] = 0; }
        slow_[0] = -1; shigh_[0] = inst.x_max; slow_[1] = -1; shigh_[1] = inst.y_max;          // OmplSetUp.cpp:194-200 / 263-289
        if (D == 4) { slow_[2] = -M_PI; shigh_[2] = M_PI; slow_[3] = 0.01; shigh_[3] = 10.0; for (int d = 0; d < 4; ++d) { clow_[d] = slow_[d]; chigh_[d] = shigh_[d]; } }
        else { clow_[0] = clow_[1] = -1.0; chigh_[0] = chigh_[1] = 1.0; }
        const double accep = svc_kind == SVC_CHISQUARED ? 0.9 : inst.p_safe_obs;                  // :213-222
        svc_ = make_svc(svc_kind, D, slow_, shigh_, inst.obstacles, inst.robots[robot].shape, accep);
        belief_default(start_);
        for (int d = 0; d < D; ++d) start_.v[d] = 0;
        start_.v[0] = inst.robots[robot].sx; start_.v[1] = inst.robots[robot].sy;
        if (D == 4) { start_.v[2] = 0.0; start_.v[3] = 0.1; }                                    // :315-320
        gx_ = inst.robots[robot].gx; gy_ = inst.robots[robot].gy;
    }
    Model model;
    Stats stats;
    void set_constraints(const std::vector<BeliefConstraint<D>>& c) { cons_ = c; motions_.clear(); wit_rep_.clear(); nodes_.clear(); wits_.clear(); }

    // -> interpolated trajectory (one state per dt) or empty
    std::vector<Belief<D>> solve(double seconds) {
        const double t0 = now_s();
        std::uniform_real_distribution<double> U(0.0, 1.0);
        if (motions_.empty()) {
            // nextStart(): the start state must be valid (:186-195); with constraints the reference checks every state of the
            // path, the start included (time 0)
            double c2[4];
            cov_block2(start_, c2);
            if (!svc_is_valid(svc_, start_.v, c2)) return {};
            Motion m; m.state = start_;
            motions_.push_back(m); nodes_.add(start_);
            wits_.add(start_); wit_rep_.push_back(0);
            start_ok_ = cons_.empty() || satisfies_constraints<D>(pvc_, std::vector<Belief<D>>{start_}, 0.2, cons_);
        }
        double max_eig = 10.0;                                                                    // :221
        int solution = -1;
        while (solution < 0) {
            if ((stats.iterations & 63) == 0 && now_s() - t0 > seconds) break;
            stats.iterations++;
            // sample (:226-236): uniform or goal-biased mean, Sigma(0,0), Sigma(1,1) = U * max_eigenvalue, rest = allocState default
            Belief<D> rs;
            belief_default(rs);
            if (U(rng_) < prm_.goal_bias) { rs.v[0] = gx_; rs.v[1] = gy_; for (int d = 2; d < D; ++d) rs.v[d] = uni(slow_[d], shigh_[d]); }
            else for (int d = 0; d < D; ++d) rs.v[d] = uni(slow_[d], shigh_[d]);
            rs.S[0] = U(rng_) * max_eig; rs.S[D + 1] = U(rng_) * max_eig;
            const int nm = select_node(rs);                                                       // :239
            double ctrl[4];
            for (int d = 0; d < D; ++d) ctrl[d] = uni(clow_[d], chigh_[d]);                       // :244
            const int cd = prm_.min_steps + (int)(rng_() % (uint64_t)(prm_.max_steps - prm_.min_steps + 1));   // :247
            Belief<D> res;
            std::vector<Belief<D>> seg(cd);
            const int prop = propagate_while_valid<D>(model, svc_, 0.2, motions_[nm].state, ctrl, cd, res, nullptr, seg.data());   // :248
            if (prop != cd) continue;                                                             // :250
            double inc = 0;
            for (int d = 0; d < D; ++d) { const double e = motions_[nm].state.v[d] - res.v[d]; inc += e * e; }
            const double cost = motions_[nm].cost + std::sqrt(inc);                               // :252-256
            // findClosestWitness (:150-178): distance(witness, node)
            double wd;
            int w = wits_.nearest(res.v, [&](int i) { return belief_distance((const double*)&wits_.at(i), (const double*)&res, D); }, [](int) { return true; }, &wd);
            bool is_new = false;
            if (w < 0 || wd > prm_.pruning_radius) { wits_.add(res); wit_rep_.push_back(-1); w = (int)wit_rep_.size() - 1; is_new = true; }
            // rep_ == rmotion holds for a fresh witness AND for one a constraint-rejected candidate left behind (rep -1)
            if (!(wit_rep_[w] < 0 || cost < motions_[wit_rep_[w]].cost)) continue;                // :261
            const int old_rep = wit_rep_[w];
            (void)is_new;
            if (!cons_.empty()) {
                // satisfiesConstraints on the interpolated root-to-node path (:283-300): only the new segment can fail
                if (!start_ok_ || !segment_ok(seg, motions_[nm].depth + 1)) continue;
            }
            Motion m;
            m.state = res; m.steps = cd; m.parent = nm; m.depth = motions_[nm].depth + cd; m.cost = cost; m.seg = std::move(seg);
            for (int d = 0; d < D; ++d) m.ctrl[d] = ctrl[d];
            const int idx = (int)motions_.size();
            motions_.push_back(std::move(m)); nodes_.add(res);
            motions_[nm].children++;
            wit_rep_[w] = idx;                                                                    // linkRep
            stats.accepted++;
            const double c00 = res.S[0] + res.L[0], c11 = res.S[D + 1] + res.L[D + 1];
            if (c00 > max_eig) max_eig = c00; else if (c11 > max_eig) max_eig = c11;              // :313-320
            double c2[4], gd;
            cov_block2(res, c2);
            if (goal_is_satisfied(res.v, c2, gx_, gy_, prm_.goal_threshold, prm_.goal_p_safe, &gd)) solution = idx;   // :326
            if (prm_.prune && old_rep >= 0) {                                                     // stock SST only; dead code in the reference
                int cur = old_rep;
                motions_[cur].inactive = true;
                while (cur >= 0 && motions_[cur].inactive && motions_[cur].children == 0) {
                    const int par = motions_[cur].parent;
                    motions_[cur].children = -1;
                    if (par >= 0) motions_[par].children--;
                    cur = par;
                }
            }
        }
        stats.nodes = (long long)motions_.size();
        stats.seconds += now_s() - t0;
        if (solution < 0) return {};
        cost_ = motions_[solution].cost;
        std::vector<int> chain;
        for (int c = solution; c >= 0; c = motions_[c].parent) chain.push_back(c);
        std::reverse(chain.begin(), chain.end());
        std::vector<Belief<D>> tr{motions_[chain[0]].state};
        for (size_t s = 1; s < chain.size(); ++s) for (const auto& b : motions_[chain[s]].seg) tr.push_back(b);
        return tr;
    }
    double last_cost() const { return cost_; }
    // OMPL answers INVALID_START when the only start state fails the validity checker (:186-195)
    bool start_valid() const { double c2[4]; cov_block2(start_, c2); return svc_is_valid(svc_, start_.v, c2); }

private:
    double uni(double lo, double hi) { return lo + (hi - lo) * std::uniform_real_distribution<double>(0.0, 1.0)(rng_); }
    // selectNode (:119-148): best-cost active node within the selection radius, else the nearest active node; distance(sample, node)
    int select_node(const Belief<D>& rs) {
        int best = -1;
        double bc = HUGE_VAL;
        auto dist = [&](int i) { return belief_distance((const double*)&rs, (const double*)&nodes_.at(i), D); };
        auto ok = [&](int i) { return !motions_[i].inactive && motions_[i].children >= 0; };
        nodes_.within(rs.v, prm_.selection_radius, dist, ok, [&](int i, double) { if (motions_[i].cost < bc) { bc = motions_[i].cost; best = i; } });
        if (best >= 0) return best;
        return nodes_.nearest(rs.v, dist, ok, nullptr);
    }
    // the constraint entries whose step falls on the new segment (absolute steps first .. first + seg.size() - 1)
    bool segment_ok(const std::vector<Belief<D>>& seg, int first) const {
        for (const auto& c : cons_) {
            for (size_t i = 0; i < c.steps.size(); ++i) {
                const int s = c.steps[i] - first;
                if (s < 0 || s >= (int)seg.size()) continue;
                std::vector<BeliefConstraint<D>> one(1);
                one[0].constrained = c.constrained; one[0].constraining = c.constraining;
                one[0].steps = {0}; one[0].times = {0.0}; one[0].states = {c.states[i]};
                if (!satisfies_constraints<D>(pvc_, std::vector<Belief<D>>{seg[s]}, 0.2, one)) return false;
            }
        }
        return true;
    }
    const Instance& inst_;
    int robot_;
    Pvc pvc_;
    Params prm_;
    std::mt19937_64 rng_;
    Svc svc_;
    double slow_[4], shigh_[4], clow_[4], chigh_[4];
    Belief<D> start_;
    double gx_ = 0, gy_ = 0, cost_ = 0;
    bool start_ok_ = true;
    std::vector<BeliefConstraint<D>> cons_;
    std::vector<Motion> motions_;
    std::vector<int> wit_rep_;
    MeanGrid<D> nodes_, wits_;
};

struct Result { bool solved = false; double seconds = 0, cost = 0; int cbs_nodes = 0, replans = 0; long long iterations = 0, tree_nodes = 0;
                std::vector<std::vector<std::vector<double>>> plan; };   // plan[robot][state][W]

template <int D, class Model>
Result kcbs(const Instance& inst, int svc_kind, int pvc_kind, Params prm, double seconds, double ll_seconds, Model model) {
    const int k = (int)inst.robots.size();
    std::vector<RobotShape> shapes;
    for (const auto& r : inst.robots) shapes.push_back(r.shape);
    const Pvc pvc = make_pvc(pvc_kind, shapes, pvc_kind == PVC_CHISQUARED ? 0.9 : inst.p_safe_agents);     // demos/main.cpp:105-138
    std::vector<std::unique_ptr<OracleBSST<D, Model>>> pl;
    for (int r = 0; r < k; ++r) { pl.emplace_back(new OracleBSST<D, Model>(inst, r, svc_kind, pvc, prm)); pl.back()->model = model; }
    typedef std::vector<std::vector<Belief<D>>> Plan;
    struct Node { Plan plan; double cost; std::vector<BeliefConstraint<D>> cons; };
    auto cmp = [](const std::shared_ptr<Node>& a, const std::shared_ptr<Node>& b) { return a->cost > b->cost; };
    std::priority_queue<std::shared_ptr<Node>, std::vector<std::shared_ptr<Node>>, decltype(cmp)> pq(cmp);
    const double t0 = now_s();
    Result res;
    auto root = std::make_shared<Node>();
    root->cost = 0;
    for (int r = 0; r < k; ++r) {                                                                 // KCBS.cpp:213-228
        pl[r]->set_constraints({});
        std::vector<Belief<D>> t;
        while (t.empty() && pl[r]->start_valid() && now_s() - t0 < seconds) t = pl[r]->solve(ll_seconds);
        if (t.empty()) { res.seconds = now_s() - t0; return res; }
        root->cost += 0.2 * (double)(t.size() - 1);
        root->plan.push_back(std::move(t));
    }
    pq.push(root);
    while (!pq.empty() && now_s() - t0 < seconds) {
        auto cur = pq.top(); pq.pop();
        res.cbs_nodes++;
        const std::vector<Conflict> confs = validate_plan<D>(pvc, cur->plan);                    // :289
        if (confs.empty()) {
            res.solved = true; res.cost = cur->cost;
            for (const auto& tr : cur->plan) {
                res.plan.emplace_back();
                for (const auto& b : tr) res.plan.back().emplace_back((const double*)&b, (const double*)&b + sizeof(b) / sizeof(double));
            }
            break;
        }
        const int agents[2] = {confs.front().a, confs.front().b};
        for (int i = 0; i < 2; ++i) {                                                             // :389-467
            const int a = agents[i];
            auto child = std::make_shared<Node>(*cur);
            child->cons.push_back(create_constraint<D>(cur->plan, confs, a, 0.2));
            std::vector<BeliefConstraint<D>> mine;
            for (const auto& c : child->cons) if (c.constrained == a) mine.push_back(c);
            pl[a]->set_constraints(mine);
            res.replans++;
            std::vector<Belief<D>> t = pl[a]->solve(ll_seconds);
            if (t.empty()) continue;
            child->cost += 0.2 * ((double)t.size() - (double)child->plan[a].size());
            child->plan[a] = std::move(t);
            pq.push(child);
        }
    }
    res.seconds = now_s() - t0;
    for (int r = 0; r < k; ++r) { res.iterations += pl[r]->stats.iterations; res.tree_nodes += pl[r]->stats.nodes; }
    return res;
}

std::string slurp(const std::string& p) { std::ifstream f(p); std::stringstream ss; ss << f.rdbuf(); return ss.str(); }

}  // namespace

int main(int argc, char** argv) {
    std::string map, scen, svc = "Blackmore", pvc = "Blackmore", output;
    int k = 1;
    double p_safe = 0.9, time_s = 300.0, lltime = 20.0;
    Params prm;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto next = [&]() -> std::string { if (i + 1 >= argc) { std::fprintf(stderr, "missing value for %s\n", a.c_str()); std::exit(2); } return argv[++i]; };
        if (a == "-m" || a == "--map") map = next();
        else if (a == "-s" || a == "--scen") scen = next();
        else if (a == "-k" || a == "--numAgents") k = std::atoi(next().c_str());
        else if (a == "--svc") svc = next();
        else if (a == "--pvc") pvc = next();
        else if (a == "-p" || a == "--p_safe") p_safe = std::atof(next().c_str());
        else if (a == "-t" || a == "--time") time_s = std::atof(next().c_str());
        else if (a == "--lltime") lltime = std::atof(next().c_str());
        else if (a == "--seed") prm.seed = std::strtoull(next().c_str(), nullptr, 10);
        else if (a == "--prune") prm.prune = true;
        else if (a == "-o" || a == "--output") output = next();
        else if (a == "--K" || a == "--devices") next();     // accepted for CLI compatibility with cck_plan, meaningless here (K = 1)
        else { std::fprintf(stderr, "unknown flag %s\n", a.c_str()); return 2; }
    }
    Instance inst;
    if (!load_map_text(slurp(map), inst) || !load_scen_text(slurp(scen), k, inst)) { std::fprintf(stderr, "cannot load %s / %s\n", map.c_str(), scen.c_str()); return 3; }
    split_risk(inst, p_safe);
    const int svc_kind = svc == "Blackmore" ? SVC_BLACKMORE : svc == "AdaptiveBlackmore" ? SVC_ADAPTIVE : SVC_CHISQUARED;
    const int pvc_kind = pvc == "Blackmore" ? PVC_BLACKMORE : pvc == "BoundingBox" ? PVC_BOUNDINGBOX : pvc == "ChiSquared" ? PVC_CHISQUARED
                         : pvc == "AdaptiveBlackmore" ? PVC_ADAPTIVE_BLACKMORE : PVC_ADAPTIVE_BOUNDINGBOX;
    Result r;
    if (inst.robots[0].model == "Uncertain-Unicycle-Model") r = kcbs<4, Model4>(inst, svc_kind, pvc_kind, prm, time_s, lltime, Model4{make_unicycle_params(0.2)});
    else r = kcbs<2, Model2>(inst, svc_kind, pvc_kind, prm, time_s, lltime, Model2{make_linear_params(0.2)});
    std::printf("{\"solved\": %s, \"seconds\": %.6f, \"cost\": %.6f, \"cbs_nodes\": %d, \"replans\": %d, \"launches\": 0, \"candidates\": %lld, "
                "\"tree_nodes\": %lld, \"obstacles\": %zu, \"agents\": %d, \"K\": 1, \"impl\": \"cpu-oracle\", \"prune\": %s}\n",
                r.solved ? "true" : "false", r.seconds, r.cost, r.cbs_nodes, r.replans, r.iterations, r.tree_nodes, inst.obstacles.size(), k, prm.prune ? "true" : "false");
    if (!output.empty() && r.solved) {     // same text format as host/cck_plan: "robot r T W" then T rows
        FILE* f = std::fopen(output.c_str(), "w");
        for (int i = 0; i < k; ++i) {
            std::fprintf(f, "robot %d %zu %zu\n", i, r.plan[i].size(), r.plan[i][0].size());
            for (const auto& st : r.plan[i]) { for (double v : st) std::fprintf(f, "%.17g ", v); std::fprintf(f, "\n"); }
        }
        std::fclose(f);
    }
    return r.solved ? 0 : 1;
}
