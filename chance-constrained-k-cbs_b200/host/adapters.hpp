// adapters.hpp — the reference's plugin classes for the hot path, re-implemented as thin hosts
// over the C ABI (include/cck_b200.h).  Same class names, constructor arguments and virtual
// signatures as the reference (cited per class) so they install with
// si->setStatePropagator / si->setStateValidityChecker / mrmp_pdef->setPlanValidator.
// They live in namespace cck_b200 so a binary may link them next to the reference's own.
//
// A single-item propagate()/isValid() is a batch-of-1 launch: correct, but slow by design —
// the fast path is the batched planner (planner.hpp) that calls the C ABI once per K candidates.
// There is no CPU fallback: without a usable GPU the GpuContext constructor throws.
#pragma once
#include <algorithm>
#include <cstring>
#include <fstream>
#include <limits>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/cck_b200.h"
#include "ompl_shim.hpp"

namespace cck_b200 {
namespace ob = ompl::base;
namespace oc = ompl::control;

// ---------------------------------------------------------------------------
// RAII over cck_ctx
// ---------------------------------------------------------------------------
class GpuContext {
public:
    explicit GpuContext(int device = 0) {
        const int rc = cck_ctx_create(device, &ctx_);
        if (rc != CCK_OK) throw std::runtime_error(std::string("cck_ctx_create: ") + cck_last_error(nullptr));
    }
    ~GpuContext() { cck_ctx_destroy(ctx_); }
    GpuContext(const GpuContext&) = delete;
    GpuContext& operator=(const GpuContext&) = delete;
    cck_ctx* get() const { return ctx_; }
    void check(int rc, const char* what) const {
        if (rc != CCK_OK) throw std::runtime_error(std::string(what) + ": " + cck_last_error(ctx_));
    }
private:
    cck_ctx* ctx_ = nullptr;
};
using GpuContextPtr = std::shared_ptr<GpuContext>;

// ---------------------------------------------------------------------------
// RealVectorBeliefSpace (include/Spaces/RealVectorBeliefSpace.h:10-121,
// src/Spaces/RealVectorBeliefSpace.cpp:7-76): values[] + sigma_ + lambda_, covariance = sum.
// Without Eigen the two matrices are small row-major dense matrices with (i,j) access.
// ---------------------------------------------------------------------------
struct Mat {
    int n = 0;
    std::vector<double> d;
    Mat() = default;
    explicit Mat(int dim, double diag = 0.0) : n(dim), d((size_t)dim * dim, 0.0) { for (int i = 0; i < dim; ++i) d[(size_t)i * dim + i] = diag; }
    double& operator()(int i, int j) { return d[(size_t)i * n + j]; }
    double operator()(int i, int j) const { return d[(size_t)i * n + j]; }
    int rows() const { return n; }
    Mat operator+(const Mat& o) const { Mat r(n); for (size_t i = 0; i < d.size(); ++i) r.d[i] = d[i] + o.d[i]; return r; }
};

class RealVectorBeliefSpace : public ob::RealVectorStateSpace {
public:
    explicit RealVectorBeliefSpace(unsigned dim) : ob::RealVectorStateSpace(dim) {}
    class StateType : public ob::RealVectorStateSpace::StateType {
    public:
        Mat getCovariance() const { return sigma_ + lambda_; }
        Mat sigma_, lambda_;
    };
    ob::State* allocState() const override {                      // RealVectorBeliefSpace.cpp:7-14
        auto* s = new StateType();
        s->values = new double[dimension_]();
        s->sigma_ = Mat((int)dimension_, 0.01);
        s->lambda_ = Mat((int)dimension_, 0.01);
        return s;
    }
    void copyState(ob::State* d, const ob::State* s) const override {   // :64-69
        ob::RealVectorStateSpace::copyState(d, s);
        d->as<StateType>()->sigma_ = s->as<StateType>()->sigma_;
        d->as<StateType>()->lambda_ = s->as<StateType>()->lambda_;
    }
    void freeState(ob::State* s) const override { delete[] s->as<StateType>()->values; delete s->as<StateType>(); }
};

// AoS state <-> one column of an SoA batch (W planes, leading dimension ld)
inline int belief_width(int dim) { return dim + 2 * dim * dim; }
inline void gather_state(const ob::State* s, int dim, double* soa, int64_t ld, int64_t i) {
    const auto* b = s->as<RealVectorBeliefSpace::StateType>();
    for (int f = 0; f < dim; ++f) soa[f * ld + i] = b->values[f];
    for (int f = 0; f < dim * dim; ++f) { soa[(dim + f) * ld + i] = b->sigma_.d[f]; soa[(dim + dim * dim + f) * ld + i] = b->lambda_.d[f]; }
}
inline void scatter_state(const double* soa, int64_t ld, int64_t i, int dim, ob::State* s) {
    auto* b = s->as<RealVectorBeliefSpace::StateType>();
    if (b->sigma_.n != dim) { b->sigma_ = Mat(dim); b->lambda_ = Mat(dim); }
    for (int f = 0; f < dim; ++f) b->values[f] = soa[f * ld + i];
    for (int f = 0; f < dim * dim; ++f) { b->sigma_.d[f] = soa[(dim + f) * ld + i]; b->lambda_.d[f] = soa[(dim + dim * dim + f) * ld + i]; }
}

// ---------------------------------------------------------------------------
// World model (src/utils/Instance.cpp, src/utils/common.cpp)
// ---------------------------------------------------------------------------
struct Location { double x_, y_; };
struct Obstacle { double x_, y_; };                 // RectangularObstacle(r, c, 1, 1)  (Instance.cpp:108)
struct Robot {
    std::string name_, dyn_model_;
    Location start_, goal_;
    double bounding_rad_;
    const std::string& getName() const { return name_; }
    const std::string& getDynamicsModel() const { return dyn_model_; }
    double getBoundingRadius() const { return bounding_rad_; }
    Location getStartLocation() const { return start_; }
    Location getGoalLocation() const { return goal_; }
};

class Instance {
public:
    // Instance::Instance + load_map_ + load_agents_ (Instance.cpp:4-47, 78-174)
    Instance(const std::string& map_path, const std::string& scen_path, int num_agents, double p_safe, const std::string& svc = "Blackmore",
             const std::string& pvc = "Blackmore")
        : num_agents_(num_agents), p_safe_(p_safe), svc_(svc), pvc_(pvc) {
        if (!load_map_(map_path)) throw std::runtime_error("Instance: unable to load map " + map_path);
        if (!load_agents_(scen_path)) throw std::runtime_error("Instance: unable to load scen file " + scen_path);
        cck_split_risk(p_safe_, (int)obstacles_.size(), (int)robots_.size(), &p_safe_obs_, &p_safe_agnts_);   // :38-44
    }
    double getPsafe() const { return p_safe_; }
    double getPsafeObs() const { return p_safe_obs_; }
    double getPsafeAgents() const { return p_safe_agnts_; }
    const std::string& getSVC() const { return svc_; }
    const std::string& getPVC() const { return pvc_; }
    std::vector<double> getDimensions() const { return {x_max_, y_max_}; }
    const std::vector<Obstacle>& getObstacles() const { return obstacles_; }
    const std::vector<Robot>& getRobots() const { return robots_; }
    bool mapOverrun() const { return map_overrun_; }
private:
    bool load_map_(const std::string& path) {
        std::ifstream f(path);
        if (!f.is_open()) return false;
        std::string line;
        std::getline(f, line);
        auto second = [](const std::string& l) { std::istringstream s(l); std::string a, b; s >> a >> b; return std::atoi(b.c_str()); };
        if (!line.empty() && line[0] == 't') {
            std::getline(f, line); y_max_ = second(line);
            std::getline(f, line); x_max_ = second(line);
            std::getline(f, line);
        }
        if (!(x_max_ > 0 && y_max_ > 0)) return false;
        int c = 0;
        while (std::getline(f, line)) {
            for (int r = 0; r < x_max_; ++r) {
                // quirk 13: line[r] with r == size() reads '\0' (an obstacle); r > size() is UB in the reference
                char ch;
                if ((size_t)r < line.size()) ch = line[r];
                else if ((size_t)r == line.size()) ch = '\0';
                else { map_overrun_ = true; continue; }
                if (ch != '.') obstacles_.push_back(Obstacle{(double)r, (double)c});
            }
            c++;
        }
        return true;
    }
    bool load_agents_(const std::string& path) {
        std::ifstream f(path);
        if (!f.is_open()) return false;
        std::string line;
        std::getline(f, line);
        for (int i = 0; i < num_agents_; ++i) {
            if (!std::getline(f, line) || line.empty()) return false;
            std::vector<std::string> tok;
            std::string cur;
            for (char ch : line) { if (ch == '\t') { if (!cur.empty()) tok.push_back(cur); cur.clear(); } else if (ch != '\r') cur.push_back(ch); }
            if (!cur.empty()) tok.push_back(cur);
            if (tok.size() < 10) return false;
            if (tok[8] != "Rectangle") return false;   // belief checkers need createBoundingShape (quirk 11)
            Robot r;
            r.start_ = Location{(double)std::atoi(tok[4].c_str()), (double)std::atoi(tok[5].c_str())};
            r.goal_ = Location{(double)std::atoi(tok[6].c_str()), (double)std::atoi(tok[7].c_str())};
            r.dyn_model_ = tok[9];
            r.name_ = "Robot " + std::to_string(i);
            r.bounding_rad_ = cck_rect_robot_bounding_radius(r.start_.x_, r.start_.y_, 0.25, 0.25);   // :166
            robots_.push_back(r);
        }
        return true;
    }
    double x_max_ = -1, y_max_ = -1;
    int num_agents_;
    double p_safe_, p_safe_obs_ = -1, p_safe_agnts_ = -1;
    std::string svc_, pvc_;
    std::vector<Obstacle> obstacles_;
    std::vector<Robot> robots_;
    bool map_overrun_ = false;
};
using InstancePtr = std::shared_ptr<Instance>;

// ---------------------------------------------------------------------------
// State propagators
// ---------------------------------------------------------------------------
inline void install_model(const GpuContextPtr& g, int model, double dt) {
    cck_model_params p;
    g->check(cck_model_params_default(model, dt, &p), "cck_model_params_default");
    g->check(cck_set_model(g->get(), &p), "cck_set_model");
}

class BatchOfOnePropagator : public oc::StatePropagator {
public:
    BatchOfOnePropagator(const oc::SpaceInformationPtr& si, GpuContextPtr gpu, int model, int dim)
        : oc::StatePropagator(si.get()), gpu_(std::move(gpu)), dim_(dim) { install_model(gpu_, model, si->getPropagationStepSize()); }
    void propagate(const ob::State* state, const oc::Control* control, const double duration, ob::State* result) const override {
        const int W = belief_width(dim_);
        double in[36], out[36], u[4];
        gather_state(state, dim_, in, 1, 0);
        for (int f = 0; f < dim_; ++f) u[f] = control->as<oc::RealVectorControlSpace::ControlType>()->values[f];
        gpu_->check(cck_propagate(gpu_->get(), 1, in, 1, u, 1, duration, out, 1), "cck_propagate");
        (void)W;
        scatter_state(out, 1, 0, dim_, result);
    }
    bool canPropagateBackward() const override { return false; }
protected:
    GpuContextPtr gpu_;
    int dim_;
};

// R2_UncertainLinearStatePropagator(si)  (include/StatePropogators/UncertainLinearSP.h:14-50)
class R2_UncertainLinearStatePropagator : public BatchOfOnePropagator {
public:
    R2_UncertainLinearStatePropagator(const oc::SpaceInformationPtr& si, GpuContextPtr gpu)
        : BatchOfOnePropagator(si, std::move(gpu), CCK_MODEL_LINEAR2D, 2) {}
};
// DynUnicycleControlSpace(si)  (include/StatePropogators/UncertainUnicycleSP.h:56-127)
class DynUnicycleControlSpace : public BatchOfOnePropagator {
public:
    DynUnicycleControlSpace(const oc::SpaceInformationPtr& si, GpuContextPtr gpu)
        : BatchOfOnePropagator(si, std::move(gpu), CCK_MODEL_UNICYCLE, 4) {}
};

// ---------------------------------------------------------------------------
// State validity checkers
// ---------------------------------------------------------------------------
struct ObstacleTables { std::vector<double> A, B, centres, rects; int M = 0; };
inline ObstacleTables build_tables(const Instance& inst, const Robot& r) {
    ObstacleTables t;
    t.M = (int)inst.getObstacles().size();
    t.A.resize((size_t)t.M * 8); t.B.resize((size_t)t.M * 4); t.centres.resize((size_t)t.M * 2); t.rects.resize((size_t)t.M * 4);
    for (int o = 0; o < t.M; ++o) { t.centres[2 * o] = inst.getObstacles()[o].x_; t.centres[2 * o + 1] = inst.getObstacles()[o].y_; }
    if (t.M) cck_build_obstacle_tables(t.M, t.centres.data(), 1.0, 1.0, r.getBoundingRadius(), t.A.data(), t.B.data(), t.rects.data());
    return t;
}

class GpuSVC : public ob::StateValidityChecker {
public:
    GpuSVC(const oc::SpaceInformationPtr& si, GpuContextPtr gpu, InstancePtr inst, const Robot* r, double accep_prob, int kind)
        : ob::StateValidityChecker(si.get()), gpu_(std::move(gpu)) {
        dim_ = (int)si->getStateSpace()->getDimension();
        const auto& b = si->getStateSpace()->as<ob::RealVectorStateSpace>()->getBounds();
        cck_svc_params p{};
        p.kind = kind; p.dim = dim_;
        for (int i = 0; i < dim_; ++i) { p.low[i] = b.low[i]; p.high[i] = b.high[i]; }
        const ObstacleTables t = build_tables(*inst, *r);
        p.p_collision = 1 - accep_prob;                                                    // PCCBlackmoreSVC.cpp:5
        if (t.M > 0) p.erf_inv_result = cck_erf_inv(1 - 2 * (p.p_collision / t.M));        // :10
        p.sc = cck_chi2_quantile_2dof(accep_prob);                                         // ChiSquaredBoundarySVC.cpp:7
        p.max_rad = r->getBoundingRadius();
        gpu_->check(cck_set_obstacles(gpu_->get(), &p, t.A.data(), t.B.data(), t.centres.data(), t.rects.data(), t.M), "cck_set_obstacles");
    }
    bool isValid(const ob::State* state) const override {
        double in[36];
        uint8_t v = 0;
        gather_state(state, dim_, in, 1, 0);
        gpu_->check(cck_is_valid(gpu_->get(), 1, in, 1, &v), "cck_is_valid");
        return v != 0;
    }
protected:
    GpuContextPtr gpu_;
    int dim_;
};
// PCCBlackmoreSVC(si, instance, robot, accep_prob)  (include/StateValidityCheckers/PCCBlackmoreSVC.h:15-43)
class PCCBlackmoreSVC : public GpuSVC {
public:
    PCCBlackmoreSVC(const oc::SpaceInformationPtr& si, GpuContextPtr g, InstancePtr i, const Robot* r, double p) : GpuSVC(si, std::move(g), std::move(i), r, p, CCK_SVC_BLACKMORE) {}
};
// AdaptiveRiskBlackmoreSVC  (include/StateValidityCheckers/AdaptiveRiskBlackmoreSVC.h:14-39)
class AdaptiveRiskBlackmoreSVC : public GpuSVC {
public:
    AdaptiveRiskBlackmoreSVC(const oc::SpaceInformationPtr& si, GpuContextPtr g, InstancePtr i, const Robot* r, double p) : GpuSVC(si, std::move(g), std::move(i), r, p, CCK_SVC_ADAPTIVE) {}
};
// ChiSquaredBoundarySVC  (include/StateValidityCheckers/ChiSquaredBoundarySVC.h:16-57)
class ChiSquaredBoundarySVC : public GpuSVC {
public:
    ChiSquaredBoundarySVC(const oc::SpaceInformationPtr& si, GpuContextPtr g, InstancePtr i, const Robot* r, double p) : GpuSVC(si, std::move(g), std::move(i), r, p, CCK_SVC_CHISQUARED) {}
};

// ---------------------------------------------------------------------------
// Centralized variants: the composite state of k agents is a RealVectorBeliefSpace(2k) state whose
// sigma_ / lambda_ carry one 2x2 block per agent on the diagonal.
// ---------------------------------------------------------------------------
// CentralizedUncertainLinearStatePropagator(si)  (include/StatePropogators/CentralizedUncertainLinearSP.h:17-70,
// src/StatePropogators/CentralizedUncertainLinearSP.cpp:42-102): singleAgentPropogate_ is the single-agent linear step with
// the same constants (F = H = I, Q = R = 0.01 I, A_cl = (1 - 0.3) I, mean += duration_ * u), applied to every agent's block;
// entries of the result outside the diagonal blocks are left as they are.  ONE cck_propagate over the k agents.
class CentralizedUncertainLinearStatePropagator : public oc::StatePropagator {
public:
    CentralizedUncertainLinearStatePropagator(const oc::SpaceInformationPtr& si, GpuContextPtr gpu)
        : oc::StatePropagator(si.get()), gpu_(std::move(gpu)), k_((int)si->getStateSpace()->getDimension() / 2) {
        install_model(gpu_, CCK_MODEL_LINEAR2D, si->getPropagationStepSize());
    }
    void propagate(const ob::State* state, const oc::Control* control, const double duration, ob::State* result) const override {
        const auto* s = state->as<RealVectorBeliefSpace::StateType>();
        auto* r = result->as<RealVectorBeliefSpace::StateType>();
        const double* u = control->as<oc::RealVectorControlSpace::ControlType>()->values;
        std::vector<double> in((size_t)10 * k_), ctl((size_t)2 * k_), out((size_t)10 * k_);
        for (int a = 0; a < k_; ++a) {
            in[0 * k_ + a] = s->values[2 * a]; in[1 * k_ + a] = s->values[2 * a + 1];
            for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) {
                in[(2 + i * 2 + j) * k_ + a] = s->sigma_(2 * a + i, 2 * a + j);
                in[(6 + i * 2 + j) * k_ + a] = s->lambda_(2 * a + i, 2 * a + j);
            }
            ctl[a] = u[2 * a]; ctl[k_ + a] = u[2 * a + 1];
        }
        gpu_->check(cck_propagate(gpu_->get(), k_, in.data(), k_, ctl.data(), k_, duration, out.data(), k_), "cck_propagate (centralized)");
        if (r->sigma_.n != 2 * k_) { r->sigma_ = Mat(2 * k_); r->lambda_ = Mat(2 * k_); }
        for (int a = 0; a < k_; ++a) {
            r->values[2 * a] = out[0 * k_ + a]; r->values[2 * a + 1] = out[1 * k_ + a];
            for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) {
                r->sigma_(2 * a + i, 2 * a + j) = out[(2 + i * 2 + j) * k_ + a];
                r->lambda_(2 * a + i, 2 * a + j) = out[(6 + i * 2 + j) * k_ + a];
            }
        }
    }
    bool canPropagateBackward() const override { return false; }
private:
    GpuContextPtr gpu_;
    int k_;
};

// CentralizedChiSquaredBoundarySVC(si, instance, accep_prob)  (include/StateValidityCheckers/CentralizedChiSquaredBoundarySVC.h,
// src/StateValidityCheckers/CentralizedChiSquaredBoundarySVC.cpp:24-99): bounds of the composite state; per agent the decagon
// of radius lambda_max(SIGMA block only - not Sigma + Lambda) * sc + r_a against every inflated obstacle (ONE cck_is_valid
// over the k agents, with Lambda = 0 and the block symmetrised from its lower triangle as SelfAdjointEigenSolver reads it);
// then every agent pair's decagons must be disjoint (ONE cck_chisq_pairs_disjoint; see the note on the reference's
// uninitialised solver_a2 in the oracle).
class CentralizedChiSquaredBoundarySVC : public ob::StateValidityChecker {
public:
    CentralizedChiSquaredBoundarySVC(const oc::SpaceInformationPtr& si, GpuContextPtr gpu, InstancePtr inst, double accep_prob)
        : ob::StateValidityChecker(si.get()), gpu_(std::move(gpu)), inst_(std::move(inst)), k_((int)inst_->getRobots().size()) {
        sc_ = cck_chi2_quantile_2dof(accep_prob);                                          // :8
        const auto& b = si->getStateSpace()->as<ob::RealVectorStateSpace>()->getBounds();
        low_ = b.low; high_ = b.high;
        install_model(gpu_, CCK_MODEL_LINEAR2D, si->getPropagationStepSize());
        cck_svc_params p{};
        p.kind = CCK_SVC_CHISQUARED; p.dim = 2;
        // the composite bounds are checked here (satisfiesBounds on all 2k values); the kernel gets agent 0's (x, y) bounds,
        // which set_up_CentralizedBSST_Problem repeats for every agent (OmplSetUp.cpp:367-376)
        for (int i = 0; i < 2; ++i) { p.low[i] = low_[i]; p.high[i] = high_[i]; }
        const ObstacleTables t = build_tables(*inst_, inst_->getRobots()[0]);              // all robots share one bounding radius in the scenario files
        p.p_collision = 1 - accep_prob; p.sc = sc_; p.max_rad = inst_->getRobots()[0].getBoundingRadius();
        gpu_->check(cck_set_obstacles(gpu_->get(), &p, t.A.data(), t.B.data(), t.centres.data(), t.rects.data(), t.M), "cck_set_obstacles");
    }
    bool isValid(const ob::State* state) const override {
        const auto* s = state->as<RealVectorBeliefSpace::StateType>();
        for (int d = 0; d < 2 * k_; ++d)                                                    // RealVectorStateSpace::satisfiesBounds
            if (s->values[d] - std::numeric_limits<double>::epsilon() > high_[d] || s->values[d] + std::numeric_limits<double>::epsilon() < low_[d]) return false;
        std::vector<double> in((size_t)10 * k_, 0.0);
        for (int a = 0; a < k_; ++a) {
            in[0 * k_ + a] = s->values[2 * a]; in[1 * k_ + a] = s->values[2 * a + 1];
            const double s00 = s->sigma_(2 * a, 2 * a), s10 = s->sigma_(2 * a + 1, 2 * a), s11 = s->sigma_(2 * a + 1, 2 * a + 1);
            in[2 * k_ + a] = s00; in[3 * k_ + a] = s10; in[4 * k_ + a] = s10; in[5 * k_ + a] = s11;    // lower triangle, Lambda planes stay 0
        }
        std::vector<uint8_t> ok(k_);
        gpu_->check(cck_is_valid(gpu_->get(), k_, in.data(), k_, ok.data()), "cck_is_valid (centralized)");
        for (int a = 0; a < k_; ++a) if (!ok[a]) return false;
        const int np = k_ * (k_ - 1) / 2;
        if (np == 0) return true;
        std::vector<double> mua((size_t)2 * np), cova((size_t)4 * np), rada(np), mub((size_t)2 * np), covb((size_t)4 * np), radb(np);
        int q = 0;
        for (int a = 0; a < k_; ++a)
            for (int b = a + 1; b < k_; ++b, ++q) {
                mua[q] = in[a]; mua[np + q] = in[k_ + a]; mub[q] = in[b]; mub[np + q] = in[k_ + b];
                for (int f = 0; f < 4; ++f) { cova[(size_t)f * np + q] = in[(2 + f) * k_ + a]; covb[(size_t)f * np + q] = in[(2 + f) * k_ + b]; }
                rada[q] = inst_->getRobots()[a].getBoundingRadius(); radb[q] = inst_->getRobots()[b].getBoundingRadius();
            }
        std::vector<uint8_t> dj(np);
        gpu_->check(cck_chisq_pairs_disjoint(gpu_->get(), np, mua.data(), cova.data(), rada.data(), mub.data(), covb.data(), radb.data(), sc_, dj.data()),
                    "cck_chisq_pairs_disjoint");
        for (int i = 0; i < np; ++i) if (!dj[i]) return false;
        return true;
    }
private:
    GpuContextPtr gpu_;
    InstancePtr inst_;
    int k_;
    double sc_ = 0;
    std::vector<double> low_, high_;
};

// ---------------------------------------------------------------------------
// Conflicts, constraints, plan validity checkers
// ---------------------------------------------------------------------------
struct Conflict {                                   // include/utils/Conflict.h:6-13
    int agent1Idx_, agent2Idx_, timeStep_;
};
using ConflictPtr = std::shared_ptr<Conflict>;

struct BeliefState { int dim = 2; std::vector<double> v; };   // flat [W] record: values | Sigma | Lambda
using Trajectory = std::vector<BeliefState>;                  // one state per dt (an interpolated PathControl)
using Plan = std::vector<Trajectory>;

// Constraint / BeliefConstraint (include/Constraints/Constraint.h:10-45, BeliefConstraint.h:8-16)
class BeliefConstraint {
public:
    BeliefConstraint(int constrained, int constraining, std::vector<double> times, std::vector<int> steps, std::vector<BeliefState> states)
        : times_(std::move(times)), steps_(std::move(steps)), belief_states_(std::move(states)), constrained_agent_(constrained), constraining_agent_(constraining) {}
    const std::vector<double>& getTimes() const { return times_; }
    const std::vector<int>& getSteps() const { return steps_; }
    int getConstrainedAgent() const { return constrained_agent_; }
    int getConstrainingAgent() const { return constraining_agent_; }
    const std::vector<BeliefState>& getStates() const { return belief_states_; }
private:
    std::vector<double> times_;
    std::vector<int> steps_;
    std::vector<BeliefState> belief_states_;
    int constrained_agent_, constraining_agent_;
};
using ConstraintPtr = std::shared_ptr<BeliefConstraint>;

// PlanValidityChecker (include/PlanValidityCheckers/PlanValidityChecker.h:12-29) over the C ABI.
// kind selects MinkowskiSumBlackmorePVC / BoundingBoxBlackmorePVC / ChiSquaredBoundaryPVC /
// AdaptiveRiskBlackmorePVC / AdaptiveRiskBoundingBoxPVC.
class BeliefPVC {
public:
    BeliefPVC(GpuContextPtr gpu, InstancePtr inst, int kind, int dim, double p_safe_arg, double step_size, std::string name, int grid_n = 2)
        : gpu_(std::move(gpu)), inst_(std::move(inst)), kind_(kind), dim_(dim), step_(step_size), name_(std::move(name)) {
        k_ = (int)inst_->getRobots().size();
        std::vector<double> radii(k_), A((size_t)k_ * k_ * 8, 0.0), B((size_t)k_ * k_ * 4, 0.0);
        for (int i = 0; i < k_; ++i) radii[i] = inst_->getRobots()[i].getBoundingRadius();
        const bool bb = kind == CCK_PVC_BOUNDINGBOX || kind == CCK_PVC_ADAPTIVE_BOUNDINGBOX;
        for (int a = 0; a < k_; ++a)
            for (int b = a + 1; b < k_; ++b) cck_build_pair_halfplanes(bb ? 1 : 0, radii[a], radii[b], &A[((size_t)a * k_ + b) * 8], &B[((size_t)a * k_ + b) * 4]);
        std::vector<int32_t> order(k_);
        for (int i = 0; i < k_; ++i) order[i] = i;
        std::sort(order.begin(), order.end(), [&](int x, int y) { return inst_->getRobots()[x].getName() < inst_->getRobots()[y].getName(); });   // std::map order (quirk 6)
        cck_pvc_params p{};
        p.kind = kind; p.k = k_; p.dim = dim; p.grid_n = grid_n;
        p.p_coll_agnts = 1 - p_safe_arg;                                                 // BeliefPVC.cpp:7
        const double p_coll_dist = k_ > 1 ? p.p_coll_agnts / (k_ - 1) : p.p_coll_agnts;  // MinkowskiSumBlackmorePVC.cpp:7-8
        p.c_pair = cck_erf_inv(1 - (2 * p_coll_dist));
        p.sc = cck_chi2_quantile_2dof(p_safe_arg);
        p.radii = radii.data(); p.pair_A = A.data(); p.pair_B = B.data(); p.map_order = order.data();
        gpu_->check(cck_set_pvc(gpu_->get(), &p), "cck_set_pvc");
    }
    const std::string& getName() const { return name_; }

    // validatePlan(Plan): the run of conflicts of the earliest conflicting pair (§3.4)
    std::vector<ConflictPtr> validatePlan(const Plan& p) const {
        const int W = belief_width(dim_);
        std::vector<int32_t> lens(k_);
        int T = 0;
        for (int r = 0; r < k_; ++r) { lens[r] = (int)p[r].size(); T = std::max(T, lens[r]); }
        std::vector<double> buf((size_t)k_ * W * T, 0.0);
        for (int r = 0; r < k_; ++r)
            for (int t = 0; t < lens[r]; ++t)
                for (int f = 0; f < W; ++f) buf[((size_t)r * W + f) * T + t] = p[r][t].v[f];
        cck_conflict_run run{};
        gpu_->check(cck_validate_plan(gpu_->get(), k_, lens.data(), buf.data(), T, &run), "cck_validate_plan");
        std::vector<ConflictPtr> out;
        for (int i = 0; i < run.run_len && run.first_step >= 0; ++i) out.push_back(std::make_shared<Conflict>(Conflict{run.a, run.b, run.first_step + i}));
        return out;
    }
    // createConstraint(Plan, conflicts, constrained_robot)  (BeliefPVC.cpp:10-40)
    ConstraintPtr createConstraint(const Plan& p, const std::vector<ConflictPtr>& conflicts, int constrained_robot) const {
        const int constraining = (constrained_robot == conflicts.front()->agent1Idx_) ? conflicts.front()->agent2Idx_ : conflicts.front()->agent1Idx_;
        std::vector<double> times; std::vector<int> steps; std::vector<BeliefState> states;
        for (const auto& c : conflicts) {
            times.push_back(c->timeStep_ * step_);
            steps.push_back(c->timeStep_);
            const Trajectory& t = p[constraining];
            states.push_back(c->timeStep_ < (int)t.size() ? t[c->timeStep_] : t.back());     // waits at its last state (quirk 8)
        }
        return std::make_shared<BeliefConstraint>(constrained_robot, constraining, times, steps, states);
    }
    // upload the constraints of one agent; they stay installed for the batched expansion kernel
    void installConstraints(const std::vector<ConstraintPtr>& constraints) const {
        std::vector<int32_t> ent_step, pair_ab;
        std::vector<double> mu, cov;
        for (const auto& c : constraints) {
            const int a = c->getConstrainedAgent(), b = c->getConstrainingAgent();   // planning robot first (cck_b200.h)
            for (size_t i = 0; i < c->getSteps().size(); ++i) {
                const std::vector<double>& s = c->getStates()[i].v;
                ent_step.push_back(c->getSteps()[i]);
                pair_ab.push_back(a * k_ + b);
                mu.push_back(s[0]); mu.push_back(s[1]);
                const double* S = &s[dim_]; const double* L = &s[dim_ + dim_ * dim_];
                if (dim_ == 4) { cov.push_back(S[0] + L[0]); cov.push_back(S[2] + L[2]); cov.push_back(S[8] + L[8]); cov.push_back(S[10] + L[10]); }   // BeliefPVC.cpp:94-97
                else for (int f = 0; f < 4; ++f) cov.push_back(S[f] + L[f]);
            }
        }
        gpu_->check(cck_set_constraints(gpu_->get(), (int)ent_step.size(), ent_step.data(), pair_ab.data(), mu.data(), cov.data()), "cck_set_constraints");
    }
    // satisfiesConstraints(path, constraints)  (MinkowskiSumBlackmorePVC.cpp:64-104)
    bool satisfiesConstraints(const Trajectory& path, const std::vector<ConstraintPtr>& constraints) const {
        installConstraints(constraints);
        return satisfiesConstraintsInstalled(path);
    }
    // the same check against the constraints already on the device (installConstraints)
    bool satisfiesConstraintsInstalled(const Trajectory& path) const {
        const int W = belief_width(dim_), T = (int)path.size();
        std::vector<double> buf((size_t)T * W);
        for (int t = 0; t < T; ++t) for (int f = 0; f < W; ++f) buf[(size_t)t * W + f] = path[t].v[f];   // ld = 1 path
        int32_t len = T, s0 = 0;
        uint8_t ok = 0;
        gpu_->check(cck_satisfies_constraints(gpu_->get(), 1, buf.data(), 1, T, &len, &s0, &ok), "cck_satisfies_constraints");
        return ok != 0;
    }
protected:
    GpuContextPtr gpu_;
    InstancePtr inst_;
    int kind_, dim_, k_ = 0;
    double step_;
    std::string name_;
};
using PlanValidityCheckerPtr = std::shared_ptr<BeliefPVC>;

// the reference's five belief PVCs (demos/main.cpp:105-138 wiring: Blackmore-type get p_safe_agents,
// ChiSquared gets the hard-coded 0.9)
inline PlanValidityCheckerPtr make_pvc(const std::string& name, GpuContextPtr g, InstancePtr inst, int dim, double step) {
    const double pa = inst->getPsafeAgents();
    if (name == "Blackmore") return std::make_shared<BeliefPVC>(g, inst, CCK_PVC_BLACKMORE, dim, pa, step, "MinkowskiSumBlackmorePVC");
    if (name == "BoundingBox") return std::make_shared<BeliefPVC>(g, inst, CCK_PVC_BOUNDINGBOX, dim, pa, step, "BoundingBoxBlackmorePVC");
    if (name == "ChiSquared") return std::make_shared<BeliefPVC>(g, inst, CCK_PVC_CHISQUARED, dim, 0.9, step, "ChiSquaredBoundaryPVC");
    if (name == "AdaptiveBlackmore") return std::make_shared<BeliefPVC>(g, inst, CCK_PVC_ADAPTIVE_BLACKMORE, dim, pa, step, "AdaptiveRiskBlackmorePVC");
    if (name == "AdaptiveBoundingBox") return std::make_shared<BeliefPVC>(g, inst, CCK_PVC_ADAPTIVE_BOUNDINGBOX, dim, pa, step, "AdaptiveRiskBoundingBoxPVC");
    if (name == "Blackmore2") return std::make_shared<BeliefPVC>(g, inst, CCK_PVC_BLACKMORE2, dim, pa, step, "Blackmore2PVC");     // not reachable from the reference CLI
    if (name.rfind("CDFGrid", 0) == 0) {                                  // `--pvc CDFGrid-n` (demos/main.cpp:126-134: token after '-')
        const size_t dash = name.find('-');
        const int disks = dash == std::string::npos ? 0 : std::atoi(name.c_str() + dash + 1);
        return std::make_shared<BeliefPVC>(g, inst, CCK_PVC_CDFGRID, dim, pa, step, "CDFGridPVC(" + std::to_string(disks) + ")", disks);
    }
    throw std::runtime_error("unknown --pvc " + name);   // the reference asserts (demos/main.cpp:40,109-137)
}

}  // namespace cck_b200
