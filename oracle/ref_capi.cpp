// ref_capi.cpp — TEST INFRASTRUCTURE: C entry points over the REFERENCE'S OWN SOURCES for the default-configuration path,
// compiled from where they lie under /root/reference (never copied):
//   src/utils/common.cpp, src/utils/Instance.cpp                 (map/scen parser, obstacles, robots, risk split)
//   src/StatePropogators/UncertainLinearSP.cpp                    (R2_UncertainLinearStatePropagator::propagate)
//   src/StateValidityCheckers/PCCBlackmoreSVC.cpp                 (Minkowski sums, half-planes, erf_inv constant, isValid)
// against the stand-in headers of oracle/shim/ (Eigen / Boost / OMPL subsets, see the notes there: every matrix
// product on this path has inner dimension 2, so the shim's evaluation order cannot differ from Eigen's).
// src/Spaces/RealVectorBeliefSpace.cpp is compiled in place too (round 2; the entry points over it, the unicycle
// propagator, the goal and the chi-squared checkers are in ref_capi2.cpp).
// propagateWhileValid is OMPL's (SpaceInformation.cpp), restated here around the reference's two virtuals.
#include <algorithm>
#include <any>
#include <cassert>
#include <cmath>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "shim_eigen.hpp"
#include "shim_boost.hpp"
#include "shim_ompl.hpp"

#define private public      // test glue only: read A_list_, B_list_, erf_inv_result_ of the reference's checker
#include "StateValidityCheckers/PCCBlackmoreSVC.h"
#include "StateValidityCheckers/AdaptiveRiskBlackmoreSVC.h"
#undef private
#include "StatePropogators/UncertainLinearSP.h"

namespace {
struct RefWorld { InstancePtr inst; };
struct RefSvc {
    std::shared_ptr<RealVectorBeliefSpace> space;
    oc::SpaceInformationPtr si;
    std::shared_ptr<PCCBlackmoreSVC> pcc;                 // kind 0
    std::shared_ptr<AdaptiveRiskBlackmoreSVC> adaptive;   // kind 1
    ob::StateValidityChecker* svc = nullptr;
    std::shared_ptr<R2_UncertainLinearStatePropagator> prop;
};
struct Quiet {   // Instance's constructor prints every obstacle
    std::ostringstream sink;     // declared (hence constructed) before `old` uses it
    std::streambuf* old;
    Quiet() : old(std::cout.rdbuf(sink.rdbuf())) {}
    ~Quiet() { std::cout.rdbuf(old); }
};
void load(RealVectorBeliefSpace::StateType* s, const double* rec, int dim) {
    for (int i = 0; i < dim; ++i) s->values[i] = rec[i];
    for (int i = 0; i < dim; ++i) for (int j = 0; j < dim; ++j) { s->sigma_(i, j) = rec[dim + i * dim + j]; s->lambda_(i, j) = rec[dim + dim * dim + i * dim + j]; }
}
void store(const RealVectorBeliefSpace::StateType* s, double* rec, int dim) {
    for (int i = 0; i < dim; ++i) rec[i] = s->values[i];
    for (int i = 0; i < dim; ++i) for (int j = 0; j < dim; ++j) { rec[dim + i * dim + j] = s->sigma_(i, j); rec[dim + dim * dim + i * dim + j] = s->lambda_(i, j); }
}
}  // namespace

extern "C" {

void* ref_world_create(const char* map, const char* scen, int k, double p_safe) {
    po::variables_map vm;
    vm.set<std::string>("solver", "K-CBS"); vm.set<std::string>("lowlevel", "BSST");
    vm.set<int>("numAgents", k); vm.set<std::string>("map", map); vm.set<std::string>("scen", scen);
    vm.set<double>("p_safe", p_safe); vm.set<std::string>("pvc", "Blackmore"); vm.set<std::string>("svc", "Blackmore");
    Quiet q;
    RefWorld* w = new RefWorld();
    w->inst = std::make_shared<Instance>(vm);
    return w;
}
void ref_world_free(void* w) { delete (RefWorld*)w; }
// out: n_obstacles, n_robots, x_max, y_max, p_safe_obs, p_safe_agents
void ref_world_info(void* w, double* out) {
    const Instance& I = *((RefWorld*)w)->inst;
    out[0] = (double)I.getObstacles().size(); out[1] = (double)I.getRobots().size();
    out[2] = I.getDimensions()[0]; out[3] = I.getDimensions()[1]; out[4] = I.getPsafeObs(); out[5] = I.getPsafeAgents();
}
void ref_world_obstacles(void* w, double* centres) {
    const auto obs = ((RefWorld*)w)->inst->getObstacles();
    for (size_t o = 0; o < obs.size(); ++o) { centres[2 * o] = obs[o]->x_; centres[2 * o + 1] = obs[o]->y_; }
}
// out per robot: start x, y, goal x, y, bounding radius
void ref_world_robots(void* w, double* out) {
    const auto rs = ((RefWorld*)w)->inst->getRobots();
    for (size_t r = 0; r < rs.size(); ++r) {
        out[5 * r] = rs[r]->getStartLocation().x_; out[5 * r + 1] = rs[r]->getStartLocation().y_;
        out[5 * r + 2] = rs[r]->getGoalLocation().x_; out[5 * r + 3] = rs[r]->getGoalLocation().y_;
        out[5 * r + 4] = rs[r]->getBoundingRadius();
    }
}

// the wiring of set_up_ConstraintBSST_MP_Problems for the linear model (src/utils/OmplSetUp.cpp:183-228)
void* ref_svc_create_kind(void* w, int kind, int robot, const double* low, const double* high, double step);
void* ref_svc_create(void* w, int robot, const double* low, const double* high, double step) { return ref_svc_create_kind(w, 0, robot, low, high, step); }
// kind 0: PCCBlackmoreSVC, kind 1: AdaptiveRiskBlackmoreSVC (OmplSetUp.cpp:213-222)
void* ref_svc_create_kind(void* w, int kind, int robot, const double* low, const double* high, double step) {
    RefWorld* W = (RefWorld*)w;
    RefSvc* s = new RefSvc();
    s->space = std::make_shared<RealVectorBeliefSpace>(2);
    ob::RealVectorBounds b(2);
    for (int i = 0; i < 2; ++i) { b.low[i] = low[i]; b.high[i] = high[i]; }
    s->space->setBounds(b);
    s->si = std::make_shared<oc::SpaceInformation>(s->space);
    s->si->setPropagationStepSize(step);
    if (kind == 0) { s->pcc = std::make_shared<PCCBlackmoreSVC>(s->si, W->inst, W->inst->getRobots()[robot], W->inst->getPsafeObs()); s->svc = s->pcc.get(); }
    else { s->adaptive = std::make_shared<AdaptiveRiskBlackmoreSVC>(s->si, W->inst, W->inst->getRobots()[robot], W->inst->getPsafeObs()); s->svc = s->adaptive.get(); }
    s->prop = std::make_shared<R2_UncertainLinearStatePropagator>(s->si);
    return s;
}
void ref_svc_free(void* s) { delete (RefSvc*)s; }
// A[M][4][2], B[M][4]; info: erf_inv_result, p_collision, n_obstacles
void ref_svc_tables(void* s, double* A, double* B, double* info) {
    if (!((RefSvc*)s)->pcc) {
        const AdaptiveRiskBlackmoreSVC& v = *((RefSvc*)s)->adaptive;
        for (size_t o = 0; o < v.A_list_.size(); ++o)
            for (int i = 0; i < 4; ++i) { A[o * 8 + i * 2] = v.A_list_[o](i, 0); A[o * 8 + i * 2 + 1] = v.A_list_[o](i, 1); B[o * 4 + i] = v.B_list_[o](i, 0); }
        info[0] = 0.0; info[1] = v.p_collision_; info[2] = (double)v.A_list_.size();
        return;
    }
    const PCCBlackmoreSVC& v = *((RefSvc*)s)->pcc;
    for (size_t o = 0; o < v.A_list_.size(); ++o)
        for (int i = 0; i < 4; ++i) { A[o * 8 + i * 2] = v.A_list_[o](i, 0); A[o * 8 + i * 2 + 1] = v.A_list_[o](i, 1); B[o * 4 + i] = v.B_list_[o](i, 0); }
    info[0] = v.n_obstacles_ > 0 ? v.erf_inv_result_ : 0.0; info[1] = v.p_collision_; info[2] = v.n_obstacles_;
}
void ref_is_valid(void* s, long n, const double* bel, unsigned char* out) {
    RefSvc* S = (RefSvc*)s;
    auto* st = S->space->allocState()->as<RealVectorBeliefSpace::StateType>();
    for (long i = 0; i < n; ++i) { load(st, bel + i * 10, 2); out[i] = S->svc->isValid(st) ? 1 : 0; }
    S->space->freeState(st);
}
void ref_propagate(void* s, long n, const double* bel, const double* ctrl, double* out) {
    RefSvc* S = (RefSvc*)s;
    auto* a = S->space->allocState()->as<RealVectorBeliefSpace::StateType>();
    auto* r = S->space->allocState()->as<RealVectorBeliefSpace::StateType>();
    oc::RealVectorControlSpace::ControlType c;
    double cv[2];
    c.values = cv;
    for (long i = 0; i < n; ++i) {
        load(a, bel + i * 10, 2);
        cv[0] = ctrl[2 * i]; cv[1] = ctrl[2 * i + 1];
        S->prop->propagate(a, &c, S->si->getPropagationStepSize(), r);
        store(r, out + i * 10, 2);
    }
    S->space->freeState(a); S->space->freeState(r);
}
// OMPL control::SpaceInformation::propagateWhileValid around the reference's propagate / isValid
void ref_rollout(void* s, long n, const double* bel, const double* ctrl, const int* steps, double* out, int* valid) {
    RefSvc* S = (RefSvc*)s;
    auto* state = S->space->allocState()->as<RealVectorBeliefSpace::StateType>();
    auto* result = S->space->allocState()->as<RealVectorBeliefSpace::StateType>();
    oc::RealVectorControlSpace::ControlType c;
    double cv[2];
    c.values = cv;
    const double step = S->si->getPropagationStepSize();
    for (long i = 0; i < n; ++i) {
        load(state, bel + i * 10, 2);
        cv[0] = ctrl[2 * i]; cv[1] = ctrl[2 * i + 1];
        int st = steps[i], r = 0;
        if (st <= 0) { S->space->copyState(result, state); }
        else {
            S->prop->propagate(state, &c, step, result);
            if (S->svc->isValid(result)) {
                ob::State* t1 = result; ob::State* t2 = S->space->allocState(); ob::State* del = t2;
                r = st;
                for (int k = 1; k < st; ++k) {
                    S->prop->propagate(t1, &c, step, t2);
                    if (S->svc->isValid(t2)) std::swap(t1, t2);
                    else { r = k; break; }
                }
                if (result != t1) S->space->copyState(result, t1);
                S->space->freeState(del);
            } else { S->space->copyState(result, state); r = 0; }
        }
        store(result, out + i * 10, 2);
        valid[i] = r;
    }
    S->space->freeState(state); S->space->freeState(result);
}

}  // extern "C"

// ---------------------------------------------------------------------------
// Plan validity checkers: src/PlanValidityCheckers/BeliefPVC.cpp, MinkowskiSumBlackmorePVC.cpp,
// BoundingBoxBlackmorePVC.cpp, src/Constraints/BeliefConstraint.cpp, src/utils/Conflict.cpp compiled in place.
// MultiRobotProblemDefinition is the stand-in of oracle/shim/utils/ (a holder of Instance + SpaceInformation),
// oc::PathControl the stand-in of shim_ompl.hpp (already-interpolated paths only).
// ---------------------------------------------------------------------------
#define private public
#define protected public
#include "PlanValidityCheckers/MinkowskiSumBlackmorePVC.h"
#include "PlanValidityCheckers/BoundingBoxBlackmorePVC.h"
#include "PlanValidityCheckers/AdaptiveRiskBlackmorePVC.h"
#include "PlanValidityCheckers/AdaptiveRiskBoundingBoxPVC.h"
#undef private
#undef protected

namespace {
struct RefPvc {
    RefWorld* world;
    std::vector<oc::SpaceInformationPtr> sis;
    MultiRobotProblemDefinitionPtr pdef;
    std::shared_ptr<BeliefPVC> pvc;
    int kind;
};
oc::PathControl make_path(RefPvc* P, int robot, int len, const double* states, double step) {
    oc::PathControl path(P->sis[robot]);
    auto* st = P->sis[robot]->allocState()->as<RealVectorBeliefSpace::StateType>();
    for (int t = 0; t < len; ++t) {
        load(st, states + (size_t)t * 10, 2);
        if (t == 0) path.append(st); else path.append(st, step);
    }
    P->sis[robot]->freeState(st);
    return path;
}
}  // namespace

extern "C" {

void* ref_pvc_create(void* w, int kind, double step) {
    RefWorld* W = (RefWorld*)w;
    RefPvc* P = new RefPvc();
    P->world = W; P->kind = kind;
    const auto robots = W->inst->getRobots();
    for (size_t r = 0; r < robots.size(); ++r) {
        auto space = std::make_shared<RealVectorBeliefSpace>(2);
        auto si = std::make_shared<oc::SpaceInformation>(space);
        si->setPropagationStepSize(step);
        P->sis.push_back(si);
    }
    P->pdef = std::make_shared<MultiRobotProblemDefinition>(W->inst, P->sis);
    if (kind == 0) P->pvc = std::make_shared<MinkowskiSumBlackmorePVC>(P->pdef, W->inst->getPsafeAgents());      // demos/main.cpp:100-125
    else if (kind == 1) P->pvc = std::make_shared<BoundingBoxBlackmorePVC>(P->pdef, W->inst->getPsafeAgents());
    else if (kind == 2) P->pvc = std::make_shared<AdaptiveRiskBlackmorePVC>(P->pdef, W->inst->getPsafeAgents());
    else P->pvc = std::make_shared<AdaptiveRiskBoundingBoxPVC>(P->pdef, W->inst->getPsafeAgents());
    return P;
}
void ref_pvc_free(void* p) { delete (RefPvc*)p; }
// out: p_coll_agnts, p_coll_dist, erf_inv(1 - 2 p_coll_dist) as the stand-in erf_inv evaluates it
void ref_pvc_info(void* p, double* out) {
    RefPvc* P = (RefPvc*)p;
    out[0] = P->pvc->p_coll_agnts_;
    if (P->kind >= 2) { out[1] = 0; out[2] = 0; return; }       // adaptive checkers allocate the risk per state
    const double pd = P->kind == 0 ? std::static_pointer_cast<MinkowskiSumBlackmorePVC>(P->pvc)->p_coll_dist_
                                   : std::static_pointer_cast<BoundingBoxBlackmorePVC>(P->pvc)->p_coll_dist_;
    out[1] = pd;
    out[2] = boost::math::erf_inv(1 - (2 * pd));
}
// half-planes of the pair (a, b), a < b, as stored in halfPlane_map_
void ref_pvc_pair_halfplanes(void* p, int a, int b, double* A8, double* B4) {
    RefPvc* P = (RefPvc*)p;
    const std::string key = std::to_string(a) + "," + std::to_string(b);
    auto& hp = P->kind == 0 ? std::static_pointer_cast<MinkowskiSumBlackmorePVC>(P->pvc)->halfPlane_map_
             : P->kind == 1 ? std::static_pointer_cast<BoundingBoxBlackmorePVC>(P->pvc)->halfPlane_map_
             : P->kind == 2 ? std::static_pointer_cast<AdaptiveRiskBlackmorePVC>(P->pvc)->halfPlane_map_
                            : std::static_pointer_cast<AdaptiveRiskBoundingBoxPVC>(P->pvc)->halfPlane_map_;
    const auto& pr = hp.at(key);
    for (int i = 0; i < 4; ++i) { A8[2 * i] = pr.first(i, 0); A8[2 * i + 1] = pr.first(i, 1); B4[i] = pr.second(i, 0); }
}
// validatePlan on interpolated trajectories: lens[k], states concatenated AoS; out rows (a, b, step); returns count
int ref_validate_plan(void* p, int k, const int* lens, const double* states, int* out, int cap) {
    RefPvc* P = (RefPvc*)p;
    Plan plan;
    size_t off = 0;
    const double step = P->sis[0]->getPropagationStepSize();
    for (int r = 0; r < k; ++r) { plan.push_back(make_path(P, r, lens[r], states + off * 10, step)); off += lens[r]; }
    std::vector<ConflictPtr> confs = P->pvc->validatePlan(plan);
    int n = 0;
    for (auto& c : confs) { if (n < cap) { out[3 * n] = c->agent1Idx_; out[3 * n + 1] = c->agent2Idx_; out[3 * n + 2] = c->timeStep_; } ++n; }
    return n;
}
// createConstraint(plan, validatePlan(plan), robot): returns the number of entries; out_info = {constrained, constraining};
// times[cap], states[cap][10]
int ref_create_constraint(void* p, int k, const int* lens, const double* states, int robot, int* out_info, double* times, double* cstates, int cap) {
    RefPvc* P = (RefPvc*)p;
    Plan plan;
    size_t off = 0;
    const double step = P->sis[0]->getPropagationStepSize();
    for (int r = 0; r < k; ++r) { plan.push_back(make_path(P, r, lens[r], states + off * 10, step)); off += lens[r]; }
    std::vector<ConflictPtr> confs = P->pvc->validatePlan(plan);
    if (confs.empty()) return 0;
    ConstraintPtr c = P->pvc->createConstraint(plan, confs, robot);
    out_info[0] = c->getConstrainedAgent(); out_info[1] = c->getConstrainingAgent();
    const auto ts = c->getTimes();
    const auto ss = c->as<BeliefConstraint>()->getStates();
    for (size_t i = 0; i < ts.size() && (int)i < cap; ++i) { times[i] = ts[i]; store(ss[i]->as<RealVectorBeliefSpace::StateType>(), cstates + i * 10, 2); }
    return (int)ts.size();
}
// satisfiesConstraints(path, constraints): constraints given as (constraining robot, n_times, steps[], states[][10])
int ref_satisfies_constraints(void* p, int n_path, const double* path_states, int constrained, int n_cons, const int* constraining,
                              const int* n_times, const int* steps, const double* cstates) {
    RefPvc* P = (RefPvc*)p;
    const double step = P->sis[0]->getPropagationStepSize();
    oc::PathControl path = make_path(P, constrained, n_path, path_states, step);
    std::vector<ConstraintPtr> cons;
    size_t off = 0;
    for (int c = 0; c < n_cons; ++c) {
        std::vector<double> times;
        std::vector<ob::State*> sts;
        for (int i = 0; i < n_times[c]; ++i) {
            times.push_back(steps[off + i] * step);                                        // BeliefPVC.cpp:23
            auto* st = P->sis[constraining[c]]->allocState()->as<RealVectorBeliefSpace::StateType>();
            load(st, cstates + (off + i) * 10, 2);
            sts.push_back(st);
        }
        off += n_times[c];
        cons.push_back(std::make_shared<BeliefConstraint>(constrained, constraining[c], times, sts));
    }
    return P->pvc->satisfiesConstraints(path, cons) ? 1 : 0;
}

}  // extern "C"
