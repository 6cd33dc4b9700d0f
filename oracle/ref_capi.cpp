// ref_capi.cpp — TEST INFRASTRUCTURE: C entry points over the REFERENCE'S OWN SOURCES for the headline path,
// compiled from where they lie under /root/reference (never copied):
//   src/utils/common.cpp, src/utils/Instance.cpp                 (map/scen parser, obstacles, robots, risk split)
//   src/StatePropogators/UncertainLinearSP.cpp                    (R2_UncertainLinearStatePropagator::propagate)
//   src/StateValidityCheckers/PCCBlackmoreSVC.cpp                 (Minkowski sums, half-planes, erf_inv constant, isValid)
// against the stand-in headers of oracle/shim/ (Eigen / Boost / OMPL subsets, see the notes there: every matrix
// product on this path has inner dimension 2, so the shim's evaluation order cannot differ from Eigen's).
// RealVectorBeliefSpace.cpp is NOT compiled (its distance() needs Eigen's eigen-solver); its five one-line
// members are defined below following src/Spaces/RealVectorBeliefSpace.cpp:7-14,64-76.
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
#undef private
#include "StatePropogators/UncertainLinearSP.h"

// ---- RealVectorBeliefSpace members (src/Spaces/RealVectorBeliefSpace.cpp:7-14, 64-76) ----
ompl::base::State* RealVectorBeliefSpace::allocState(void) const {
    StateType* rstate = new StateType();
    rstate->values = new double[dimension_];
    rstate->sigma_ = 0.01 * Eigen::MatrixXd::Identity(dimension_, dimension_);
    rstate->lambda_ = 0.01 * Eigen::MatrixXd::Identity(dimension_, dimension_);
    return rstate;
}
void RealVectorBeliefSpace::copyState(State* destination, const State* source) const {
    RealVectorStateSpace::copyState(destination, source);
    destination->as<StateType>()->sigma_ = source->as<StateType>()->sigma_;
    destination->as<StateType>()->lambda_ = source->as<StateType>()->lambda_;
}
void RealVectorBeliefSpace::freeState(State* state) const {
    StateType* s = state->as<StateType>();
    delete[] s->values;
    delete s;
}
double RealVectorBeliefSpace::distance(const State*, const State*) const { return 0.0; }   // not on the pinned path
void RealVectorBeliefSpace::printState(const State*, std::ostream&) const {}

namespace {
struct RefWorld { InstancePtr inst; };
struct RefSvc {
    std::shared_ptr<RealVectorBeliefSpace> space;
    oc::SpaceInformationPtr si;
    std::shared_ptr<PCCBlackmoreSVC> svc;
    std::shared_ptr<R2_UncertainLinearStatePropagator> prop;
};
struct Quiet {   // Instance's constructor prints every obstacle
    std::streambuf* old;
    std::ostringstream sink;
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
void* ref_svc_create(void* w, int robot, const double* low, const double* high, double step) {
    RefWorld* W = (RefWorld*)w;
    RefSvc* s = new RefSvc();
    s->space = std::make_shared<RealVectorBeliefSpace>(2);
    ob::RealVectorBounds b(2);
    for (int i = 0; i < 2; ++i) { b.low[i] = low[i]; b.high[i] = high[i]; }
    s->space->setBounds(b);
    s->si = std::make_shared<oc::SpaceInformation>(s->space);
    s->si->setPropagationStepSize(step);
    s->svc = std::make_shared<PCCBlackmoreSVC>(s->si, W->inst, W->inst->getRobots()[robot], W->inst->getPsafeObs());
    s->prop = std::make_shared<R2_UncertainLinearStatePropagator>(s->si);
    return s;
}
void ref_svc_free(void* s) { delete (RefSvc*)s; }
// A[M][4][2], B[M][4]; info: erf_inv_result, p_collision, n_obstacles
void ref_svc_tables(void* s, double* A, double* B, double* info) {
    const PCCBlackmoreSVC& v = *((RefSvc*)s)->svc;
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
