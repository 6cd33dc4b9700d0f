// ref_capi2.cpp — TEST INFRASTRUCTURE (round 2): C entry points over MORE of the reference's own sources, compiled from
// where they lie under /root/reference (never copied) against the stand-in headers of oracle/shim/:
//   src/StatePropogators/UncertainUnicycleSP.cpp        DynUnicycleControlSpace ctor (MatrixXd::exp, zero-order hold) + propagate
//   src/Spaces/RealVectorBeliefSpace.cpp                allocState / copyState / freeState / distance (SelfAdjointEigenSolver)
//   src/Goals/BeliefSpaceGoals.cpp                      ChanceConstrainedGoal::isSatisfied (EigenSolver, bg::buffer, bg::within)
//   src/StateValidityCheckers/ChiSquaredBoundarySVC.cpp isValid (EigenSolver, bg::buffer, bg::disjoint)
//   src/PlanValidityCheckers/ChiSquaredBoundaryPVC.cpp  validatePlan / satisfiesConstraints / isSafe_
//   src/PlanValidityCheckers/CDFGridPVC.cpp, Blackmore2PVC.cpp   
// What these pins can and cannot prove is stated in DESIGN.md section 4: the reference's own control flow, constants,
// sub-block choices and operation order are executed as written; the third-party numerics underneath (Eigen's 4x4 inverse,
// eigen-solvers and Pade exponential, Boost.Geometry's buffer / within / disjoint, Boost.Math's quantile / erf) are the
// stand-ins' restatements of the published algorithms, so values are pinned to 1e-9 relative and verdicts outside the
// +-1e-6 band, not to the last bit.
// Records are AoS: [values[dim], Sigma row-major, Lambda row-major], W = dim + 2 dim^2 doubles.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "shim_eigen.hpp"
#include "shim_boost.hpp"
#include "shim_ompl.hpp"

#define private public
#define protected public
#include "StateValidityCheckers/PCCBlackmoreSVC.h"
#include "StateValidityCheckers/AdaptiveRiskBlackmoreSVC.h"
#include "StateValidityCheckers/ChiSquaredBoundarySVC.h"
#include "StatePropogators/UncertainLinearSP.h"
#include "StatePropogators/UncertainUnicycleSP.h"
#include "StatePropogators/CentralizedUncertainLinearSP.h"
#include "Goals/BeliefSpaceGoals.h"
#include "PlanValidityCheckers/MinkowskiSumBlackmorePVC.h"
#include "PlanValidityCheckers/BoundingBoxBlackmorePVC.h"
#include "PlanValidityCheckers/AdaptiveRiskBlackmorePVC.h"
#include "PlanValidityCheckers/AdaptiveRiskBoundingBoxPVC.h"
#include "PlanValidityCheckers/ChiSquaredBoundaryPVC.h"
#include "PlanValidityCheckers/CDFGridPVC.h"
#include "PlanValidityCheckers/Blackmore2PVC.h"
#undef private
#undef protected

namespace {
struct RefWorld { InstancePtr inst; };   // same layout as in ref_capi.cpp (the handle comes from ref_world_create)
struct Quiet {                           // the reference prints from constructors and from the goal test
    std::ostringstream sink;
    std::streambuf* old;
    Quiet() : old(std::cout.rdbuf(sink.rdbuf())) {}
    ~Quiet() { std::cout.rdbuf(old); }
};
struct Ref2Svc {
    int dim = 2;
    std::shared_ptr<RealVectorBeliefSpace> space;
    oc::SpaceInformationPtr si;
    std::shared_ptr<ob::StateValidityChecker> svc;
    std::shared_ptr<oc::StatePropagator> prop;
};
typedef RealVectorBeliefSpace::StateType BState;
void load(BState* s, const double* rec, int dim) {
    if (s->sigma_.rows() != dim) { s->sigma_ = Eigen::MatrixXd(dim, dim); s->lambda_ = Eigen::MatrixXd(dim, dim); }
    for (int i = 0; i < dim; ++i) s->values[i] = rec[i];
    for (int i = 0; i < dim; ++i) for (int j = 0; j < dim; ++j) { s->sigma_(i, j) = rec[dim + i * dim + j]; s->lambda_(i, j) = rec[dim + dim * dim + i * dim + j]; }
}
void store(const BState* s, double* rec, int dim) {
    for (int i = 0; i < dim; ++i) rec[i] = s->values[i];
    for (int i = 0; i < dim; ++i) for (int j = 0; j < dim; ++j) { rec[dim + i * dim + j] = s->sigma_(i, j); rec[dim + dim * dim + i * dim + j] = s->lambda_(i, j); }
}
inline int width(int dim) { return dim + 2 * dim * dim; }
}  // namespace

extern "C" {

// The wiring of set_up_ConstraintBSST_MP_Problems (src/utils/OmplSetUp.cpp:183-353) for one robot.
// model 0: RealVectorBeliefSpace(2) + R2_UncertainLinearStatePropagator; 1: RealVectorBeliefSpace(4) + DynUnicycleControlSpace.
// svc_kind 0 PCCBlackmoreSVC, 1 AdaptiveRiskBlackmoreSVC, 2 ChiSquaredBoundarySVC; accep_prob is the constructor's argument
// (the caller passes getPsafeObs() resp. the hard-coded 0.9, OmplSetUp.cpp:213-222).
void* ref2_svc_create(void* w, int svc_kind, int robot, int model, const double* low, const double* high, double step, double accep_prob) {
    RefWorld* W = (RefWorld*)w;
    Quiet q;
    Ref2Svc* s = new Ref2Svc();
    s->dim = model == 1 ? 4 : 2;
    s->space = std::make_shared<RealVectorBeliefSpace>(s->dim);
    ob::RealVectorBounds b(s->dim);
    for (int i = 0; i < s->dim; ++i) { b.low[i] = low[i]; b.high[i] = high[i]; }
    s->space->setBounds(b);
    s->si = std::make_shared<oc::SpaceInformation>(s->space);
    s->si->setPropagationStepSize(step);
    const Robot* r = W->inst->getRobots()[robot];
    if (svc_kind == 0) s->svc = std::make_shared<PCCBlackmoreSVC>(s->si, W->inst, r, accep_prob);
    else if (svc_kind == 1) s->svc = std::make_shared<AdaptiveRiskBlackmoreSVC>(s->si, W->inst, r, accep_prob);
    else s->svc = std::make_shared<ChiSquaredBoundarySVC>(s->si, W->inst, r, accep_prob);
    if (model == 1) s->prop = std::make_shared<DynUnicycleControlSpace>(s->si);
    else s->prop = std::make_shared<R2_UncertainLinearStatePropagator>(s->si);
    return s;
}
void ref2_svc_free(void* s) { delete (Ref2Svc*)s; }

// the matrices the DynUnicycleControlSpace constructor computed (UncertainUnicycleSP.cpp:88-127), row-major
void ref2_uni_params(void* s, double* F, double* Acl, double* Q, double* R) {
    auto* p = dynamic_cast<DynUnicycleControlSpace*>(((Ref2Svc*)s)->prop.get());
    if (!p) return;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { F[i * 4 + j] = p->F(i, j); Acl[i * 4 + j] = p->A_cl_d_(i, j); Q[i * 4 + j] = p->Q(i, j); R[i * 4 + j] = p->R(i, j); }
}
// ChiSquaredBoundarySVC's constants: sc_, max_rad_, number of inflated obstacle polygons
void ref2_chisq_info(void* s, double* out) {
    auto* p = dynamic_cast<ChiSquaredBoundarySVC*>(((Ref2Svc*)s)->svc.get());
    if (!p) { out[0] = out[1] = out[2] = 0; return; }
    out[0] = p->sc_; out[1] = p->max_rad_; out[2] = (double)p->obs_list_.size();
}

void ref2_propagate(void* s, long n, const double* bel, const double* ctrl, double duration, double* out) {
    Ref2Svc* S = (Ref2Svc*)s;
    const int d = S->dim, W = width(d);
    auto* a = S->space->allocState()->as<BState>();
    auto* r = S->space->allocState()->as<BState>();
    oc::RealVectorControlSpace::ControlType c;
    double cv[4];
    c.values = cv;
    for (long i = 0; i < n; ++i) {
        load(a, bel + i * W, d);
        for (int j = 0; j < d; ++j) cv[j] = ctrl[d * i + j];
        S->prop->propagate(a, &c, duration, r);
        store(r, out + i * W, d);
    }
    S->space->freeState(a); S->space->freeState(r);
}
void ref2_is_valid(void* s, long n, const double* bel, unsigned char* out) {
    Ref2Svc* S = (Ref2Svc*)s;
    const int d = S->dim, W = width(d);
    auto* st = S->space->allocState()->as<BState>();
    for (long i = 0; i < n; ++i) { load(st, bel + i * W, d); out[i] = S->svc->isValid(st) ? 1 : 0; }
    S->space->freeState(st);
}
// OMPL control::SpaceInformation::propagateWhileValid (src/ompl/control/src/SpaceInformation.cpp, OMPL 1.6) around the
// reference's two virtuals
void ref2_rollout(void* s, long n, const double* bel, const double* ctrl, const int* steps, double* out, int* valid) {
    Ref2Svc* S = (Ref2Svc*)s;
    const int d = S->dim, W = width(d);
    auto* state = S->space->allocState()->as<BState>();
    auto* result = S->space->allocState()->as<BState>();
    oc::RealVectorControlSpace::ControlType c;
    double cv[4];
    c.values = cv;
    const double step = S->si->getPropagationStepSize();
    for (long i = 0; i < n; ++i) {
        load(state, bel + i * W, d);
        for (int j = 0; j < d; ++j) cv[j] = ctrl[d * i + j];
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
        store(result, out + i * W, d);
        valid[i] = r;
    }
    S->space->freeState(state); S->space->freeState(result);
}

// RealVectorBeliefSpace::distance(state1 = a_i, state2 = b_i)   (src/Spaces/RealVectorBeliefSpace.cpp:16-62)
void ref2_distance(int dim, long n, const double* a, const double* b, double* out) {
    RealVectorBeliefSpace space(dim);
    const int W = width(dim);
    auto* s1 = space.allocState()->as<BState>();
    auto* s2 = space.allocState()->as<BState>();
    for (long i = 0; i < n; ++i) { load(s1, a + i * W, dim); load(s2, b + i * W, dim); out[i] = space.distance(s1, s2); }
    space.freeState(s1); space.freeState(s2);
}

// ChanceConstrainedGoal(si, goal, toll, p_safe)::isSatisfied(state, &distance)   (src/Goals/BeliefSpaceGoals.cpp:9-147)
void ref2_goal(int dim, long n, const double* bel, double gx, double gy, double toll, double p_safe, unsigned char* ok, double* dist) {
    Quiet q;   // isSafe_ prints distances below 2.0 (:102-103)
    auto space = std::make_shared<RealVectorBeliefSpace>(dim);
    auto si = std::make_shared<oc::SpaceInformation>(space);
    ChanceConstrainedGoal goal(si, Location(gx, gy), toll, p_safe);
    const int W = width(dim);
    auto* st = space->allocState()->as<BState>();
    for (long i = 0; i < n; ++i) {
        load(st, bel + i * W, dim);
        double d = 0;
        ok[i] = goal.isSatisfied(st, &d) ? 1 : 0;
        dist[i] = d;
    }
    space->freeState(st);
}

// CentralizedUncertainLinearStatePropagator::propagate (src/StatePropogators/CentralizedUncertainLinearSP.cpp:42-70) on n
// composite states of k agents (dimension 2k)
void ref2_central_propagate(int k, long n, const double* states, const double* ctrl, double step, double* out) {
    const int dim = 2 * k, W = width(dim);
    auto space = std::make_shared<RealVectorBeliefSpace>(dim);
    auto si = std::make_shared<oc::SpaceInformation>(space);
    si->setPropagationStepSize(step);
    CentralizedUncertainLinearStatePropagator prop(si);
    auto* a = space->allocState()->as<BState>();
    auto* r = space->allocState()->as<BState>();
    oc::RealVectorControlSpace::ControlType c;
    std::vector<double> cv(dim);
    c.values = cv.data();
    for (long i = 0; i < n; ++i) {
        load(a, states + i * W, dim);
        for (int j = 0; j < dim; ++j) cv[j] = ctrl[i * dim + j];
        prop.propagate(a, &c, step, r);
        store(r, out + i * W, dim);
    }
    space->freeState(a); space->freeState(r);
}

}  // extern "C"

// ---------------------------------------------------------------------------
// plan validity checkers for either state dimension (BeliefPVC::getDistribution_ picks a different sub-block for 4-D)
// ---------------------------------------------------------------------------
namespace {
struct Ref2Pvc {
    RefWorld* world;
    int dim, kind;
    std::vector<oc::SpaceInformationPtr> sis;
    MultiRobotProblemDefinitionPtr pdef;
    std::shared_ptr<BeliefPVC> pvc;
};
oc::PathControl make_path(Ref2Pvc* P, int robot, int len, const double* states, double step) {
    oc::PathControl path(P->sis[robot]);
    auto* st = P->sis[robot]->allocState()->as<BState>();
    const int W = width(P->dim);
    for (int t = 0; t < len; ++t) {
        load(st, states + (size_t)t * W, P->dim);
        if (t == 0) path.append(st); else path.append(st, step);
    }
    P->sis[robot]->freeState(st);
    return path;
}
}  // namespace

extern "C" {

// kind 0 MinkowskiSumBlackmorePVC, 1 BoundingBoxBlackmorePVC, 2 AdaptiveRiskBlackmorePVC, 3 AdaptiveRiskBoundingBoxPVC,
// 4 ChiSquaredBoundaryPVC, 5 CDFGridPVC(p_arg, n_steps), 6 Blackmore2PVC; p_arg is the constructor's p_safe
// (demos/main.cpp:105-138: getPsafeAgents(), or the hard-coded 0.9 for ChiSquared)
void* ref2_pvc_create(void* w, int kind, int dim, double step, double p_arg, int n_steps) {
    RefWorld* W = (RefWorld*)w;
    Quiet q;
    Ref2Pvc* P = new Ref2Pvc();
    P->world = W; P->dim = dim; P->kind = kind;
    const auto robots = W->inst->getRobots();
    for (size_t r = 0; r < robots.size(); ++r) {
        auto space = std::make_shared<RealVectorBeliefSpace>(dim);
        auto si = std::make_shared<oc::SpaceInformation>(space);
        si->setPropagationStepSize(step);
        P->sis.push_back(si);
    }
    P->pdef = std::make_shared<MultiRobotProblemDefinition>(W->inst, P->sis);
    if (kind == 0) P->pvc = std::make_shared<MinkowskiSumBlackmorePVC>(P->pdef, p_arg);
    else if (kind == 1) P->pvc = std::make_shared<BoundingBoxBlackmorePVC>(P->pdef, p_arg);
    else if (kind == 2) P->pvc = std::make_shared<AdaptiveRiskBlackmorePVC>(P->pdef, p_arg);
    else if (kind == 3) P->pvc = std::make_shared<AdaptiveRiskBoundingBoxPVC>(P->pdef, p_arg);
    else if (kind == 4) P->pvc = std::make_shared<ChiSquaredBoundaryPVC>(P->pdef, p_arg);
    else if (kind == 5) P->pvc = std::make_shared<CDFGridPVC>(P->pdef, p_arg, n_steps);
    else P->pvc = std::make_shared<Blackmore2PVC>(P->pdef, p_arg);
    return P;
}
void ref2_pvc_free(void* p) { delete (Ref2Pvc*)p; }
// out: p_coll_agnts_, sc_ (ChiSquared / CDFGrid, else 0), p_coll_dist_ (Blackmore2, else 0)
void ref2_pvc_info(void* p, double* out) {
    Ref2Pvc* P = (Ref2Pvc*)p;
    out[0] = P->pvc->p_coll_agnts_; out[1] = 0; out[2] = 0;
    if (P->kind == 4) out[1] = std::static_pointer_cast<ChiSquaredBoundaryPVC>(P->pvc)->sc_;
    if (P->kind == 5) out[1] = std::static_pointer_cast<CDFGridPVC>(P->pvc)->sc_;
    if (P->kind == 6) out[2] = std::static_pointer_cast<Blackmore2PVC>(P->pvc)->p_coll_dist_;
}
int ref2_validate_plan(void* p, int k, const int* lens, const double* states, int* out, int cap) {
    Ref2Pvc* P = (Ref2Pvc*)p;
    Plan plan;
    size_t off = 0;
    const int W = width(P->dim);
    const double step = P->sis[0]->getPropagationStepSize();
    for (int r = 0; r < k; ++r) { plan.push_back(make_path(P, r, lens[r], states + off * W, step)); off += lens[r]; }
    std::vector<ConflictPtr> confs = P->pvc->validatePlan(plan);
    int n = 0;
    for (auto& c : confs) { if (n < cap) { out[3 * n] = c->agent1Idx_; out[3 * n + 1] = c->agent2Idx_; out[3 * n + 2] = c->timeStep_; } ++n; }
    return n;
}
int ref2_satisfies_constraints(void* p, int n_path, const double* path_states, int constrained, int n_cons, const int* constraining,
                               const int* n_times, const int* steps, const double* cstates) {
    Ref2Pvc* P = (Ref2Pvc*)p;
    const int W = width(P->dim);
    const double step = P->sis[0]->getPropagationStepSize();
    oc::PathControl path = make_path(P, constrained, n_path, path_states, step);
    std::vector<ConstraintPtr> cons;
    size_t off = 0;
    for (int c = 0; c < n_cons; ++c) {
        std::vector<double> times;
        std::vector<ob::State*> sts;
        for (int i = 0; i < n_times[c]; ++i) {
            times.push_back(steps[off + i] * step);                                        // BeliefPVC.cpp:23
            auto* st = P->sis[constraining[c]]->allocState()->as<BState>();
            load(st, cstates + (off + i) * W, P->dim);
            sts.push_back(st);
        }
        off += n_times[c];
        cons.push_back(std::make_shared<BeliefConstraint>(constrained, constraining[c], times, sts));
    }
    return P->pvc->satisfiesConstraints(path, cons) ? 1 : 0;
}
// the pair test in isolation: *PVC::isSafe_ through independentCheck where the class has one (ChiSquared, CDFGrid,
// Blackmore2: robots "Robot 0" / "Robot 1"), for n pairs of states
void ref2_independent_check(void* p, long n, const double* a, const double* b, unsigned char* out) {
    Ref2Pvc* P = (Ref2Pvc*)p;
    const int W = width(P->dim);
    auto* s1 = P->sis[0]->allocState()->as<BState>();
    auto* s2 = P->sis[0]->allocState()->as<BState>();
    for (long i = 0; i < n; ++i) {
        load(s1, a + i * W, P->dim); load(s2, b + i * W, P->dim);
        bool ok = false;
        if (P->kind == 4) ok = std::static_pointer_cast<ChiSquaredBoundaryPVC>(P->pvc)->independentCheck(s1, s2);
        else if (P->kind == 5) ok = std::static_pointer_cast<CDFGridPVC>(P->pvc)->independentCheck(s1, s2);
        else if (P->kind == 6) ok = std::static_pointer_cast<Blackmore2PVC>(P->pvc)->independentCheck(s1, s2);
        out[i] = ok ? 1 : 0;
    }
    P->sis[0]->freeState(s1); P->sis[0]->freeState(s2);
}

}  // extern "C"
