// host_selftest — exercises the OMPL-shaped adapters one call at a time (the drop-in surface):
// oc::SpaceInformation::propagateWhileValid driving R2_UncertainLinearStatePropagator::propagate
// and PCCBlackmoreSVC::isValid through the C ABI must agree with one batched
// cck_propagate_while_valid launch on the same inputs; PathControl::interpolate + validatePlan +
// createConstraint + satisfiesConstraints round trip.
#include <cstdio>
#include <random>

#include "planner.hpp"

using namespace cck_b200;

#define REQUIRE(c) do { if (!(c)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

int main(int argc, char** argv) {
    if (argc < 3) { std::printf("usage: host_selftest map scen\n"); return 2; }
    try {
        auto inst = std::make_shared<Instance>(argv[1], argv[2], 2, 0.9);
        PlannerParams prm; prm.K = 256;
        BatchedBSST pl(inst, 0, "Blackmore", "Blackmore", prm);
        auto si = pl.getSpaceInformation();
        const int dim = pl.dim(), W = belief_width(dim);
        std::mt19937_64 rng(5);
        std::uniform_real_distribution<double> U(0, 1);
        const int n = 200;
        std::vector<double> bel((size_t)W * n), ctrl((size_t)dim * n), out((size_t)W * n);
        std::vector<int32_t> steps(n), valid(n);
        ob::State* s = si->allocState();
        ob::State* res = si->allocState();
        oc::Control* c = si->allocControl();
        int agree = 0, full = 0;
        for (int i = 0; i < n; ++i) {
            auto* b = s->as<RealVectorBeliefSpace::StateType>();
            b->values[0] = U(rng) * inst->getDimensions()[0]; b->values[1] = U(rng) * inst->getDimensions()[1];
            for (int d = 0; d < dim; ++d) c->as<oc::RealVectorControlSpace::ControlType>()->values[d] = 2 * U(rng) - 1;
            gather_state(s, dim, bel.data(), n, i);
            for (int d = 0; d < dim; ++d) ctrl[(size_t)d * n + i] = c->as<oc::RealVectorControlSpace::ControlType>()->values[d];
            steps[i] = 1 + (int)(rng() % 10);
        }
        pl.gpu()->check(cck_propagate_while_valid(pl.gpu()->get(), n, bel.data(), n, ctrl.data(), n, steps.data(), out.data(), n, valid.data(), nullptr, 0, 0), "batch");
        for (int i = 0; i < n; ++i) {
            scatter_state(bel.data(), n, i, dim, s);
            for (int d = 0; d < dim; ++d) c->as<oc::RealVectorControlSpace::ControlType>()->values[d] = ctrl[(size_t)d * n + i];
            const unsigned r = si->propagateWhileValid(s, c, steps[i], res);      // OMPL loop, one virtual call per step
            double one[36];
            gather_state(res, dim, one, 1, 0);
            bool same = (int)r == valid[i];
            for (int f = 0; f < W; ++f) same = same && one[f] == out[(size_t)f * n + i];
            agree += same; full += ((int)r == steps[i]);
        }
        std::printf("adapter-vs-batch: %d/%d identical, %d fully valid\n", agree, n, full);
        REQUIRE(agree == n);
        REQUIRE(full > 0 && full < n);

        // PathControl::interpolate through the adapters, then the PVC round trip
        oc::PathControl path0(si), path1(si);
        auto make_path = [&](oc::PathControl& p, double x, double y, double ux, double uy, int segs) {
            auto* b = s->as<RealVectorBeliefSpace::StateType>();
            si->copyState(s, si->allocState());
            b->values[0] = x; b->values[1] = y;
            for (int k = 0; k < segs; ++k) {
                c->as<oc::RealVectorControlSpace::ControlType>()->values[0] = ux; c->as<oc::RealVectorControlSpace::ControlType>()->values[1] = uy;
                p.append(s, c, 5 * 0.2);
                std::vector<ob::State*> v;
                si->propagate(s, c, 5, v, true);
                si->copyState(s, v.back());
            }
            p.append(s);
            p.interpolate();
        };
        make_path(path0, 1.0, 4.0, 1.0, 0.0, 6);
        make_path(path1, 7.0, 4.0, -1.0, 0.0, 6);
        REQUIRE(path0.getStateCount() == 31 && path1.getStateCount() == 31);
        auto to_traj = [&](oc::PathControl& p) { Trajectory t; for (auto* st : p.getStates()) { BeliefState b; b.dim = dim; b.v.resize(W); gather_state(st, dim, b.v.data(), 1, 0); t.push_back(b); } return t; };
        Plan plan{to_traj(path0), to_traj(path1)};
        auto pv = pl.getPlanValidator();
        auto confs = pv->validatePlan(plan);
        std::printf("validatePlan: %zu conflicts, first at step %d between %d and %d\n", confs.size(), confs.empty() ? -1 : confs[0]->timeStep_,
                    confs.empty() ? -1 : confs[0]->agent1Idx_, confs.empty() ? -1 : confs[0]->agent2Idx_);
        REQUIRE(!confs.empty() && confs[0]->agent1Idx_ == 0 && confs[0]->agent2Idx_ == 1);
        auto con = pv->createConstraint(plan, confs, 0);
        REQUIRE(con->getConstrainingAgent() == 1 && con->getTimes().size() == confs.size());
        REQUIRE(!pv->satisfiesConstraints(plan[0], {con}));
        Trajectory stay(plan[0].size(), plan[0][0]);
        REQUIRE(pv->satisfiesConstraints(stay, {con}));
        // ---- centralized variants: composite state of k = 2 agents ----
        {
            const int k = 2;
            auto cspace4 = std::make_shared<RealVectorBeliefSpace>(2 * k);
            ob::RealVectorBounds b4(2 * k);
            for (int a = 0; a < k; ++a) { b4.setLow(2 * a, -1); b4.setHigh(2 * a, inst->getDimensions()[0]); b4.setLow(2 * a + 1, -1); b4.setHigh(2 * a + 1, inst->getDimensions()[1]); }
            cspace4->setBounds(b4);
            auto ctl4 = std::make_shared<oc::RealVectorControlSpace>(cspace4, 2 * k);
            auto si4 = std::make_shared<oc::SpaceInformation>(cspace4, ctl4);
            si4->setPropagationStepSize(0.5);                                             // OmplSetUp.cpp:357
            auto g4 = std::make_shared<GpuContext>(0);
            si4->setStatePropagator(std::make_shared<CentralizedUncertainLinearStatePropagator>(si4, g4));
            auto g5 = std::make_shared<GpuContext>(0);
            si4->setStateValidityChecker(std::make_shared<CentralizedChiSquaredBoundarySVC>(si4, g5, inst, 0.9));
            ob::State* cs = si4->allocState();
            ob::State* cr = si4->allocState();
            oc::Control* cc = si4->allocControl();
            auto* cb = cs->as<RealVectorBeliefSpace::StateType>();
            cb->values[0] = 1.0; cb->values[1] = 2.0; cb->values[2] = 5.0; cb->values[3] = 2.0;
            cb->sigma_(0, 1) = cb->sigma_(1, 0) = 0.002; cb->sigma_(0, 3) = 0.5;          // an off-block entry: must not influence anything
            double* cu = cc->as<oc::RealVectorControlSpace::ControlType>()->values;
            cu[0] = 1.0; cu[1] = 0.0; cu[2] = -1.0; cu[3] = 0.5;
            REQUIRE(si4->isValid(cs));
            si4->getStatePropagator()->propagate(cs, cc, 0.5, cr);
            auto* rb = cr->as<RealVectorBeliefSpace::StateType>();
            // each agent's block against the single-agent batched propagator on a context with the same step size
            auto g6 = std::make_shared<GpuContext>(0);
            install_model(g6, CCK_MODEL_LINEAR2D, 0.5);
            for (int a = 0; a < k; ++a) {
                double in1[10], u1[2] = {cu[2 * a], cu[2 * a + 1]}, out1[10];
                in1[0] = cb->values[2 * a]; in1[1] = cb->values[2 * a + 1];
                for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) { in1[2 + 2 * i + j] = cb->sigma_(2 * a + i, 2 * a + j); in1[6 + 2 * i + j] = cb->lambda_(2 * a + i, 2 * a + j); }
                g6->check(cck_propagate(g6->get(), 1, in1, 1, u1, 1, 0.5, out1, 1), "single");
                REQUIRE(rb->values[2 * a] == out1[0] && rb->values[2 * a + 1] == out1[1]);
                for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) REQUIRE(rb->sigma_(2 * a + i, 2 * a + j) == out1[2 + 2 * i + j] && rb->lambda_(2 * a + i, 2 * a + j) == out1[6 + 2 * i + j]);
            }
            REQUIRE(rb->values[0] == 1.0 + 0.5 * 1.0 && rb->values[2] == 5.0 + 0.5 * -1.0);
            // agents on top of each other: the decagons meet -> invalid; outside the bounds -> invalid
            cb->values[2] = 1.2; cb->values[3] = 2.1;
            REQUIRE(!si4->isValid(cs));
            cb->values[2] = 5.0; cb->values[3] = 2.0; cb->values[1] = -1.5;
            REQUIRE(!si4->isValid(cs));
            std::printf("centralized adapters OK\n");
        }
        std::printf("host_selftest OK\n");
        return 0;
    } catch (const std::exception& e) {
        std::printf("host_selftest exception: %s\n", e.what());
        return 3;
    }
}
