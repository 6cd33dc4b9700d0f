// ompl_shim.hpp — just enough of the OMPL 1.6 API surface for the hot-path plugins to compile
// and be exercised where OMPL is not installed (this image has no OMPL/Eigen/Boost).
// When the real headers are present (__has_include) they are used instead and the adapter
// classes in adapters.hpp derive from the genuine ompl::control::StatePropagator /
// ompl::base::StateValidityChecker.
//
// Semantics restated from OMPL 1.6.0 (source not in the container):
//   RealVectorStateSpace::satisfiesBounds, control::SpaceInformation::propagateWhileValid,
//   control::SpaceInformation::propagate(n steps), control::PathControl::interpolate.
#pragma once
#if defined(__has_include)
#if __has_include(<ompl/control/StatePropagator.h>) && !defined(CCK_FORCE_OMPL_SHIM)
#define CCK_HAVE_OMPL 1
#endif
#endif

#ifdef CCK_HAVE_OMPL
#include <ompl/base/StateValidityChecker.h>
#include <ompl/base/spaces/RealVectorStateSpace.h>
#include <ompl/control/PathControl.h>
#include <ompl/control/SpaceInformation.h>
#include <ompl/control/StatePropagator.h>
#include <ompl/control/spaces/RealVectorControlSpace.h>
#else
#include <cfloat>
#include <cmath>
#include <memory>
#include <vector>

namespace ompl {
namespace base {

class State {
public:
    virtual ~State() = default;
    template <class T> T* as() { return static_cast<T*>(this); }
    template <class T> const T* as() const { return static_cast<const T*>(this); }
};

struct RealVectorBounds {
    std::vector<double> low, high;
    explicit RealVectorBounds(unsigned dim = 0) : low(dim, 0.0), high(dim, 0.0) {}
    void setLow(unsigned i, double v) { low[i] = v; }
    void setHigh(unsigned i, double v) { high[i] = v; }
    void setLow(double v) { for (auto& x : low) x = v; }
    void setHigh(double v) { for (auto& x : high) x = v; }
};

class StateSpace {
public:
    virtual ~StateSpace() = default;
    virtual unsigned getDimension() const = 0;
    virtual State* allocState() const = 0;
    virtual void freeState(State* s) const = 0;
    virtual void copyState(State* dst, const State* src) const = 0;
    virtual bool satisfiesBounds(const State* s) const = 0;
    template <class T> T* as() { return static_cast<T*>(this); }
    template <class T> const T* as() const { return static_cast<const T*>(this); }
};
using StateSpacePtr = std::shared_ptr<StateSpace>;

class RealVectorStateSpace : public StateSpace {
public:
    class StateType : public State {
    public:
        double* values = nullptr;
    };
    explicit RealVectorStateSpace(unsigned dim) : dimension_(dim), bounds_(dim) {}
    unsigned getDimension() const override { return dimension_; }
    void setBounds(const RealVectorBounds& b) { bounds_ = b; }
    const RealVectorBounds& getBounds() const { return bounds_; }
    State* allocState() const override { auto* s = new StateType(); s->values = new double[dimension_](); return s; }
    void freeState(State* s) const override { delete[] s->as<StateType>()->values; delete s; }
    void copyState(State* d, const State* s) const override {
        for (unsigned i = 0; i < dimension_; ++i) d->as<StateType>()->values[i] = s->as<StateType>()->values[i];
    }
    // OMPL 1.6: v - eps > high || v + eps < low  => false   (a NaN component passes)
    bool satisfiesBounds(const State* s) const override {
        const auto* r = s->as<StateType>();
        for (unsigned i = 0; i < dimension_; ++i)
            if (r->values[i] - DBL_EPSILON > bounds_.high[i] || r->values[i] + DBL_EPSILON < bounds_.low[i]) return false;
        return true;
    }
protected:
    unsigned dimension_;
    RealVectorBounds bounds_;
};

class SpaceInformation;
class StateValidityChecker {
public:
    explicit StateValidityChecker(SpaceInformation* si) : si_(si) {}
    virtual ~StateValidityChecker() = default;
    virtual bool isValid(const State* state) const = 0;
protected:
    SpaceInformation* si_;
};
using StateValidityCheckerPtr = std::shared_ptr<StateValidityChecker>;

class SpaceInformation {
public:
    explicit SpaceInformation(StateSpacePtr space) : space_(std::move(space)) {}
    virtual ~SpaceInformation() = default;
    const StateSpacePtr& getStateSpace() const { return space_; }
    void setStateValidityChecker(const StateValidityCheckerPtr& c) { svc_ = c; }
    const StateValidityCheckerPtr& getStateValidityChecker() const { return svc_; }
    bool isValid(const State* s) const { return svc_->isValid(s); }
    bool satisfiesBounds(const State* s) const { return space_->satisfiesBounds(s); }
    State* allocState() const { return space_->allocState(); }
    void freeState(State* s) const { space_->freeState(s); }
    void copyState(State* d, const State* s) const { space_->copyState(d, s); }
    State* cloneState(const State* s) const { State* c = allocState(); copyState(c, s); return c; }
protected:
    StateSpacePtr space_;
    StateValidityCheckerPtr svc_;
};
using SpaceInformationPtr = std::shared_ptr<SpaceInformation>;

}  // namespace base

namespace control {

class Control {
public:
    virtual ~Control() = default;
    template <class T> T* as() { return static_cast<T*>(this); }
    template <class T> const T* as() const { return static_cast<const T*>(this); }
};

class RealVectorControlSpace {
public:
    class ControlType : public Control {
    public:
        double* values = nullptr;
    };
    RealVectorControlSpace(const base::StateSpacePtr&, unsigned dim) : dimension_(dim), bounds_(dim) {}
    unsigned getDimension() const { return dimension_; }
    void setBounds(const base::RealVectorBounds& b) { bounds_ = b; }
    const base::RealVectorBounds& getBounds() const { return bounds_; }
    Control* allocControl() const { auto* c = new ControlType(); c->values = new double[dimension_](); return c; }
    void freeControl(Control* c) const { delete[] c->as<ControlType>()->values; delete c; }
    void copyControl(Control* d, const Control* s) const {
        for (unsigned i = 0; i < dimension_; ++i) d->as<ControlType>()->values[i] = s->as<ControlType>()->values[i];
    }
private:
    unsigned dimension_;
    base::RealVectorBounds bounds_;
};
using ControlSpacePtr = std::shared_ptr<RealVectorControlSpace>;

class SpaceInformation;
class StatePropagator {
public:
    explicit StatePropagator(SpaceInformation* si) : si_(si) {}
    virtual ~StatePropagator() = default;
    virtual void propagate(const base::State* state, const Control* control, double duration, base::State* result) const = 0;
    virtual bool canPropagateBackward() const { return true; }
protected:
    SpaceInformation* si_;
};
using StatePropagatorPtr = std::shared_ptr<StatePropagator>;

class SpaceInformation : public base::SpaceInformation {
public:
    SpaceInformation(const base::StateSpacePtr& space, ControlSpacePtr cspace) : base::SpaceInformation(space), cspace_(std::move(cspace)) {}
    const ControlSpacePtr& getControlSpace() const { return cspace_; }
    void setStatePropagator(const StatePropagatorPtr& p) { sp_ = p; }
    const StatePropagatorPtr& getStatePropagator() const { return sp_; }
    void setPropagationStepSize(double s) { stepSize_ = s; }
    double getPropagationStepSize() const { return stepSize_; }
    void setMinMaxControlDuration(unsigned lo, unsigned hi) { minSteps_ = lo; maxSteps_ = hi; }
    unsigned getMinControlDuration() const { return minSteps_; }
    unsigned getMaxControlDuration() const { return maxSteps_; }
    Control* allocControl() const { return cspace_->allocControl(); }
    void freeControl(Control* c) const { cspace_->freeControl(c); }
    Control* cloneControl(const Control* c) const { Control* n = allocControl(); cspace_->copyControl(n, c); return n; }

    // OMPL 1.6 control::SpaceInformation::propagateWhileValid (restated)
    unsigned propagateWhileValid(const base::State* state, const Control* control, int steps, base::State* result) const {
        if (steps == 0) { if (result != state) copyState(result, state); return 0; }
        const double signedStep = steps > 0 ? stepSize_ : -stepSize_;
        steps = std::abs(steps);
        sp_->propagate(state, control, signedStep, result);
        if (isValid(result)) {
            base::State* temp1 = result;
            base::State* temp2 = allocState();
            base::State* toDelete = temp2;
            unsigned r = steps;
            for (int i = 1; i < steps; ++i) {
                sp_->propagate(temp1, control, signedStep, temp2);
                if (isValid(temp2)) std::swap(temp1, temp2);
                else { r = i; break; }
            }
            if (result != temp1) copyState(result, temp1);
            freeState(toDelete);
            return r;
        }
        if (result != state) copyState(result, state);
        return 0;
    }
    // OMPL 1.6 propagate(state, control, steps, result vector, alloc=true): no validity checks
    void propagate(const base::State* state, const Control* control, int steps, std::vector<base::State*>& result, bool) const {
        const base::State* cur = state;
        for (int i = 0; i < steps; ++i) {
            base::State* nxt = allocState();
            sp_->propagate(cur, control, stepSize_, nxt);
            result.push_back(nxt);
            cur = nxt;
        }
    }
private:
    ControlSpacePtr cspace_;
    StatePropagatorPtr sp_;
    double stepSize_ = 0.0;
    unsigned minSteps_ = 0, maxSteps_ = 0;
};
using SpaceInformationPtr = std::shared_ptr<SpaceInformation>;

// PathControl with OMPL 1.6's interpolate() (restated)
class PathControl {
public:
    explicit PathControl(const SpaceInformationPtr& si) : si_(si) {}
    void append(const base::State* s) { states_.push_back(si_->cloneState(s)); }
    void append(const base::State* s, const Control* c, double duration) {
        states_.push_back(si_->cloneState(s)); controls_.push_back(si_->cloneControl(c)); durations_.push_back(duration);
    }
    std::vector<base::State*>& getStates() { return states_; }
    base::State* getState(unsigned i) { return states_[i]; }
    std::size_t getStateCount() const { return states_.size(); }
    double getControlDuration(unsigned i) const { return durations_[i]; }
    std::vector<Control*>& getControls() { return controls_; }
    void interpolate() {
        if (states_.size() <= controls_.size()) return;
        std::vector<base::State*> ns; std::vector<Control*> nc; std::vector<double> nd;
        const double res = si_->getPropagationStepSize();
        for (unsigned i = 0; i < controls_.size(); ++i) {
            const int steps = (int)std::floor(0.5 + durations_[i] / res);
            if (steps <= 1) { ns.push_back(states_[i]); nc.push_back(controls_[i]); nd.push_back(durations_[i]); continue; }
            std::vector<base::State*> ist;
            si_->propagate(states_[i], controls_[i], steps, ist, true);
            if (!ist.empty()) { si_->freeState(ist.back()); ist.pop_back(); }
            ns.push_back(states_[i]);
            ns.insert(ns.end(), ist.begin(), ist.end());
            nc.push_back(controls_[i]); nd.push_back(res);
            for (int j = 1; j < steps; ++j) { nc.push_back(si_->cloneControl(controls_[i])); nd.push_back(res); }
        }
        ns.push_back(states_[controls_.size()]);
        states_.swap(ns); controls_.swap(nc); durations_.swap(nd);
    }
private:
    SpaceInformationPtr si_;
    std::vector<base::State*> states_;
    std::vector<Control*> controls_;
    std::vector<double> durations_;
};

}  // namespace control
}  // namespace ompl
#endif  // CCK_HAVE_OMPL
