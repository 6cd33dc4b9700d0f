// shim_ompl.hpp — TEST INFRASTRUCTURE.  The OMPL 1.6 types the reference's belief space, linear propagator and
// PCCBlackmoreSVC are written against (State::as<T>, RealVectorStateSpace::StateType, control::StatePropagator,
// base::StateValidityChecker, SpaceInformation::getPropagationStepSize / satisfiesBounds with OMPL's
// RealVectorBounds test: v - eps > high || v + eps < low, eps = DBL_EPSILON), no planning machinery.
#pragma once
#include <algorithm>
#include <cfloat>
#include <map>
#include <unordered_map>
#include <cmath>
#include <stdexcept>
#include <iostream>
#include <limits>
#include <memory>
#include <string>
#include <vector>

#define OMPL_CLASS_FORWARD(C) class C; typedef std::shared_ptr<C> C##Ptr
#define OMPL_INFORM(...) ((void)0)
#define OMPL_WARN(...) ((void)0)
#define OMPL_ERROR(...) ((void)0)
#define OMPL_DEBUG(...) ((void)0)

namespace ompl {
namespace base {

class State {
public:
    template <class T> const T* as() const { return static_cast<const T*>(this); }
    template <class T> T* as() { return static_cast<T*>(this); }
protected:
    State() = default;
    ~State() = default;
};

enum StateSpaceType { STATE_SPACE_UNKNOWN = 0, STATE_SPACE_REAL_VECTOR = 1 };

struct RealVectorBounds {
    std::vector<double> low, high;
    explicit RealVectorBounds(unsigned dim = 0) : low(dim, 0.0), high(dim, 0.0) {}
};

class StateSpace {
public:
    virtual ~StateSpace() = default;
    const std::string& getName() const { return name_; }
    void setName(const std::string& n) { name_ = n; }
    virtual State* allocState() const = 0;
    virtual void freeState(State*) const = 0;
    virtual void copyState(State*, const State*) const = 0;
    virtual double distance(const State*, const State*) const = 0;
    virtual void printState(const State*, std::ostream&) const {}
    virtual bool satisfiesBounds(const State*) const = 0;
    virtual unsigned getDimension() const = 0;
protected:
    int type_ = STATE_SPACE_UNKNOWN;
    std::string name_ = "Space";
};

class RealVectorStateSpace : public StateSpace {
public:
    class StateType : public State {
    public:
        StateType() = default;
        double* values = nullptr;
    };
    explicit RealVectorStateSpace(unsigned dim = 0) : dimension_(dim), bounds_(dim) {}
    unsigned getDimension() const override { return dimension_; }
    void setBounds(const RealVectorBounds& b) { bounds_ = b; }
    const RealVectorBounds& getBounds() const { return bounds_; }
    State* allocState() const override { auto* s = new StateType(); s->values = new double[dimension_]; return s; }
    void freeState(State* s) const override { auto* r = static_cast<StateType*>(s); delete[] r->values; delete r; }
    void copyState(State* d, const State* s) const override {
        for (unsigned i = 0; i < dimension_; ++i) static_cast<StateType*>(d)->values[i] = static_cast<const StateType*>(s)->values[i];
    }
    double distance(const State*, const State*) const override { return 0; }
    void printState(const State*, std::ostream&) const override {}
    // RealVectorStateSpace::satisfiesBounds (OMPL 1.6, src/ompl/base/spaces/src/RealVectorStateSpace.cpp)
    bool satisfiesBounds(const State* state) const override {
        const auto* r = static_cast<const StateType*>(state);
        for (unsigned i = 0; i < dimension_; ++i)
            if (r->values[i] - std::numeric_limits<double>::epsilon() > bounds_.high[i] ||
                r->values[i] + std::numeric_limits<double>::epsilon() < bounds_.low[i])
                return false;
        return true;
    }
protected:
    unsigned dimension_;
    RealVectorBounds bounds_;
};

using StateSpacePtr = std::shared_ptr<StateSpace>;

class SpaceInformation {
public:
    explicit SpaceInformation(StateSpacePtr s) : space_(std::move(s)) {}
    virtual ~SpaceInformation() = default;
    const StateSpacePtr& getStateSpace() const { return space_; }
    bool satisfiesBounds(const State* s) const { return space_->satisfiesBounds(s); }
    State* allocState() const { return space_->allocState(); }
    void freeState(State* s) const { space_->freeState(s); }
    void copyState(State* d, const State* s) const { space_->copyState(d, s); }
    State* cloneState(const State* s) const { State* c = space_->allocState(); space_->copyState(c, s); return c; }
protected:
    StateSpacePtr space_;
};
using SpaceInformationPtr = std::shared_ptr<SpaceInformation>;

// base::Goal: what ChanceConstrainedGoal overrides (ompl/base/Goal.h)
class Goal {
public:
    explicit Goal(const SpaceInformationPtr& si) : si_(si) {}
    virtual ~Goal() = default;
    virtual bool isSatisfied(const State*) const = 0;
    virtual bool isSatisfied(const State* st, double* distance) const { if (distance) *distance = std::numeric_limits<double>::max(); return isSatisfied(st); }
protected:
    SpaceInformationPtr si_;
};

class StateValidityChecker {
public:
    explicit StateValidityChecker(const SpaceInformationPtr& si) : si_(si.get()) {}
    explicit StateValidityChecker(SpaceInformation* si) : si_(si) {}
    virtual ~StateValidityChecker() = default;
    virtual bool isValid(const State*) const = 0;
protected:
    SpaceInformation* si_;
};

}  // namespace base

namespace control {

class Control {
public:
    template <class T> const T* as() const { return static_cast<const T*>(this); }
    template <class T> T* as() { return static_cast<T*>(this); }
protected:
    Control() = default;
    ~Control() = default;
};

class RealVectorControlSpace {
public:
    class ControlType : public Control {
    public:
        double* values = nullptr;
    };
};

class SpaceInformation : public base::SpaceInformation {
public:
    explicit SpaceInformation(base::StateSpacePtr s) : base::SpaceInformation(std::move(s)) {}
    void setPropagationStepSize(double s) { step_ = s; }
    double getPropagationStepSize() const { return step_; }
private:
    double step_ = 0.05;
};
using SpaceInformationPtr = std::shared_ptr<SpaceInformation>;

class StatePropagator {
public:
    explicit StatePropagator(const SpaceInformationPtr& si) : si_(si.get()) {}
    virtual ~StatePropagator() = default;
    virtual void propagate(const base::State*, const Control*, double, base::State*) const = 0;
    virtual bool canPropagateBackward() const { return true; }
protected:
    SpaceInformation* si_;
};

// PathControl: states / control durations with OMPL's value semantics (copies own their states).  Only paths that
// are ALREADY interpolated (every control lasts one propagation step) are supported: interpolate() then has
// nothing to do, exactly as in OMPL (steps <= 1 keeps state, control and duration).
class PathControl {
public:
    explicit PathControl(const base::SpaceInformationPtr& si) : si_(si) {}
    PathControl(const PathControl& o) : si_(o.si_), durations_(o.durations_) { for (auto* s : o.states_) states_.push_back(si_->cloneState(s)); }
    PathControl& operator=(const PathControl& o) {
        if (this != &o) { clear(); si_ = o.si_; durations_ = o.durations_; for (auto* s : o.states_) states_.push_back(si_->cloneState(s)); }
        return *this;
    }
    ~PathControl() { clear(); }
    void append(const base::State* s) { states_.push_back(si_->cloneState(s)); }
    void append(const base::State* s, double duration) { states_.push_back(si_->cloneState(s)); durations_.push_back(duration); }
    void interpolate() {
        const double res = static_cast<const SpaceInformation*>(si_.get())->getPropagationStepSize();
        for (double d : durations_)
            if ((int)std::floor(0.5 + d / res) > 1) throw std::runtime_error("shim PathControl: feed interpolated paths");
    }
    std::size_t getStateCount() const { return states_.size(); }
    std::size_t getControlCount() const { return durations_.size(); }
    base::State* getState(unsigned i) { return states_[i]; }
    const base::State* getState(unsigned i) const { return states_[i]; }
    std::vector<base::State*>& getStates() { return states_; }
    double getControlDuration(unsigned i) const { return durations_[i]; }
private:
    void clear() { for (auto* s : states_) si_->freeState(s); states_.clear(); }
    base::SpaceInformationPtr si_;
    std::vector<base::State*> states_;
    std::vector<double> durations_;
};

}  // namespace control
}  // namespace ompl
