// TEST INFRASTRUCTURE.  Stand-in for include/utils/MultiRobotProblemDefinition.h of the reference: the real header
// drags in the whole planner stack (Mergers, ConstraintRespectingBSST, ob::ProblemDefinition).  The plan validity
// checkers only use it as a holder of the Instance, the per-robot SpaceInformation and the propagation step size
// (getInstance, getRobotSpaceInformationPtr, getSystemStepSize), which is all this class provides.
#pragma once
#include <vector>

#include "utils/Instance.h"
#include <ompl/control/SpaceInformation.h>

namespace ob = ompl::base;
namespace oc = ompl::control;

OMPL_CLASS_FORWARD(MultiRobotProblemDefinition);
class MultiRobotProblemDefinition {
public:
    MultiRobotProblemDefinition(InstancePtr inst, std::vector<oc::SpaceInformationPtr> sis) : mrmp_instance_(inst), sis_(sis) {}
    const InstancePtr getInstance() const { return mrmp_instance_; }
    const oc::SpaceInformationPtr getRobotSpaceInformationPtr(const int idx) { return sis_[idx]; }
    double getSystemStepSize() { return sis_[0]->getPropagationStepSize(); }
private:
    InstancePtr mrmp_instance_;
    std::vector<oc::SpaceInformationPtr> sis_;
};
