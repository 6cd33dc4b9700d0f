"""ctypes wrapper over oracle/_ref/libcck_ref.so: the REFERENCE'S OWN SOURCES for the headline path
(Instance/Robot/Obstacle, R2_UncertainLinearStatePropagator, PCCBlackmoreSVC) compiled from /root/reference
against the stand-in headers of oracle/shim/.  TEST INFRASTRUCTURE: used to pin the oracle and to generate
the golden vectors tests/golden/ref_*.npz (oracle/gen_ref_golden.py).  Available only where /root/reference
is mounted (this container); the prebuilt .so travels to the GPU box, the sources do not."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libcck_ref.so")
REFERENCE = "/root/reference"
SOURCES = ["src/utils/common.cpp", "src/utils/Instance.cpp", "src/utils/Conflict.cpp", "src/StatePropogators/UncertainLinearSP.cpp",
           "src/StateValidityCheckers/PCCBlackmoreSVC.cpp", "src/StateValidityCheckers/AdaptiveRiskBlackmoreSVC.cpp",
           "src/PlanValidityCheckers/AdaptiveRiskBlackmorePVC.cpp", "src/PlanValidityCheckers/AdaptiveRiskBoundingBoxPVC.cpp", "src/Constraints/BeliefConstraint.cpp", "src/PlanValidityCheckers/BeliefPVC.cpp",
           "src/PlanValidityCheckers/MinkowskiSumBlackmorePVC.cpp", "src/PlanValidityCheckers/BoundingBoxBlackmorePVC.cpp",
           "src/Spaces/RealVectorBeliefSpace.cpp", "src/StatePropogators/UncertainUnicycleSP.cpp", "src/Goals/BeliefSpaceGoals.cpp",
           "src/StateValidityCheckers/ChiSquaredBoundarySVC.cpp", "src/PlanValidityCheckers/ChiSquaredBoundaryPVC.cpp",
           "src/PlanValidityCheckers/CDFGridPVC.cpp", "src/PlanValidityCheckers/Blackmore2PVC.cpp", "src/StatePropogators/CentralizedUncertainLinearSP.cpp"]


def available():
    return os.path.exists(_SO)


def build(force=False):
    """Compile the reference sources where they lie (Release flags of its CMakeLists: -O3 -DNDEBUG, C++17,
    baseline x86-64).  Returns the library path, or None when /root/reference is not mounted."""
    if not os.path.isdir(REFERENCE):
        return _SO if os.path.exists(_SO) else None
    srcs = [os.path.join(REFERENCE, s) for s in SOURCES]
    deps = srcs + [os.path.join(_HERE, "ref_capi.cpp"), os.path.join(_HERE, "ref_capi2.cpp")] + [os.path.join(_HERE, "shim", f) for f in ("shim_eigen.hpp", "shim_boost.hpp", "shim_ompl.hpp", "utils/MultiRobotProblemDefinition.h")]
    if not force and os.path.exists(_SO) and all(os.path.getmtime(_SO) >= os.path.getmtime(d) for d in deps):
        return _SO
    subprocess.check_call(["make", "-C", _HERE, "-s", "ref"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(_SO)
        dp, ip, vp, u8p = C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_void_p, C.POINTER(C.c_ubyte)
        L.ref_world_create.restype = vp
        L.ref_world_create.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_double]
        L.ref_world_free.argtypes = [vp]
        L.ref_world_info.argtypes = [vp, dp]
        L.ref_world_obstacles.argtypes = [vp, dp]
        L.ref_world_robots.argtypes = [vp, dp]
        L.ref_svc_create.restype = vp
        L.ref_svc_create.argtypes = [vp, C.c_int, dp, dp, C.c_double]
        L.ref_svc_create_kind.restype = vp
        L.ref_svc_create_kind.argtypes = [vp, C.c_int, C.c_int, dp, dp, C.c_double]
        L.ref_svc_free.argtypes = [vp]
        L.ref_svc_tables.argtypes = [vp, dp, dp, dp]
        L.ref_is_valid.argtypes = [vp, C.c_long, dp, u8p]
        L.ref_propagate.argtypes = [vp, C.c_long, dp, dp, dp]
        L.ref_rollout.argtypes = [vp, C.c_long, dp, dp, ip, dp, ip]
        L.ref_pvc_create.restype = vp
        L.ref_pvc_create.argtypes = [vp, C.c_int, C.c_double]
        L.ref_pvc_free.argtypes = [vp]
        L.ref_pvc_info.argtypes = [vp, dp]
        L.ref_pvc_pair_halfplanes.argtypes = [vp, C.c_int, C.c_int, dp, dp]
        L.ref_validate_plan.restype = C.c_int
        L.ref_validate_plan.argtypes = [vp, C.c_int, ip, dp, ip, C.c_int]
        L.ref_create_constraint.restype = C.c_int
        L.ref_create_constraint.argtypes = [vp, C.c_int, ip, dp, C.c_int, ip, dp, dp, C.c_int]
        L.ref_satisfies_constraints.restype = C.c_int
        L.ref_satisfies_constraints.argtypes = [vp, C.c_int, dp, C.c_int, C.c_int, ip, ip, ip, dp]
        u8 = u8p
        L.ref2_svc_create.restype = vp
        L.ref2_svc_create.argtypes = [vp, C.c_int, C.c_int, C.c_int, dp, dp, C.c_double, C.c_double]
        L.ref2_svc_free.argtypes = [vp]
        L.ref2_uni_params.argtypes = [vp, dp, dp, dp, dp]
        L.ref2_chisq_info.argtypes = [vp, dp]
        L.ref2_propagate.argtypes = [vp, C.c_long, dp, dp, C.c_double, dp]
        L.ref2_is_valid.argtypes = [vp, C.c_long, dp, u8]
        L.ref2_rollout.argtypes = [vp, C.c_long, dp, dp, ip, dp, ip]
        L.ref2_distance.argtypes = [C.c_int, C.c_long, dp, dp, dp]
        L.ref2_goal.argtypes = [C.c_int, C.c_long, dp, C.c_double, C.c_double, C.c_double, C.c_double, u8, dp]
        L.ref2_pvc_create.restype = vp
        L.ref2_pvc_create.argtypes = [vp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int]
        L.ref2_pvc_free.argtypes = [vp]
        L.ref2_pvc_info.argtypes = [vp, dp]
        L.ref2_validate_plan.restype = C.c_int
        L.ref2_validate_plan.argtypes = [vp, C.c_int, ip, dp, ip, C.c_int]
        L.ref2_satisfies_constraints.restype = C.c_int
        L.ref2_satisfies_constraints.argtypes = [vp, C.c_int, dp, C.c_int, C.c_int, ip, ip, ip, dp]
        L.ref2_independent_check.argtypes = [vp, C.c_long, dp, dp, u8]
        L.ref2_central_propagate.argtypes = [C.c_int, C.c_long, dp, dp, C.c_double, dp]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class World:
    def __init__(self, map_path, scen_path, k, p_safe):
        self.h = lib().ref_world_create(map_path.encode(), scen_path.encode(), k, p_safe)
        info = np.zeros(6)
        lib().ref_world_info(self.h, _dp(info))
        self.n_obstacles, self.n_robots = int(info[0]), int(info[1])
        self.x_max, self.y_max, self.p_safe_obs, self.p_safe_agents = info[2:6]
        self.centres = np.zeros((self.n_obstacles, 2))
        if self.n_obstacles:
            lib().ref_world_obstacles(self.h, _dp(self.centres))
        r = np.zeros((self.n_robots, 5))
        lib().ref_world_robots(self.h, _dp(r))
        self.starts, self.goals, self.bounding_rad = r[:, 0:2].copy(), r[:, 2:4].copy(), r[:, 4].copy()


class Svc:
    """PCCBlackmoreSVC + R2_UncertainLinearStatePropagator wired as OmplSetUp.cpp:183-228 does."""

    def __init__(self, world, robot=0, step=0.2, kind=0):
        """kind 0: PCCBlackmoreSVC, kind 1: AdaptiveRiskBlackmoreSVC (its erf_inv is the stand-in's)."""
        self.world = world
        self.low = np.array([-1.0, -1.0])
        self.high = np.array([world.x_max, world.y_max], dtype=np.float64)
        self.h = lib().ref_svc_create_kind(world.h, kind, robot, _dp(self.low), _dp(self.high), step)
        M = world.n_obstacles
        self.A, self.B = np.zeros((M, 4, 2)), np.zeros((M, 4))
        info = np.zeros(3)
        lib().ref_svc_tables(self.h, _dp(self.A), _dp(self.B), _dp(info))
        self.erf_inv_result, self.p_collision = info[0], info[1]

    def is_valid(self, bel):
        b = np.ascontiguousarray(bel, dtype=np.float64)
        out = np.zeros(len(b), dtype=np.uint8)
        lib().ref_is_valid(self.h, len(b), _dp(b), out.ctypes.data_as(C.POINTER(C.c_ubyte)))
        return out.astype(bool)

    def propagate(self, bel, ctrl):
        b, u = np.ascontiguousarray(bel, dtype=np.float64), np.ascontiguousarray(ctrl, dtype=np.float64)
        out = np.zeros_like(b)
        lib().ref_propagate(self.h, len(b), _dp(b), _dp(u), _dp(out))
        return out

    def rollout(self, bel, ctrl, steps):
        b, u = np.ascontiguousarray(bel, dtype=np.float64), np.ascontiguousarray(ctrl, dtype=np.float64)
        st = np.ascontiguousarray(steps, dtype=np.int32)
        out = np.zeros_like(b)
        valid = np.zeros(len(b), dtype=np.int32)
        lib().ref_rollout(self.h, len(b), _dp(b), _dp(u), st.ctypes.data_as(C.POINTER(C.c_int)), _dp(out), valid.ctypes.data_as(C.POINTER(C.c_int)))
        return out, valid


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


class Pvc:
    """MinkowskiSumBlackmorePVC (kind 0) / BoundingBoxBlackmorePVC (1) / AdaptiveRiskBlackmorePVC (2) /
    AdaptiveRiskBoundingBoxPVC (3) as demos/main.cpp:100-125 wires them
    (p_safe = Instance::getPsafeAgents()).  Paths are interpolated trajectories [T_r, 10]."""

    def __init__(self, world, kind=0, step=0.2):
        self.world, self.kind, self.k = world, kind, world.n_robots
        self.h = lib().ref_pvc_create(world.h, kind, step)
        info = np.zeros(3)
        lib().ref_pvc_info(self.h, _dp(info))
        self.p_coll_agnts, self.p_coll_dist, self.c_pair = info

    def pair_halfplanes(self, a, b):
        A, B = np.zeros((4, 2)), np.zeros(4)
        lib().ref_pvc_pair_halfplanes(self.h, a, b, _dp(A), _dp(B))
        return A, B

    def _pack(self, trajs):
        lens = np.array([len(t) for t in trajs], dtype=np.int32)
        states = np.ascontiguousarray(np.concatenate([np.asarray(t, dtype=np.float64).reshape(-1, 10) for t in trajs], axis=0))
        return lens, states

    def validate_plan(self, trajs):
        lens, states = self._pack(trajs)
        cap = int(lens.max()) + 2
        out = np.zeros((cap, 3), dtype=np.int32)
        n = lib().ref_validate_plan(self.h, len(trajs), _ip(lens), _dp(states), _ip(out), cap)
        return [tuple(int(x) for x in out[i]) for i in range(n)]

    def create_constraint(self, trajs, robot):
        """createConstraint(plan, validatePlan(plan), robot) -> (constrained, constraining, times, states) or None."""
        lens, states = self._pack(trajs)
        cap = int(lens.max()) + 2
        info = np.zeros(2, dtype=np.int32)
        times, cst = np.zeros(cap), np.zeros((cap, 10))
        n = lib().ref_create_constraint(self.h, len(trajs), _ip(lens), _dp(states), robot, _ip(info), _dp(times), _dp(cst), cap)
        if n == 0:
            return None
        return int(info[0]), int(info[1]), times[:n].copy(), cst[:n].copy()

    def satisfies_constraints(self, path, constrained, constraints):
        """constraints: list of (constraining_robot, steps[int], states[len(steps), 10])."""
        p = np.ascontiguousarray(np.asarray(path, dtype=np.float64).reshape(-1, 10))
        cr = np.array([c[0] for c in constraints], dtype=np.int32)
        nt = np.array([len(c[1]) for c in constraints], dtype=np.int32)
        st = np.ascontiguousarray(np.concatenate([np.asarray(c[1], dtype=np.int32) for c in constraints]), dtype=np.int32)
        states = np.ascontiguousarray(np.concatenate([np.asarray(c[2], dtype=np.float64).reshape(-1, 10) for c in constraints], axis=0))
        return bool(lib().ref_satisfies_constraints(self.h, len(p), _dp(p), constrained, len(constraints), _ip(cr), _ip(nt), _ip(st), _dp(states)))


# ---------------------------------------------------------------------------
# round 2: unicycle propagator, belief-space distance, goal, chi-squared / CDF-grid / Blackmore2 checkers
# (oracle/ref_capi2.cpp over the reference's own UncertainUnicycleSP.cpp, RealVectorBeliefSpace.cpp, BeliefSpaceGoals.cpp,
#  ChiSquaredBoundarySVC.cpp, ChiSquaredBoundaryPVC.cpp, CDFGridPVC.cpp, Blackmore2PVC.cpp)
# ---------------------------------------------------------------------------
def _u8(a):
    return a.ctypes.data_as(C.POINTER(C.c_ubyte))


SVC2_BLACKMORE, SVC2_ADAPTIVE, SVC2_CHISQUARED = 0, 1, 2
PVC2_MINKOWSKI, PVC2_BBOX, PVC2_ADAPT_MINKOWSKI, PVC2_ADAPT_BBOX, PVC2_CHISQUARED, PVC2_CDFGRID, PVC2_BLACKMORE2 = range(7)


class Svc2:
    """One robot's SpaceInformation as OmplSetUp.cpp:183-353 wires it: model 'linear' (dim 2) or 'unicycle' (dim 4),
    svc_kind 0 PCCBlackmoreSVC / 1 AdaptiveRiskBlackmoreSVC / 2 ChiSquaredBoundarySVC (hard-coded 0.9)."""

    def __init__(self, world, model="linear", svc_kind=0, robot=0, step=0.2, accep_prob=None):
        self.world, self.dim = world, (4 if model == "unicycle" else 2)
        self.W = self.dim + 2 * self.dim * self.dim
        if self.dim == 4:      # OmplSetUp.cpp:275-289: control bounds == state bounds
            self.low = np.array([-1.0, -1.0, -np.pi, 0.01])
            self.high = np.array([world.x_max, world.y_max, np.pi, 10.0], dtype=np.float64)
        else:
            self.low = np.array([-1.0, -1.0])
            self.high = np.array([world.x_max, world.y_max], dtype=np.float64)
        if accep_prob is None:
            accep_prob = 0.9 if svc_kind == SVC2_CHISQUARED else world.p_safe_obs
        self.step = step
        self.h = lib().ref2_svc_create(world.h, svc_kind, robot, 1 if self.dim == 4 else 0, _dp(self.low), _dp(self.high), step, accep_prob)

    def uni_params(self):
        F, A, Q, R = (np.zeros((4, 4)) for _ in range(4))
        lib().ref2_uni_params(self.h, _dp(F), _dp(A), _dp(Q), _dp(R))
        return F, A, Q, R

    def chisq_info(self):
        o = np.zeros(3)
        lib().ref2_chisq_info(self.h, _dp(o))
        return {"sc": o[0], "max_rad": o[1], "n_obstacles": int(o[2])}

    def propagate(self, bel, ctrl, duration=None):
        b, u = np.ascontiguousarray(bel, dtype=np.float64), np.ascontiguousarray(ctrl, dtype=np.float64)
        out = np.zeros_like(b)
        lib().ref2_propagate(self.h, len(b), _dp(b), _dp(u), self.step if duration is None else duration, _dp(out))
        return out

    def is_valid(self, bel):
        b = np.ascontiguousarray(bel, dtype=np.float64)
        out = np.zeros(len(b), dtype=np.uint8)
        lib().ref2_is_valid(self.h, len(b), _dp(b), _u8(out))
        return out.astype(bool)

    def rollout(self, bel, ctrl, steps):
        b, u = np.ascontiguousarray(bel, dtype=np.float64), np.ascontiguousarray(ctrl, dtype=np.float64)
        st = np.ascontiguousarray(steps, dtype=np.int32)
        out = np.zeros_like(b)
        valid = np.zeros(len(b), dtype=np.int32)
        lib().ref2_rollout(self.h, len(b), _dp(b), _dp(u), _ip(st), _dp(out), _ip(valid))
        return out, valid


def distance(dim, a, b):
    """RealVectorBeliefSpace::distance(a_i, b_i) (src/Spaces/RealVectorBeliefSpace.cpp:16-62)."""
    a, b = np.ascontiguousarray(a, dtype=np.float64), np.ascontiguousarray(b, dtype=np.float64)
    out = np.zeros(len(a))
    lib().ref2_distance(dim, len(a), _dp(a), _dp(b), _dp(out))
    return out


def goal(dim, bel, gx, gy, toll=2.0, p_safe=0.95):
    """ChanceConstrainedGoal(si, goal, toll, p_safe)::isSatisfied(state, &distance) (OmplSetUp.cpp:239)."""
    b = np.ascontiguousarray(bel, dtype=np.float64)
    ok, dist = np.zeros(len(b), dtype=np.uint8), np.zeros(len(b))
    lib().ref2_goal(dim, len(b), _dp(b), gx, gy, toll, p_safe, _u8(ok), _dp(dist))
    return ok.astype(bool), dist


class Pvc2:
    """Any of the reference's belief plan checkers for state dimension 2 or 4 (kinds PVC2_*), constructed with the
    arguments demos/main.cpp:105-138 passes (ChiSquared: the hard-coded 0.9)."""

    def __init__(self, world, kind, dim=2, step=0.2, p_arg=None, n_steps=2):
        self.world, self.kind, self.dim, self.k = world, kind, dim, world.n_robots
        self.W = dim + 2 * dim * dim
        if p_arg is None:
            p_arg = 0.9 if kind == PVC2_CHISQUARED else world.p_safe_agents
        self.h = lib().ref2_pvc_create(world.h, kind, dim, step, p_arg, n_steps)
        o = np.zeros(3)
        lib().ref2_pvc_info(self.h, _dp(o))
        self.p_coll_agnts, self.sc, self.p_coll_dist = o

    def validate_plan(self, trajs):
        lens = np.array([len(t) for t in trajs], dtype=np.int32)
        states = np.ascontiguousarray(np.concatenate([np.asarray(t, dtype=np.float64).reshape(-1, self.W) for t in trajs], axis=0))
        cap = int(lens.max()) + 2
        out = np.zeros((cap, 3), dtype=np.int32)
        n = lib().ref2_validate_plan(self.h, len(trajs), _ip(lens), _dp(states), _ip(out), cap)
        return [tuple(int(x) for x in out[i]) for i in range(n)]

    def satisfies_constraints(self, path, constrained, constraints):
        p = np.ascontiguousarray(np.asarray(path, dtype=np.float64).reshape(-1, self.W))
        cr = np.array([c[0] for c in constraints], dtype=np.int32)
        nt = np.array([len(c[1]) for c in constraints], dtype=np.int32)
        st = np.ascontiguousarray(np.concatenate([np.asarray(c[1], dtype=np.int32) for c in constraints]), dtype=np.int32)
        states = np.ascontiguousarray(np.concatenate([np.asarray(c[2], dtype=np.float64).reshape(-1, self.W) for c in constraints], axis=0))
        return bool(lib().ref2_satisfies_constraints(self.h, len(p), _dp(p), constrained, len(constraints), _ip(cr), _ip(nt), _ip(st), _dp(states)))

    def independent_check(self, a, b):
        a, b = np.ascontiguousarray(a, dtype=np.float64), np.ascontiguousarray(b, dtype=np.float64)
        out = np.zeros(len(a), dtype=np.uint8)
        lib().ref2_independent_check(self.h, len(a), _dp(a), _dp(b), _u8(out))
        return out.astype(bool)


def central_propagate(k, states, controls, step=0.5):
    """CentralizedUncertainLinearStatePropagator::propagate (CentralizedUncertainLinearSP.cpp:42-70) on n composite states of
    k agents: records [2k values | Sigma (2k x 2k) | Lambda (2k x 2k)], controls [n, 2k].  Entries outside the agents'
    diagonal blocks are left at the allocState default of the result state."""
    s, u = np.ascontiguousarray(states, dtype=np.float64), np.ascontiguousarray(controls, dtype=np.float64)
    out = np.zeros_like(s)
    lib().ref2_central_propagate(k, len(s), _dp(s), _dp(u), step, _dp(out))
    return out
