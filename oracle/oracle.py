"""ctypes wrapper over the CPU ORACLE (oracle/cck_oracle.hpp).

TEST INFRASTRUCTURE ONLY: importable from tests/, bench.py's cpu_baseline /
``--impl reference`` leg and ``__graft_entry__.smoke()``; never from the product
package.  Beliefs are AoS numpy arrays ``[n, D + 2*D*D]`` (values | Sigma | Lambda,
matrices row-major), mirroring the reference's per-state objects
(include/Spaces/RealVectorBeliefSpace.h:21-96).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libcck_oracle.so")

SVC_BLACKMORE, SVC_ADAPTIVE, SVC_CHISQUARED = 0, 1, 2
PVC_BLACKMORE, PVC_BOUNDINGBOX, PVC_CHISQUARED, PVC_ADAPTIVE_BLACKMORE, PVC_ADAPTIVE_BOUNDINGBOX, PVC_BLACKMORE2, PVC_CDFGRID = range(7)


def build(force=False):
    """Compile the oracle with g++ (seconds).  Building the checker is not using it."""
    src = [os.path.join(_HERE, f) for f in ("cck_oracle_capi.cpp", "cck_oracle.hpp")]
    if (not force and os.path.exists(_SO)
            and all(os.path.getmtime(_SO) >= os.path.getmtime(s) for s in src)):
        return _SO
    subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        dp, ip, vp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_void_p
        u8p, u64p = C.POINTER(C.c_ubyte), C.POINTER(C.c_uint64)
        L.orc_erf_inv.restype = C.c_double
        L.orc_erf_inv.argtypes = [C.c_double]
        L.orc_chi2_quantile_2dof.restype = C.c_double
        L.orc_chi2_quantile_2dof.argtypes = [C.c_double]
        L.orc_linear_params.argtypes = [C.c_double, dp]
        L.orc_unicycle_params.argtypes = [C.c_double, dp]
        L.orc_world_from_text.restype = vp
        L.orc_world_from_text.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_double]
        L.orc_world_synthetic.restype = vp
        L.orc_world_synthetic.argtypes = [C.c_int, dp, C.c_double, C.c_double, C.c_int, C.c_double]
        L.orc_world_free.argtypes = [vp]
        for f in ("orc_world_n_obstacles", "orc_world_n_robots", "orc_world_map_overrun"):
            getattr(L, f).restype = C.c_int
            getattr(L, f).argtypes = [vp]
        L.orc_world_info.argtypes = [vp, dp]
        L.orc_world_obstacle_centres.argtypes = [vp, dp]
        L.orc_world_robots.argtypes = [vp, dp]
        L.orc_world_robot_model.restype = C.c_int
        L.orc_world_robot_model.argtypes = [vp, C.c_int, C.c_char_p, C.c_int]
        L.orc_svc_create.restype = vp
        L.orc_svc_create.argtypes = [vp, C.c_int, C.c_int, C.c_int, dp, dp, C.c_double]
        L.orc_svc_free.argtypes = [vp]
        L.orc_svc_info.argtypes = [vp, dp]
        L.orc_svc_tables.argtypes = [vp, dp, dp, dp, dp]
        L.orc_svc_override_tables.argtypes = [vp, C.c_int, dp, dp, C.c_double]
        L.orc_svc_set_max_rad.argtypes = [vp, C.c_double]
        L.orc_pvc_set_grid_n.argtypes = [vp, C.c_int]
        L.orc_chisq_pairs_disjoint.argtypes = [C.c_long, dp, dp, dp, dp, dp, dp, C.c_double, C.POINTER(C.c_ubyte)]
        L.orc_cdfgrid_prob.argtypes = [C.c_long, dp, dp, dp, dp, C.c_double, C.c_int, dp]
        L.orc_propagate.argtypes = [C.c_int, C.c_double, C.c_double, C.c_long, dp, dp, dp]
        L.orc_is_valid.argtypes = [vp, C.c_int, C.c_long, dp, u8p, u64p]
        L.orc_rollout.argtypes = [vp, C.c_int, C.c_double, C.c_long, dp, dp, ip, dp, ip, dp, C.c_int, u64p, C.c_int]
        L.orc_propagate_n.argtypes = [C.c_int, C.c_double, dp, dp, C.c_int, dp]
        L.orc_pvc_create.restype = vp
        L.orc_pvc_create.argtypes = [vp, C.c_int, C.c_double]
        L.orc_pvc_free.argtypes = [vp]
        L.orc_pvc_info.argtypes = [vp, dp]
        L.orc_pvc_map_order.argtypes = [vp, ip]
        L.orc_pvc_pair_halfplanes.argtypes = [vp, C.c_int, C.c_int, dp, dp]
        L.orc_validate_plan.restype = C.c_int
        L.orc_validate_plan.argtypes = [vp, C.c_int, C.c_int, ip, dp, ip, C.c_int]
        L.orc_satisfies_constraints.restype = C.c_int
        L.orc_satisfies_constraints.argtypes = [vp, C.c_int, C.c_int, dp, C.c_double, C.c_int, C.c_int, ip, ip, ip, dp]
        L.orc_goal.argtypes = [C.c_int, C.c_long, dp, C.c_double, C.c_double, C.c_double, C.c_double, u8p, dp]
        L.orc_belief_distance.argtypes = [C.c_int, C.c_long, dp, dp, dp]
        L.orc_select_nodes.argtypes = [C.c_int, C.c_long, dp, dp, u8p, C.c_long, dp, C.c_double, ip, dp]
        L.orc_nearest_witness.argtypes = [C.c_int, C.c_long, dp, C.c_long, dp, ip, dp]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def width(dim):
    return dim + 2 * dim * dim


def erf_inv(z):
    return lib().orc_erf_inv(float(z))


def chi2_quantile_2dof(p):
    return lib().orc_chi2_quantile_2dof(float(p))


def linear_params(dt=0.2):
    out = np.zeros(13)
    lib().orc_linear_params(dt, _dp(out))
    return {"dt": out[0], "Q": out[1:5].copy(), "R": out[5:9].copy(), "Acl": out[9:13].copy()}


def unicycle_params(dt=0.2):
    out = np.zeros(76)
    lib().orc_unicycle_params(dt, _dp(out))
    return {"F": out[0:16].copy(), "Acl": out[16:32].copy(), "Q": out[32:48].copy(), "R": out[48:64].copy(),
            "gains": out[64:72].copy(), "acc": out[72:74].copy(), "turn": out[74:76].copy()}


def default_beliefs(n, dim):
    """allocState defaults: Sigma = Lambda = 0.01*I (RealVectorBeliefSpace.cpp:11-12)."""
    b = np.zeros((n, width(dim)))
    eye = 0.01 * np.eye(dim).reshape(-1)
    b[:, dim:dim + dim * dim] = eye
    b[:, dim + dim * dim:] = eye
    return b


class World:
    def __init__(self, handle):
        if not handle:
            raise ValueError("oracle: could not parse map/scen")
        self.h = handle
        L = lib()
        self.n_obstacles = L.orc_world_n_obstacles(self.h)
        self.n_robots = L.orc_world_n_robots(self.h)
        self.map_overrun = bool(L.orc_world_map_overrun(self.h))
        info = np.zeros(4)
        L.orc_world_info(self.h, _dp(info))
        self.x_max, self.y_max, self.p_safe_obs, self.p_safe_agents = info
        self.centres = np.zeros((self.n_obstacles, 2))
        if self.n_obstacles:
            L.orc_world_obstacle_centres(self.h, _dp(self.centres))
        r = np.zeros((self.n_robots, 5))
        if self.n_robots:
            L.orc_world_robots(self.h, _dp(r))
        self.starts, self.goals, self.bounding_rad = r[:, 0:2].copy(), r[:, 2:4].copy(), r[:, 4].copy()
        self.models = []
        for i in range(self.n_robots):
            buf = C.create_string_buffer(128)
            L.orc_world_robot_model(self.h, i, buf, 128)
            self.models.append(buf.value.decode())

    @classmethod
    def from_text(cls, map_text, scen_text, num_agents, p_safe):
        return cls(lib().orc_world_from_text(map_text.encode(), scen_text.encode(), num_agents, p_safe))

    @classmethod
    def from_files(cls, map_path, scen_path, num_agents, p_safe):
        with open(map_path) as f:
            m = f.read()
        with open(scen_path) as f:
            s = f.read()
        return cls.from_text(m, s, num_agents, p_safe)

    @classmethod
    def synthetic(cls, centres, x_max, y_max, k_agents, p_safe):
        c = _f64(centres).reshape(-1, 2)
        return cls(lib().orc_world_synthetic(len(c), _dp(c), x_max, y_max, k_agents, p_safe))

    def bounds(self, dim):
        """State-space bounds as wired by OmplSetUp.cpp:195-200 / :275-289."""
        if dim == 2:
            return np.array([-1.0, -1.0]), np.array([self.x_max, self.y_max])
        return (np.array([-1.0, -1.0, -np.pi, 0.01]), np.array([self.x_max, self.y_max, np.pi, 10.0]))

    def __del__(self):
        try:
            lib().orc_world_free(self.h)
        except Exception:
            pass


class Svc:
    def __init__(self, world, kind=SVC_BLACKMORE, dim=2, robot=0, accep_prob=-1.0, low=None, high=None):
        L = lib()
        self.world, self.kind, self.dim = world, kind, dim
        lo, hi = world.bounds(dim)
        self.low = _f64(low if low is not None else lo)
        self.high = _f64(high if high is not None else hi)
        self.h = L.orc_svc_create(world.h, kind, dim, robot, _dp(self.low), _dp(self.high), accep_prob)
        info = np.zeros(4)
        L.orc_svc_info(self.h, _dp(info))
        self.erf_inv_result, self.p_collision, self.sc, self.max_rad = info
        M = world.n_obstacles
        self.A, self.B = np.zeros((M, 4, 2)), np.zeros((M, 4))
        self.centres, self.rects = np.zeros((M, 2)), np.zeros((M, 4))
        if M:
            L.orc_svc_tables(self.h, _dp(self.A), _dp(self.B), _dp(self.centres), _dp(self.rects))

    def override_tables(self, A, B, erf_inv_result):
        """Test hook (Blackmore only): arbitrary half-plane tables A[M,4,2], B[M,4]."""
        assert self.kind == SVC_BLACKMORE
        self.A, self.B = _f64(A).reshape(-1, 4, 2).copy(), _f64(B).reshape(-1, 4).copy()
        M = len(self.A)
        self.centres, self.rects = np.zeros((M, 2)), np.zeros((M, 4))
        self.erf_inv_result = float(erf_inv_result)
        lib().orc_svc_override_tables(self.h, M, _dp(self.A), _dp(self.B), self.erf_inv_result)

    def set_max_rad(self, r):
        """Test hook: the robot bounding radius of the chi-squared checker (band tests shift it by +-1e-6)."""
        self.max_rad = float(r)
        lib().orc_svc_set_max_rad(self.h, self.max_rad)

    def is_valid(self, beliefs):
        b = _f64(beliefs)
        out = np.zeros(len(b), dtype=np.uint8)
        cnt = np.zeros(3, dtype=np.uint64)
        lib().orc_is_valid(self.h, self.dim, len(b), _dp(b), out.ctypes.data_as(C.POINTER(C.c_ubyte)),
                           cnt.ctypes.data_as(C.POINTER(C.c_uint64)))
        return out.astype(bool), cnt

    def rollout(self, beliefs, controls, steps, dt=0.2, traj_stride=0, nthreads=1):
        """propagateWhileValid over a batch -> (final beliefs, valid_steps, counters[, traj])."""
        b, u = _f64(beliefs), _f64(controls)
        st = np.ascontiguousarray(steps, dtype=np.int32)
        n = len(b)
        out = np.zeros_like(b)
        valid = np.zeros(n, dtype=np.int32)
        cnt = np.zeros(3, dtype=np.uint64)
        traj = np.zeros((n, traj_stride, b.shape[1])) if traj_stride else None
        lib().orc_rollout(self.h, self.dim, dt, n, _dp(b), _dp(u), _ip(st), _dp(out), _ip(valid),
                          _dp(traj) if traj is not None else None, traj_stride,
                          cnt.ctypes.data_as(C.POINTER(C.c_uint64)), nthreads)
        if traj is not None:
            return out, valid, cnt, traj
        return out, valid, cnt

    def __del__(self):
        try:
            lib().orc_svc_free(self.h)
        except Exception:
            pass


def propagate(beliefs, controls, dim, dt=0.2, duration=None):
    b, u = _f64(beliefs), _f64(controls)
    out = np.zeros_like(b)
    lib().orc_propagate(dim, dt, dt if duration is None else duration, len(b), _dp(b), _dp(u), _dp(out))
    return out


def propagate_n(belief, control, steps, dim, dt=0.2):
    b, u = _f64(belief).reshape(-1), _f64(control).reshape(-1)
    out = np.zeros((steps, width(dim)))
    lib().orc_propagate_n(dim, dt, _dp(b), _dp(u), steps, _dp(out))
    return out


class Pvc:
    def __init__(self, world, kind=PVC_BLACKMORE, p_safe_arg=-1.0, grid_n=2):
        L = lib()
        self.world, self.kind, self.k = world, kind, world.n_robots
        self.h = L.orc_pvc_create(world.h, kind, p_safe_arg)
        if kind == PVC_CDFGRID:
            L.orc_pvc_set_grid_n(self.h, int(grid_n))
        info = np.zeros(4)
        L.orc_pvc_info(self.h, _dp(info))
        self.p_coll_agnts, self.p_coll_dist, self.c_pair, self.sc = info
        self.map_order = np.zeros(self.k, dtype=np.int32)
        L.orc_pvc_map_order(self.h, _ip(self.map_order))

    def pair_halfplanes(self, a, b):
        A, B = np.zeros((4, 2)), np.zeros(4)
        lib().orc_pvc_pair_halfplanes(self.h, a, b, _dp(A), _dp(B))
        return A, B

    def validate_plan(self, trajs, dim):
        """trajs: list of [T_r, W] arrays (interpolated states).  -> [(a, b, step), ...]"""
        lens = np.array([len(t) for t in trajs], dtype=np.int32)
        states = _f64(np.concatenate([_f64(t).reshape(-1, width(dim)) for t in trajs], axis=0))
        cap = int(lens.max()) + 2
        out = np.zeros((cap, 3), dtype=np.int32)
        n = lib().orc_validate_plan(self.h, dim, len(trajs), _ip(lens), _dp(states), _ip(out), cap)
        return [tuple(int(x) for x in out[i]) for i in range(n)]

    def satisfies_constraints(self, path, dim, constrained, constraints, step_size=0.2):
        """constraints: list of (constraining_robot, steps[int], states[len(steps), W])."""
        p = _f64(path).reshape(-1, width(dim))
        cr = np.array([c[0] for c in constraints], dtype=np.int32)
        nt = np.array([len(c[1]) for c in constraints], dtype=np.int32)
        st = np.ascontiguousarray(np.concatenate([np.asarray(c[1], dtype=np.int32) for c in constraints]), dtype=np.int32)
        states = _f64(np.concatenate([_f64(c[2]).reshape(-1, width(dim)) for c in constraints], axis=0))
        return bool(lib().orc_satisfies_constraints(self.h, dim, len(p), _dp(p), step_size, constrained, len(constraints),
                                                    _ip(cr), _ip(nt), _ip(st), _dp(states)))

    def __del__(self):
        try:
            lib().orc_pvc_free(self.h)
        except Exception:
            pass


def cdfgrid_prob(mu_a, cov_a, mu_b, cov_b, r_sum, grid_n):
    """CDFGridPVC::isSafe_'s collision mass for n belief pairs (CDFGridPVC.cpp:182-291)."""
    mu_a, cov_a, mu_b, cov_b = (_f64(x) for x in (mu_a, cov_a, mu_b, cov_b))
    out = np.zeros(len(mu_a))
    lib().orc_cdfgrid_prob(len(mu_a), _dp(mu_a), _dp(cov_a), _dp(mu_b), _dp(cov_b), float(r_sum), int(grid_n), _dp(out))
    return out


def chisq_pairs_disjoint(mu_a, cov_a, rad_a, mu_b, cov_b, rad_b, sc):
    """CentralizedChiSquaredBoundarySVC's agent-agent decagon test (AoS mu [n, 2], cov [n, 4], rad [n])."""
    mu_a, cov_a, rad_a, mu_b, cov_b, rad_b = (_f64(x) for x in (mu_a, cov_a, rad_a, mu_b, cov_b, rad_b))
    out = np.zeros(len(mu_a), dtype=np.uint8)
    lib().orc_chisq_pairs_disjoint(len(mu_a), _dp(mu_a), _dp(cov_a), _dp(rad_a), _dp(mu_b), _dp(cov_b), _dp(rad_b), float(sc),
                                   out.ctypes.data_as(C.POINTER(C.c_ubyte)))
    return out.astype(bool)


def goal(beliefs, dim, gx, gy, threshold=2.0, p_safe=0.95):
    b = _f64(beliefs)
    ok = np.zeros(len(b), dtype=np.uint8)
    dist = np.zeros(len(b))
    lib().orc_goal(dim, len(b), _dp(b), gx, gy, threshold, p_safe, ok.ctypes.data_as(C.POINTER(C.c_ubyte)), _dp(dist))
    return ok.astype(bool), dist


def belief_distance(a, b, dim):
    """RealVectorBeliefSpace::distance(a_i, b_i) for AoS records [n, W]."""
    a, b = _f64(a), _f64(b)
    out = np.zeros(len(a))
    lib().orc_belief_distance(dim, len(a), _dp(a), _dp(b), _dp(out))
    return out


def select_nodes(nodes, cost, active, samples, dim, radius=0.2):
    """selectNode for each sample against the AoS node set -> (index, distance)."""
    nodes, cost, samples = _f64(nodes), _f64(cost), _f64(samples)
    act = np.ascontiguousarray(active, dtype=np.uint8)
    idx = np.zeros(len(samples), dtype=np.int32)
    dist = np.zeros(len(samples))
    lib().orc_select_nodes(dim, len(nodes), _dp(nodes), _dp(cost), act.ctypes.data_as(C.POINTER(C.c_ubyte)), len(samples), _dp(samples),
                           radius, _ip(idx), _dp(dist))
    return idx, dist


def nearest_witness(witnesses, queries, dim):
    witnesses, queries = _f64(witnesses), _f64(queries)
    idx = np.zeros(len(queries), dtype=np.int32)
    dist = np.zeros(len(queries))
    lib().orc_nearest_witness(dim, len(witnesses), _dp(witnesses), len(queries), _dp(queries), _ip(idx), _dp(dist))
    return idx, dist
