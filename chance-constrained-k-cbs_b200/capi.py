"""ctypes binding of libcck_b200.so (include/cck_b200.h) + a thin numpy/torch-friendly mirror
of the reference's plugin objects for this path.

This module is plumbing: all arithmetic happens in the CUDA library.  It never imports the
CPU oracle; when the shared library is missing or no sm_100 GPU is present every compute call
raises (there is no fallback).
"""
import ctypes as C
import weakref
import os
import subprocess

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
_SO = os.environ.get("CCK_SO") or os.path.join(_PKG, "libcck_b200.so")   # CCK_SO: alternative build (kernel tuning experiments)
_ROOT = os.path.dirname(_PKG)

OK, ERR_INVALID, ERR_CUDA, ERR_NOMEM, ERR_NODEVICE = 0, 1, 2, 3, 4
MODEL_LINEAR2D, MODEL_UNICYCLE = 0, 1
SVC_BLACKMORE, SVC_ADAPTIVE, SVC_CHISQUARED = 0, 1, 2
PVC_BLACKMORE, PVC_BOUNDINGBOX, PVC_CHISQUARED, PVC_ADAPTIVE_BLACKMORE, PVC_ADAPTIVE_BOUNDINGBOX, PVC_BLACKMORE2, PVC_CDFGRID = range(7)
FLAG_FULL, FLAG_CONS_OK, FLAG_GOAL, FLAG_COV_TABLE = 1, 2, 4, 8


class CckError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"cck_b200 error {code}: {msg}")
        self.code = code


class ModelParams(C.Structure):
    _fields_ = [("model", C.c_int32), ("dim", C.c_int32), ("dt", C.c_double),
                ("lin_Q", C.c_double * 4), ("lin_R", C.c_double * 4), ("lin_Acl", C.c_double * 4),
                ("uni_F", C.c_double * 16), ("uni_Acl", C.c_double * 16), ("uni_Q", C.c_double * 16), ("uni_R", C.c_double * 16),
                ("uni_gains", C.c_double * 8), ("uni_acc", C.c_double * 2), ("uni_turn", C.c_double * 2)]


class SvcParams(C.Structure):
    _fields_ = [("kind", C.c_int32), ("dim", C.c_int32), ("low", C.c_double * 4), ("high", C.c_double * 4),
                ("erf_inv_result", C.c_double), ("p_collision", C.c_double), ("sc", C.c_double), ("max_rad", C.c_double)]


class PvcParams(C.Structure):
    _fields_ = [("kind", C.c_int32), ("k", C.c_int32), ("dim", C.c_int32), ("grid_n", C.c_int32),
                ("c_pair", C.c_double), ("p_coll_agnts", C.c_double), ("sc", C.c_double),
                ("radii", C.POINTER(C.c_double)), ("pair_A", C.POINTER(C.c_double)), ("pair_B", C.POINTER(C.c_double)),
                ("map_order", C.POINTER(C.c_int32))]


class ConflictRun(C.Structure):
    _fields_ = [("a", C.c_int32), ("b", C.c_int32), ("first_step", C.c_int32), ("run_len", C.c_int32)]


class GoalParams(C.Structure):
    _fields_ = [("gx", C.c_double), ("gy", C.c_double), ("threshold", C.c_double), ("sc", C.c_double)]


# every symbol include/cck_b200.h declares (tests check the library exports all of them)
SYMBOLS = [
    "cck_ctx_create", "cck_ctx_destroy", "cck_last_error", "cck_ctx_synchronize", "cck_num_sms", "cck_launch_count",
    "cck_model_params_default", "cck_set_model", "cck_set_obstacles", "cck_set_pvc",
    "cck_erf_inv", "cck_chi2_quantile_2dof", "cck_split_risk", "cck_rect_robot_bounding_radius",
    "cck_build_obstacle_tables", "cck_build_pair_halfplanes",
    "cck_propagate", "cck_propagate_dev", "cck_is_valid", "cck_is_valid_dev",
    "cck_propagate_while_valid", "cck_propagate_while_valid_dev",
    "cck_validate_plan", "cck_validate_plan_dev", "cck_validate_plan_range", "cck_validate_plan_range_dev",
    "cck_set_constraints", "cck_satisfies_constraints", "cck_satisfies_constraints_dev", "cck_build_clearance_field", "cck_measure_fp64_rates", "cck_set_cov_table", "cck_chisq_pairs_disjoint", "cck_host_pair_safe", "cck_host_chisq_pair_disjoint",
    "cck_expand_batch", "cck_expand_batch_dev", "cck_rollout_summary_dev", "cck_select_winner_dev", "cck_measure_fp64_peak",
    "cck_tree_create", "cck_tree_destroy", "cck_tree_clear", "cck_tree_size", "cck_tree_append", "cck_tree_update",
    "cck_tree_select", "cck_tree_nearest", "cck_tree_witness_batch", "cck_belief_distance", "cck_belief_distance_lower", "cck_belief_distance_cols", "cck_belief_distance_cols_near", "cck_build_grid_table",
]


def build(force=False, verbose=False):
    """Compile libcck_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
    import glob
    src = sorted(glob.glob(os.path.join(_PKG, "csrc", "*.cu")) + glob.glob(os.path.join(_PKG, "csrc", "*.cuh")))
    src.append(os.path.join(_ROOT, "include", "cck_b200.h"))
    if not force and os.path.exists(_SO) and all(os.path.getmtime(_SO) >= os.path.getmtime(s) for s in src):
        return _SO
    out = subprocess.run(["make", "-C", os.path.join(_PKG, "csrc"), "-B"], capture_output=True, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout[-4000:], out.stderr[-4000:])
    if out.returncode != 0:
        raise RuntimeError("nvcc build of libcck_b200.so failed")
    return _SO


_lib = None


def lib():
    """Load the shared library; raises loudly if it is missing (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        raise CckError(ERR_NODEVICE, f"{_SO} is missing: run __graft_entry__.build() (nvcc); there is no CPU fallback")
    L = C.CDLL(_SO)
    dp, ip, vp = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.c_void_p
    u8p = C.POINTER(C.c_uint8)
    i64 = C.c_int64
    L.cck_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.cck_ctx_destroy.argtypes = [vp]
    L.cck_last_error.restype = C.c_char_p
    L.cck_last_error.argtypes = [vp]
    L.cck_ctx_synchronize.argtypes = [vp]
    L.cck_num_sms.argtypes = [vp]
    L.cck_launch_count.restype = i64
    L.cck_launch_count.argtypes = [vp]
    L.cck_model_params_default.argtypes = [C.c_int, C.c_double, C.POINTER(ModelParams)]
    L.cck_set_model.argtypes = [vp, C.POINTER(ModelParams)]
    L.cck_set_obstacles.argtypes = [vp, C.POINTER(SvcParams), vp, vp, vp, vp, C.c_int]
    L.cck_set_pvc.argtypes = [vp, C.POINTER(PvcParams)]
    L.cck_erf_inv.restype = C.c_double
    L.cck_erf_inv.argtypes = [C.c_double]
    L.cck_chi2_quantile_2dof.restype = C.c_double
    L.cck_chi2_quantile_2dof.argtypes = [C.c_double]
    L.cck_split_risk.restype = None
    L.cck_split_risk.argtypes = [C.c_double, C.c_int, C.c_int, dp, dp]
    L.cck_rect_robot_bounding_radius.restype = C.c_double
    L.cck_rect_robot_bounding_radius.argtypes = [C.c_double] * 4
    L.cck_build_obstacle_tables.argtypes = [C.c_int, vp, C.c_double, C.c_double, C.c_double, vp, vp, vp]
    L.cck_build_pair_halfplanes.argtypes = [C.c_int, C.c_double, C.c_double, vp, vp]
    L.cck_propagate.argtypes = [vp, i64, vp, i64, vp, i64, C.c_double, vp, i64]
    L.cck_propagate_dev.argtypes = [vp, vp, i64, vp, i64, vp, i64, C.c_double, vp, i64]
    L.cck_is_valid.argtypes = [vp, i64, vp, i64, vp]
    L.cck_is_valid_dev.argtypes = [vp, vp, i64, vp, i64, vp]
    L.cck_propagate_while_valid.argtypes = [vp, i64, vp, i64, vp, i64, vp, vp, i64, vp, vp, i64, C.c_int32]
    L.cck_propagate_while_valid_dev.argtypes = [vp, vp, i64, vp, i64, vp, i64, vp, vp, i64, vp, vp, i64, C.c_int32]
    L.cck_validate_plan.argtypes = [vp, C.c_int, vp, vp, i64, C.POINTER(ConflictRun)]
    L.cck_validate_plan_dev.argtypes = [vp, vp, C.c_int, vp, C.c_int, vp, i64, vp]
    L.cck_validate_plan_range.argtypes = [vp, C.c_int, vp, vp, i64, C.c_int, C.c_int, C.POINTER(ConflictRun), C.POINTER(C.c_uint64)]
    L.cck_validate_plan_range_dev.argtypes = [vp, vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, vp, i64, vp, vp]
    L.cck_set_constraints.argtypes = [vp, C.c_int, vp, vp, vp, vp]
    L.cck_set_cov_table.argtypes = [vp, vp, C.c_int]
    L.cck_host_pair_safe.argtypes = [C.c_int, vp, vp, vp, vp, vp, vp, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int]
    L.cck_host_chisq_pair_disjoint.argtypes = [vp, vp, C.c_double, vp, vp, C.c_double, C.c_double]
    L.cck_chisq_pairs_disjoint.argtypes = [vp, i64, vp, vp, vp, vp, vp, vp, C.c_double, vp]
    L.cck_satisfies_constraints.argtypes = [vp, i64, vp, i64, C.c_int32, vp, vp, vp]
    L.cck_satisfies_constraints_dev.argtypes = [vp, vp, i64, vp, i64, C.c_int32, vp, vp, vp]
    L.cck_expand_batch.argtypes = [vp, i64, vp, i64, vp, i64, vp, vp, vp, C.POINTER(GoalParams), vp, i64, vp, vp, vp, vp]
    L.cck_expand_batch_dev.argtypes = [vp, vp, i64, vp, i64, vp, i64, vp, vp, vp, C.POINTER(GoalParams), vp, i64, vp, vp, vp, vp]
    L.cck_rollout_summary_dev.argtypes = [vp, vp, i64, vp, vp, vp]
    L.cck_select_winner_dev.argtypes = [vp, vp, i64, vp, vp, C.c_int, i64, vp]
    L.cck_measure_fp64_peak.argtypes = [vp, dp]
    L.cck_measure_fp64_rates.argtypes = [vp, dp]
    L.cck_tree_create.argtypes = [vp, C.c_int, C.POINTER(vp)]
    L.cck_tree_destroy.argtypes = [vp]
    L.cck_tree_clear.argtypes = [vp]
    L.cck_tree_size.restype = i64
    L.cck_tree_size.argtypes = [vp]
    L.cck_tree_append.argtypes = [vp, i64, vp, i64, vp, C.POINTER(i64)]
    L.cck_tree_update.argtypes = [vp, i64, vp, vp, vp]
    L.cck_tree_select.argtypes = [vp, i64, vp, i64, C.c_double, vp, vp]
    L.cck_tree_nearest.argtypes = [vp, i64, vp, i64, vp, vp]
    L.cck_tree_witness_batch.argtypes = [vp, i64, vp, i64, C.c_double, vp, vp, vp, vp, vp, vp]
    L.cck_belief_distance.argtypes = [vp, C.c_int, i64, vp, i64, vp, i64, vp]
    L.cck_build_grid_table.argtypes = [C.c_int, vp, vp, C.c_double, vp, i64, C.POINTER(i64), vp]
    L.cck_build_clearance_field.argtypes = [C.c_int, vp, vp, C.c_double, vp, vp, vp, vp, C.POINTER(C.c_int)]
    L.cck_belief_distance_lower.argtypes = [vp, C.c_int, i64, vp, i64, vp]
    L.cck_belief_distance_cols.argtypes = [vp, C.c_int, i64, vp, i64, C.c_int32, vp, vp]
    L.cck_belief_distance_cols_near.argtypes = [vp, C.c_int, i64, vp, i64, C.c_int32, vp, vp, C.c_int32, vp, vp, vp]
    _lib = L
    return L


def _ptr(a):
    """numpy array / torch tensor / int / None -> void* address."""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return a.data_ptr()   # torch tensor


def width(dim):
    return dim + 2 * dim * dim


def model_params(model, dt=0.2):
    p = ModelParams()
    rc = lib().cck_model_params_default(model, dt, C.byref(p))
    if rc:
        raise CckError(rc, "cck_model_params_default")
    return p


def erf_inv(z):
    return lib().cck_erf_inv(float(z))


def chi2_quantile_2dof(p):
    return lib().cck_chi2_quantile_2dof(float(p))


def split_risk(p_safe, n_obstacles, n_robots):
    a, b = C.c_double(), C.c_double()
    lib().cck_split_risk(p_safe, n_obstacles, n_robots, C.byref(a), C.byref(b))
    return a.value, b.value


def rect_robot_bounding_radius(sx=0.0, sy=0.0, length=0.25, width_=0.25):
    return lib().cck_rect_robot_bounding_radius(sx, sy, length, width_)


def build_obstacle_tables(centres, robot_rad, length=1.0, width_=1.0):
    c = np.ascontiguousarray(centres, dtype=np.float64).reshape(-1, 2)
    M = len(c)
    A, B, R = np.zeros((M, 4, 2)), np.zeros((M, 4)), np.zeros((M, 4))
    rc = lib().cck_build_obstacle_tables(M, _ptr(c), length, width_, robot_rad, _ptr(A), _ptr(B), _ptr(R))
    if rc:
        raise CckError(rc, "cck_build_obstacle_tables")
    return A, B, R


def build_pair_halfplanes(r1, r2, bounding_box=False):
    A, B = np.zeros((4, 2)), np.zeros(4)
    rc = lib().cck_build_pair_halfplanes(int(bounding_box), r1, r2, _ptr(A), _ptr(B))
    if rc:
        raise CckError(rc, "cck_build_pair_halfplanes")
    return A, B


class Context:
    """One GPU context (device tables + streams).  Mirrors what the reference holds in its
    SpaceInformation: the installed propagator (set_model), validity checker (set_obstacles)
    and plan validator (set_pvc)."""

    def __init__(self, device=0):
        self.L = lib()
        h = C.c_void_p()
        rc = self.L.cck_ctx_create(device, C.byref(h))
        if rc:
            raise CckError(rc, self.L.cck_last_error(None).decode())
        self.h = h
        self.dim, self.dt, self.model = None, 0.2, None
        self._keep = []
        self._trees = weakref.WeakSet()   # cck_tree handles dereference the context: they must go first

    def close(self):
        if getattr(self, "h", None):
            for t in list(self._trees):
                t.close()
            self.L.cck_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc:
            raise CckError(rc, self.L.cck_last_error(self.h).decode())

    @property
    def num_sms(self):
        return self.L.cck_num_sms(self.h)

    @property
    def launch_count(self):
        return self.L.cck_launch_count(self.h)

    def synchronize(self):
        self._ck(self.L.cck_ctx_synchronize(self.h))

    def set_model(self, model, dt=0.2, params=None):
        p = params if params is not None else model_params(model, dt)
        self._ck(self.L.cck_set_model(self.h, C.byref(p)))
        self.dim, self.dt, self.model = p.dim, p.dt, p.model
        return p

    def set_obstacles(self, kind, low, high, A=None, B=None, centres=None, rects=None, erf_inv_result=0.0, p_collision=0.0,
                      sc=0.0, max_rad=0.0):
        p = SvcParams()
        p.kind, p.dim = kind, len(low)
        for i in range(len(low)):
            p.low[i], p.high[i] = float(low[i]), float(high[i])
        p.erf_inv_result, p.p_collision, p.sc, p.max_rad = erf_inv_result, p_collision, sc, max_rad
        M = 0 if A is None else len(A)
        f = lambda x: None if x is None or M == 0 else np.ascontiguousarray(x, dtype=np.float64)
        A_, B_, c_, r_ = f(A), f(B), f(centres), f(rects)
        self._ck(self.L.cck_set_obstacles(self.h, C.byref(p), _ptr(A_), _ptr(B_), _ptr(c_), _ptr(r_), M))
        self.M = M

    def set_pvc(self, kind, k, dim, radii, pair_A, pair_B, map_order, c_pair=0.0, p_coll_agnts=0.0, sc=0.0, grid_n=2):
        radii = np.ascontiguousarray(radii, dtype=np.float64)
        pair_A = np.ascontiguousarray(pair_A, dtype=np.float64)
        pair_B = np.ascontiguousarray(pair_B, dtype=np.float64)
        map_order = np.ascontiguousarray(map_order, dtype=np.int32)
        p = PvcParams()
        p.kind, p.k, p.dim, p.grid_n = kind, k, dim, int(grid_n)
        p.c_pair, p.p_coll_agnts, p.sc = c_pair, p_coll_agnts, sc
        p.radii = radii.ctypes.data_as(C.POINTER(C.c_double))
        p.pair_A = pair_A.ctypes.data_as(C.POINTER(C.c_double))
        p.pair_B = pair_B.ctypes.data_as(C.POINTER(C.c_double))
        p.map_order = map_order.ctypes.data_as(C.POINTER(C.c_int32))
        self._ck(self.L.cck_set_pvc(self.h, C.byref(p)))
        self.pvc_k, self.pvc_dim = k, dim

    # ---- host-pointer entry points (numpy SoA arrays [W, n]) ----
    def propagate(self, bel, ctrl, duration=None):
        bel = np.ascontiguousarray(bel, dtype=np.float64)
        ctrl = np.ascontiguousarray(ctrl, dtype=np.float64)
        n = bel.shape[1]
        out = np.empty_like(bel)
        self._ck(self.L.cck_propagate(self.h, n, _ptr(bel), n, _ptr(ctrl), n, self.dt if duration is None else duration, _ptr(out), n))
        return out

    def is_valid(self, bel):
        bel = np.ascontiguousarray(bel, dtype=np.float64)
        n = bel.shape[1]
        v = np.zeros(n, dtype=np.uint8)
        self._ck(self.L.cck_is_valid(self.h, n, _ptr(bel), n, _ptr(v)))
        return v.astype(bool)

    def propagate_while_valid(self, bel, ctrl, steps, traj_T=0):
        bel = np.ascontiguousarray(bel, dtype=np.float64)
        ctrl = np.ascontiguousarray(ctrl, dtype=np.float64)
        steps = np.ascontiguousarray(steps, dtype=np.int32)
        n = bel.shape[1]
        out = np.empty_like(bel)
        valid = np.zeros(n, dtype=np.int32)
        traj = np.zeros((traj_T, bel.shape[0], n)) if traj_T else None
        self._ck(self.L.cck_propagate_while_valid(self.h, n, _ptr(bel), n, _ptr(ctrl), n, _ptr(steps), _ptr(out), n, _ptr(valid),
                                                  _ptr(traj), n, traj_T))
        return (out, valid, traj) if traj_T else (out, valid)

    def validate_plan(self, trajs):
        """trajs: list of SoA arrays [W, T_r].  -> (a, b, first_step, run_len) or None."""
        k = len(trajs)
        lens = np.array([t.shape[1] for t in trajs], dtype=np.int32)
        W, T = trajs[0].shape[0], int(lens.max())
        buf = np.zeros((k, W, T))
        for r, t in enumerate(trajs):
            buf[r, :, :t.shape[1]] = t
        run = ConflictRun()
        self._ck(self.L.cck_validate_plan(self.h, k, _ptr(lens), _ptr(buf), T, C.byref(run)))
        if run.first_step < 0:
            return None
        return (run.a, run.b, run.first_step, run.run_len)

    def validate_plan_range(self, trajs, t_begin, t_end=-1):
        """One rank's share of a sharded conflict search: first conflict within the steps [t_begin, t_end) of the plan.
        -> record [key, a, b, first_step, run_len] (int64; key < 0: none) for sharding.gather_conflict."""
        k = len(trajs)
        lens = np.array([t.shape[1] for t in trajs], dtype=np.int32)
        W, T = trajs[0].shape[0], int(lens.max())
        buf = np.zeros((k, W, T))
        for r, t in enumerate(trajs):
            buf[r, :, :t.shape[1]] = t
        run, key = ConflictRun(), C.c_uint64(0)
        self._ck(self.L.cck_validate_plan_range(self.h, k, _ptr(lens), _ptr(buf), T, int(t_begin), int(t_end), C.byref(run), C.byref(key)))
        kv = -1 if key.value == 0xFFFFFFFFFFFFFFFF else int(key.value)
        return np.array([kv, run.a, run.b, run.first_step, run.run_len], dtype=np.int64)

    def chisq_pairs_disjoint(self, mu_a, cov_a, rad_a, mu_b, cov_b, rad_b, sc):
        """CentralizedChiSquaredBoundarySVC's agent-agent decagon test; SoA inputs mu [2, n], cov [4, n], rad [n]."""
        arrs = [np.ascontiguousarray(x, dtype=np.float64) for x in (mu_a, cov_a, rad_a, mu_b, cov_b, rad_b)]
        n = arrs[0].shape[1]
        out = np.zeros(n, dtype=np.uint8)
        self._ck(self.L.cck_chisq_pairs_disjoint(self.h, n, *[_ptr(x) for x in arrs], float(sc), _ptr(out)))
        return out.astype(bool)

    def set_cov_table(self, root_cov, n_steps):
        """Covariance table of a low-level tree (cck_set_cov_table); root_cov = [Sigma row-major, Lambda row-major]."""
        rc = None if root_cov is None else np.ascontiguousarray(root_cov, dtype=np.float64)
        self._keep.append(rc)
        self._ck(self.L.cck_set_cov_table(self.h, _ptr(rc), int(n_steps)))

    def set_constraints(self, ent_step, pair_ab, mu2, cov4):
        n = len(ent_step)
        if n == 0:
            self._ck(self.L.cck_set_constraints(self.h, 0, None, None, None, None))
            return
        s = np.ascontiguousarray(ent_step, dtype=np.int32)
        pr = np.ascontiguousarray(pair_ab, dtype=np.int32)
        mu = np.ascontiguousarray(mu2, dtype=np.float64)
        cv = np.ascontiguousarray(cov4, dtype=np.float64)
        self._ck(self.L.cck_set_constraints(self.h, n, _ptr(s), _ptr(pr), _ptr(mu), _ptr(cv)))

    def satisfies_constraints(self, paths, path_len, path_step0):
        """paths: [T, W, n] SoA trajectory buffer."""
        paths = np.ascontiguousarray(paths, dtype=np.float64)
        T, _, n = paths.shape
        pl = np.ascontiguousarray(path_len, dtype=np.int32)
        s0 = np.ascontiguousarray(path_step0, dtype=np.int32)
        ok = np.zeros(n, dtype=np.uint8)
        self._ck(self.L.cck_satisfies_constraints(self.h, n, _ptr(paths), n, T, _ptr(pl), _ptr(s0), _ptr(ok)))
        return ok.astype(bool)

    def expand_batch(self, bel, ctrl, steps, parent_step, parent_cost, goal_xy, threshold=2.0, goal_p_safe=0.95):
        bel = np.ascontiguousarray(bel, dtype=np.float64)
        ctrl = np.ascontiguousarray(ctrl, dtype=np.float64)
        steps = np.ascontiguousarray(steps, dtype=np.int32)
        ps = np.ascontiguousarray(parent_step, dtype=np.int32)
        pc = np.ascontiguousarray(parent_cost, dtype=np.float64)
        n = bel.shape[1]
        g = GoalParams(float(goal_xy[0]), float(goal_xy[1]), threshold, chi2_quantile_2dof(goal_p_safe))
        out = np.empty_like(bel)
        valid = np.zeros(n, dtype=np.int32)
        cost, gd, fl = np.zeros(n), np.zeros(n), np.zeros(n, dtype=np.uint8)
        self._ck(self.L.cck_expand_batch(self.h, n, _ptr(bel), n, _ptr(ctrl), n, _ptr(steps), _ptr(ps), _ptr(pc), C.byref(g),
                                         _ptr(out), n, _ptr(valid), _ptr(cost), _ptr(gd), _ptr(fl)))
        return out, valid, cost, gd, fl

    # ---- device-pointer entry points (torch CUDA tensors or raw addresses) ----
    def propagate_while_valid_dev(self, stream, n, bel, ld, ctrl, ldc, steps, out, ldo, valid, traj=None, traj_ld=0, traj_T=0):
        self._ck(self.L.cck_propagate_while_valid_dev(self.h, stream, n, _ptr(bel), ld, _ptr(ctrl), ldc, _ptr(steps), _ptr(out), ldo,
                                                      _ptr(valid), _ptr(traj), traj_ld, traj_T))

    def propagate_dev(self, stream, n, bel, ld, ctrl, ldc, out, ldo, duration=None):
        self._ck(self.L.cck_propagate_dev(self.h, stream, n, _ptr(bel), ld, _ptr(ctrl), ldc, self.dt if duration is None else duration,
                                          _ptr(out), ldo))

    def is_valid_dev(self, stream, n, bel, ld, valid):
        self._ck(self.L.cck_is_valid_dev(self.h, stream, n, _ptr(bel), ld, _ptr(valid)))

    def rollout_summary_dev(self, stream, n, steps, valid, out3):
        self._ck(self.L.cck_rollout_summary_dev(self.h, stream, n, _ptr(steps), _ptr(valid), _ptr(out3)))

    def select_winner_dev(self, stream, n, cost, flags, required_mask, index_offset, out2):
        self._ck(self.L.cck_select_winner_dev(self.h, stream, n, _ptr(cost), _ptr(flags), required_mask, index_offset, _ptr(out2)))

    def expand_batch_dev(self, stream, n, bel, ld, ctrl, ldc, steps, parent_step, parent_cost, goal_xy, out, ldo, valid, cost, goal_dist, flags,
                         threshold=2.0, goal_p_safe=0.95):
        g = GoalParams(float(goal_xy[0]), float(goal_xy[1]), threshold, chi2_quantile_2dof(goal_p_safe))
        self._ck(self.L.cck_expand_batch_dev(self.h, stream, n, _ptr(bel), ld, _ptr(ctrl), ldc, _ptr(steps), _ptr(parent_step), _ptr(parent_cost),
                                             C.byref(g), _ptr(out), ldo, _ptr(valid), _ptr(cost), _ptr(goal_dist), _ptr(flags)))

    def measure_fp64_rates(self):
        out = (C.c_double * 6)()
        self._ck(self.L.cck_measure_fp64_rates(self.h, out))
        return {"dmul_dadd": out[0], "dadd": out[1], "dmul": out[2], "with_1_fp32_per_fp64": out[3], "with_2_fp32_per_fp64": out[4],
                "with_3_fp32_per_fp64": out[5]}

    def measure_fp64_peak(self):
        t = C.c_double()
        self._ck(self.L.cck_measure_fp64_peak(self.h, C.byref(t)))
        return t.value


class Tree:
    """Device-resident tree / witness set: batched selectNode / nearest-witness queries."""

    def __init__(self, ctx, dim):
        self.ctx, self.dim = ctx, dim
        h = C.c_void_p()
        ctx._ck(ctx.L.cck_tree_create(ctx.h, dim, C.byref(h)))
        self.h = h
        ctx._trees.add(self)

    def close(self):
        # a no-op once the owning context is gone (Context.close destroys its trees before itself)
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            self.ctx.L.cck_tree_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self):
        return int(self.ctx.L.cck_tree_size(self.h))

    def clear(self):
        self.ctx._ck(self.ctx.L.cck_tree_clear(self.h))

    def append(self, bel, cost):
        bel = np.ascontiguousarray(bel, dtype=np.float64)
        cost = np.ascontiguousarray(cost, dtype=np.float64)
        n = bel.shape[1]
        first = C.c_int64(0)
        self.ctx._ck(self.ctx.L.cck_tree_append(self.h, n, _ptr(bel), n, _ptr(cost), C.byref(first)))
        return int(first.value)

    def update(self, idx, cost=None, active=None):
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        cost = None if cost is None else np.ascontiguousarray(cost, dtype=np.float64)
        active = None if active is None else np.ascontiguousarray(active, dtype=np.uint8)
        self.ctx._ck(self.ctx.L.cck_tree_update(self.h, len(idx), _ptr(idx), _ptr(cost), _ptr(active)))

    def witness_batch(self, queries, pruning_radius, eligible=None):
        """The witness rules of one launch for n <= 1024 survivors [W, n] against this tree (= the existing witnesses):
        (widx, wdist, near, near_dist, fresh), see cck_tree_witness_batch in include/cck_b200.h."""
        q = np.ascontiguousarray(queries, dtype=np.float64)
        n = q.shape[1]
        widx, near = np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int32)
        wdist, ndist = np.zeros(n), np.zeros(n)
        fresh = np.zeros(n, dtype=np.uint8)
        el = None if eligible is None else np.ascontiguousarray(eligible, dtype=np.uint8)
        self.ctx._ck(self.ctx.L.cck_tree_witness_batch(self.h, n, _ptr(q), n, float(pruning_radius), None if el is None else _ptr(el),
                                                       _ptr(widx), _ptr(wdist), _ptr(near), _ptr(ndist), _ptr(fresh)))
        return widx, wdist, near, ndist, fresh.astype(bool)

    def select(self, samples, radius=0.2):
        q = np.ascontiguousarray(samples, dtype=np.float64)
        K = q.shape[1]
        idx, dist = np.zeros(K, dtype=np.int32), np.zeros(K)
        self.ctx._ck(self.ctx.L.cck_tree_select(self.h, K, _ptr(q), K, radius, _ptr(idx), _ptr(dist)))
        return idx, dist

    def nearest(self, queries):
        q = np.ascontiguousarray(queries, dtype=np.float64)
        K = q.shape[1]
        idx, dist = np.zeros(K, dtype=np.int32), np.zeros(K)
        self.ctx._ck(self.ctx.L.cck_tree_nearest(self.h, K, _ptr(q), K, _ptr(idx), _ptr(dist)))
        return idx, dist


DF_N = 128       # CCK_DF_N


def build_clearance_field(A, B, erf_inv_result, low, high):
    """Host-only: the clearance field of the PCCBlackmore broad phase (cck_build_clearance_field) -> dict or None."""
    A = np.ascontiguousarray(A, dtype=np.float64); B = np.ascontiguousarray(B, dtype=np.float64)
    low = np.ascontiguousarray(low, dtype=np.float64); high = np.ascontiguousarray(high, dtype=np.float64)
    q = np.zeros((DF_N, DF_N), dtype=np.uint32)
    par = np.zeros(3, dtype=np.float32)
    has = C.c_int(0)
    rc = lib().cck_build_clearance_field(len(A), _ptr(A), _ptr(B), float(erf_inv_result), _ptr(low), _ptr(high), _ptr(q), _ptr(par), C.byref(has))
    if rc:
        raise CckError(rc, "cck_build_clearance_field")
    if not has.value:
        return None
    return {"q": (q & 255).astype(np.uint8), "q2": ((q >> 8) & 255).astype(np.uint8), "j1": (q >> 16).astype(np.int64),
            "icx": par[0], "icy": par[1], "k2": par[2]}


def belief_distance(ctx, a, b, dim):
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    n = a.shape[1]
    out = np.zeros(n)
    ctx._ck(ctx.L.cck_belief_distance(ctx.h, dim, n, _ptr(a), n, _ptr(b), n, _ptr(out)))
    return out


def belief_distance_lower(ctx, bel, dim):
    """[n, n] matrix whose strict lower triangle is distance(b_i, b_q) (row q, column i < q)."""
    bel = np.ascontiguousarray(bel, dtype=np.float64)
    n = bel.shape[1]
    out = np.zeros((n, n))
    ctx._ck(ctx.L.cck_belief_distance_lower(ctx.h, dim, n, _ptr(bel), n, _ptr(out)))
    return np.tril(out, -1)


def build_grid_table(A, B, erf_inv_result):
    """Host-only: parse the grid table cck_set_obstacles would upload.  Returns a dict with the header scalars,
    dstart[R*64+1] (dense row-major CSR over the cells), outer / inner boxes [M, 4] = (xlo, ylo, xhi, yhi) in row order,
    and order[M]."""
    A = np.ascontiguousarray(A, dtype=np.float64).reshape(-1, 4, 2)
    B = np.ascontiguousarray(B, dtype=np.float64).reshape(-1, 4)
    M = len(A)
    L = lib()
    nbytes = C.c_int64(0)
    rc = L.cck_build_grid_table(M, _ptr(A), _ptr(B), float(erf_inv_result), None, 0, C.byref(nbytes), None)
    if rc:
        raise CckError(rc, "cck_build_grid_table")
    buf = np.zeros(nbytes.value, dtype=np.uint8)
    order = np.zeros(max(M, 1), dtype=np.int32)
    rc = L.cck_build_grid_table(M, _ptr(A), _ptr(B), float(erf_inv_result), _ptr(buf), nbytes.value, C.byref(nbytes), _ptr(order))
    if rc:
        raise CckError(rc, "cck_build_grid_table")
    hi = buf[:40].view(np.int32)
    hf = buf[40:72].view(np.float32)
    t = {"total_bytes": int(hi[0]), "M": int(hi[1]), "n_grid": int(hi[2]), "n_cells": int(hi[3]), "R": int(hi[4]),
         "gx0": hf[0], "gy0": hf[1], "icx": hf[2], "icy": hf[3], "SX": hf[4], "SY": hf[5], "k": hf[6], "Rm1": hf[7]}
    off_cs, off_bx = int(hi[7]), int(hi[8])
    t["dstart"] = buf[off_cs:off_cs + 2 * (t["n_cells"] + 1)].view(np.uint16).copy()
    t["outer"] = buf[off_bx:off_bx + 16 * M].view(np.float32).reshape(M, 4).copy()
    t["inner"] = buf[off_bx + 16 * M:off_bx + 32 * M].view(np.float32).reshape(M, 4).copy()
    t["order"] = order[:M].copy()
    return t


def belief_distance_cols(ctx, bel, dim, cols):
    """[n, nf]: entry (q, a) = distance(b_cols[a], b_q) where cols[a] < q (0 elsewhere)."""
    bel = np.ascontiguousarray(bel, dtype=np.float64)
    cols = np.ascontiguousarray(cols, dtype=np.int32)
    n, nf = bel.shape[1], len(cols)
    out = np.zeros((n, max(nf, 1)))
    ctx._ck(ctx.L.cck_belief_distance_cols(ctx.h, dim, n, _ptr(bel), n, nf, _ptr(cols), _ptr(out)))
    out = out[:, :nf]
    out[np.arange(n)[:, None] <= cols[None, :]] = 0.0
    return out


def belief_distance_cols_near(ctx, bel, dim, cols, limit, cap=16):
    """Sparse form of belief_distance_cols: per row q the pairs (a, distance(b_cols[a], b_q)) with cols[a] < q and
    distance < limit[q], in increasing a.  Returns (count[n], idx[n, cap], dist[n, cap]); count may exceed cap."""
    bel = np.ascontiguousarray(bel, dtype=np.float64)
    cols = np.ascontiguousarray(cols, dtype=np.int32)
    limit = np.ascontiguousarray(limit, dtype=np.float64)
    n, nf = bel.shape[1], len(cols)
    idx = np.full((n, cap), -1, dtype=np.int32)
    dist = np.zeros((n, cap))
    cnt = np.zeros(n, dtype=np.int32)
    ctx._ck(ctx.L.cck_belief_distance_cols_near(ctx.h, dim, n, _ptr(bel), n, nf, _ptr(cols), _ptr(limit), cap, _ptr(idx), _ptr(dist), _ptr(cnt)))
    return cnt, idx, dist


def aos_to_soa(b):
    """[n, W] AoS records -> [W, n] SoA planes."""
    return np.ascontiguousarray(np.asarray(b, dtype=np.float64).T)


def soa_to_aos(b):
    return np.ascontiguousarray(np.asarray(b).T)


def host_pair_safe(kind, mu_a, cov_a, mu_b, cov_b, A8, B4, rad_a, rad_b, c_pair=0.0, p_coll_agnts=0.0, sc=0.0, grid_n=2):
    """Host-only: the kernels' pair-test source compiled for the CPU (cck_host_pair_safe), one pair."""
    arrs = [np.ascontiguousarray(x, dtype=np.float64) for x in (mu_a, cov_a, mu_b, cov_b, A8, B4)]
    return lib().cck_host_pair_safe(int(kind), *[_ptr(x) for x in arrs], float(rad_a), float(rad_b), float(c_pair), float(p_coll_agnts), float(sc), int(grid_n))


def robot_map_order(k):
    """Robot indices in std::map<std::string, Belief> order of the names "Robot i"
    (lexicographic: "Robot 10" < "Robot 2"; MinkowskiSumBlackmorePVC.cpp:133, quirk 6)."""
    return np.array(sorted(range(k), key=lambda i: f"Robot {i}"), dtype=np.int32)


def install_pvc(ctx, kind, dim, radii, p_safe_agents, chi_p=0.9, grid_n=2):
    """Build the plan-validity-checker constants the reference constructors compute
    (MinkowskiSumBlackmorePVC.cpp:4-30, BoundingBoxBlackmorePVC.cpp:4-32, ChiSquaredBoundaryPVC.cpp:4-20,
    Blackmore2PVC.cpp:4-30, CDFGridPVC.cpp:4-38 with grid_n = disk_steps_; demos/main.cpp:105-138) with the library's host builders and install them in `ctx`."""
    radii = np.ascontiguousarray(radii, dtype=np.float64)
    k = len(radii)
    p_coll_agnts = 1 - p_safe_agents
    p_coll_dist = p_coll_agnts / (k - 1) if k > 1 else p_coll_agnts
    c_pair = erf_inv(1 - (2 * p_coll_dist))
    bb = kind in (PVC_BOUNDINGBOX, PVC_ADAPTIVE_BOUNDINGBOX)
    A = np.zeros((k * k, 4, 2))
    B = np.zeros((k * k, 4))
    for a in range(k):
        for b in range(a + 1, k):
            A[a * k + b], B[a * k + b] = build_pair_halfplanes(radii[a], radii[b], bb)
    sc = chi2_quantile_2dof(chi_p)
    ctx.set_pvc(kind, k, dim, radii, A, B, robot_map_order(k), c_pair=c_pair, p_coll_agnts=p_coll_agnts, sc=sc, grid_n=grid_n)
    return {"c_pair": c_pair, "p_coll_dist": p_coll_dist, "pair_A": A, "pair_B": B, "sc": sc}


def constraint_entries(k, dim, constrained, constraints):
    """Flatten BeliefConstraints for cck_set_constraints.  constraints: list of
    (constraining_robot, steps[int], states AoS [len(steps), W]) as BeliefPVC::createConstraint
    builds them (BeliefPVC.cpp:10-40).  Returns (ent_step, pair_ab, mu2, cov4)."""
    ent_step, pair_ab, mu, cov = [], [], [], []
    for other, steps, states in constraints:
        st = np.asarray(states, dtype=np.float64).reshape(len(steps), -1)
        for s, b in zip(steps, st):
            ent_step.append(int(s))
            pair_ab.append(constrained * k + other)      # planning robot first (cck_b200.h)
            mu.append(b[:2])
            S, L = b[dim:dim + dim * dim], b[dim + dim * dim:]
            if dim == 4:   # BeliefPVC::getDistribution_ picks (0,0),(0,2),(2,0),(2,2)
                cov.append([S[0] + L[0], S[2] + L[2], S[8] + L[8], S[10] + L[10]])
            else:
                cov.append(S + L)
    return (np.array(ent_step, dtype=np.int32), np.array(pair_ab, dtype=np.int32),
            np.array(mu, dtype=np.float64).reshape(-1, 2), np.array(cov, dtype=np.float64).reshape(-1, 4))
