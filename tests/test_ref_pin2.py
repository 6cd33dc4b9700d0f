"""Round-2 pins: the rows that round 1 left on an unchecked restatement - the unicycle propagator, the belief-space
distance, the goal test, the chi-squared state / plan checkers (and the CDF-grid / Blackmore2 plan checkers) - against
vectors produced by the REFERENCE'S OWN SOURCES (UncertainUnicycleSP.cpp, RealVectorBeliefSpace.cpp, BeliefSpaceGoals.cpp,
ChiSquaredBoundarySVC.cpp, ChiSquaredBoundaryPVC.cpp, CDFGridPVC.cpp, Blackmore2PVC.cpp compiled in oracle/_ref against
oracle/shim/, generator oracle/gen_ref_golden2.py).

Bars (BASELINE.json north_star): propagated means and covariances within 1e-9 relative; validity / conflict verdicts and
first-violation steps exact outside a +-1e-6 band around the decision threshold.  Every mismatch must be SHOWN to lie in
that band (helpers.*band*); on these fixtures there are none, which is asserted too where it holds.

CPU part: the oracle against the vectors (and against the live reference build where /root/reference is mounted).
GPU part (-m gpu): the CUDA path against the same vectors, through the C ABI."""
import os

import numpy as np
import pytest

from conftest import load_golden
import helpers as H

RTOL = 1e-9


def close(a, b, rtol=RTOL, atol_scale=1e-13):
    """|a - b| <= rtol |b| + atol, atol = atol_scale * typical magnitude of the compared block: entries that cancel to
    ~0 (cross-covariances) cannot hold a relative bound of their own."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    atol = atol_scale * max(1.0, float(np.max(np.abs(b))) if b.size else 1.0)
    return np.allclose(a, b, rtol=rtol, atol=atol, equal_nan=True)


def g3():
    return load_golden("ref2_config3_random32_5_unicycle_k4.npz")


def g2():
    return load_golden("ref2_config2_random32_10_linear_k4.npz")


def oracle_world(O, g):
    """The oracle's own parser on the reference's map / scenario text (committed in the fixture directory by round 1's
    generator as config*.npz; here rebuilt from the stored obstacle list, which test_ref_pin.py proves equal)."""
    w = O.World.synthetic(np.asarray(g["centres"]), float(g["x_max"]), float(g["y_max"]), int(g["k"]), float(g["p_safe"]))
    assert w.n_obstacles == int(g["n_obstacles"])
    return w


# ---------------------------------------------------------------------------------------------------------------
# a2: DynUnicycleControlSpace
# ---------------------------------------------------------------------------------------------------------------
def test_unicycle_constants_vs_reference(O):
    g = g3()
    p = O.unicycle_params(0.2)
    # F = exp(A_ol dt) and A_cl_d = F - B_d K through the reference's constructor (Pade + LU in the stand-in): the oracle's
    # closed form differs by at most one ulp in A_cl_d(0,1), A_cl_d(2,3)
    assert close(p["F"].reshape(4, 4), g["uni_F"], 1e-15)
    assert close(p["Acl"].reshape(4, 4), g["uni_Acl"], 4e-16)
    assert np.array_equal(p["Q"].reshape(4, 4), g["uni_Q"]) and np.array_equal(p["R"].reshape(4, 4), g["uni_R"])
    # the A_cl_d block
    assert np.allclose(g["uni_Acl"][:2, :2], [[0.948286, 0.131132], [-0.51714, 0.31132]], rtol=0, atol=1e-12)


def test_unicycle_propagate_vs_reference(O):
    g = g3()
    b, u = g["beliefs"], g["controls"]
    out = O.propagate(b, u, 4)
    # the mean step has no third-party arithmetic: bit-exact (libm sin/cos on both sides)
    assert np.array_equal(out[:, :4], g["propagate"][:, :4])
    assert close(out[:, 4:], g["propagate"][:, 4:])
    # quirk 1: the unicycle uses the duration ARGUMENT for the mean (the covariance step does not depend on it)
    out01 = O.propagate(b[:300], u[:300], 4, duration=0.1)
    assert np.array_equal(out01[:, :4], g["propagate_d01"][:, :4])
    assert close(out01[:, 4:], g["propagate_d01"][:, 4:])
    assert not np.array_equal(out01[:, :2], out[:300, :2])
    # quirk 2: Sigma' is NOT symmetric (F Sigma F without the transpose) - in the reference's output itself
    S = g["propagate"][:, 4:20].reshape(-1, 4, 4)
    assert np.abs(S - np.transpose(S, (0, 2, 1))).max() > 1e-4


def test_unicycle_chain_vs_reference(O):
    """50 steps from each robot's start state (Sigma0 = Lambda0 = 0.01 I, yaw 0, surge 0.1: OmplSetUp.cpp:315-320)."""
    g = g3()
    cur = g["chain0"].copy()
    for t in range(50):
        cur = O.propagate(cur, g["chain_u"], 4)
        assert close(cur, g["chain"][t]), t


def blackmore_band_unicycle(svc, aos, c):
    """the (x, xdot) block the state checker reads for 4-D states (PCCBlackmoreSVC.cpp:41, quirk 2)"""
    P = np.stack([aos[:, 4] + aos[:, 20], aos[:, 5] + aos[:, 21], aos[:, 8] + aos[:, 24], aos[:, 9] + aos[:, 25]], axis=1)
    return H.band_mask_halfplane(svc, aos, P, c)


def test_unicycle_blackmore_svc_vs_reference(O):
    g = g3()
    w = oracle_world(O, g)
    svc = O.Svc(w, O.SVC_BLACKMORE, 4)
    valid, _ = svc.is_valid(g["beliefs"])
    bad = valid != g["bm_is_valid"]
    assert not (bad & ~blackmore_band_unicycle(svc, g["beliefs"], svc.erf_inv_result)).any()
    assert bad.sum() == 0
    fin, nval, _ = svc.rollout(g["beliefs"], g["controls"], g["steps"])
    assert np.array_equal(nval, g["bm_rollout_valid"])
    assert np.array_equal(fin[:, :4], g["bm_rollout_final"][:, :4])
    assert close(fin[:, 4:], g["bm_rollout_final"][:, 4:])
    assert 0.2 < (nval == g["steps"]).mean() < 0.9


# ---------------------------------------------------------------------------------------------------------------
# a6: ChiSquaredBoundarySVC (decagon of radius lambda_max * sc + r vs the inflated rectangles)
# ---------------------------------------------------------------------------------------------------------------
def chisq_band(O, world, dim, aos, eps=1e-6):
    """True where the verdict changes when the belief's decagon is grown / shrunk by eps (absolute, in distance units):
    the stated band around the decision threshold, evaluated with the oracle on a robot radius shifted by +-eps."""
    lo = O.Svc(world, O.SVC_CHISQUARED, dim)
    v0, _ = lo.is_valid(aos)
    flips = np.zeros(len(aos), dtype=bool)
    for d in (-eps, eps):
        s = O.Svc(world, O.SVC_CHISQUARED, dim)
        s.set_max_rad(s.max_rad + d)
        v, _ = s.is_valid(aos)
        flips |= v != v0
    return flips


@pytest.mark.parametrize("which", ["linear", "unicycle"])
def test_chisq_svc_vs_reference(O, which):
    g = g2() if which == "linear" else g3()
    dim = 2 if which == "linear" else 4
    w = oracle_world(O, g)
    svc = O.Svc(w, O.SVC_CHISQUARED, dim)
    sc, max_rad, n_obs = g["cs_info"]
    assert abs(svc.sc - sc) <= 2e-16 * sc and svc.max_rad == max_rad and int(n_obs) == w.n_obstacles   # quirk 3: constructed with 0.9
    assert abs(sc - 4.605170185988092) < 1e-14
    valid, _ = svc.is_valid(g["beliefs"])
    bad = valid != g["cs_is_valid"]
    assert not (bad & ~chisq_band(O, w, dim, g["beliefs"])).any()
    assert bad.sum() == 0
    n = len(g["cs_rollout_valid"])
    fin, nval, _ = svc.rollout(g["beliefs"][:n], g["controls"][:n], g["steps"][:n])
    assert np.array_equal(nval, g["cs_rollout_valid"])
    assert close(fin, g["cs_rollout_final"])
    assert 0.3 < valid.mean() < 0.95


# ---------------------------------------------------------------------------------------------------------------
# f2: RealVectorBeliefSpace::distance, f1: ChanceConstrainedGoal
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("which", ["linear", "unicycle"])
def test_distance_vs_reference(O, which):
    g = g2() if which == "linear" else g3()
    dim = 2 if which == "linear" else 4
    d = O.belief_distance(g["dist_a"], g["dist_b"], dim)
    assert close(d, g["dist"], 1e-9)
    assert (g["dist"] > 0).all()
    # identical means: exactly 0 whatever the covariances (RealVectorBeliefSpace.cpp:27-28)
    a = g["dist_a"][:50].copy(); b = g["dist_b"][:50].copy(); b[:, :dim] = a[:, :dim]
    assert (O.belief_distance(a, b, dim) == 0).all()


def goal_band(O, aos, dim, gx, gy, eps=1e-6):
    ok0, _ = O.goal(aos, dim, gx, gy)
    flips = np.zeros(len(aos), dtype=bool)
    for thr in (2.0 - eps, 2.0 + eps):
        ok, _ = O.goal(aos, dim, gx, gy, threshold=thr)
        flips |= ok != ok0
    return flips


@pytest.mark.parametrize("which", ["linear", "unicycle"])
def test_goal_vs_reference(O, which):
    g = g2() if which == "linear" else g3()
    dim = 2 if which == "linear" else 4
    gx, gy = g["goal_xy"]
    ok, dist = O.goal(g["goal_beliefs"], dim, gx, gy)
    assert np.array_equal(dist, g["goal_dist"])                      # dx*dx + dy*dy: no third-party arithmetic
    bad = ok != g["goal_ok"]
    assert not (bad & ~goal_band(O, g["goal_beliefs"], dim, gx, gy)).any()
    assert bad.sum() == 0
    assert 0.2 < ok.mean() < 0.9


# ---------------------------------------------------------------------------------------------------------------
# a9: ChiSquaredBoundaryPVC; unicycle getDistribution_ sub-block through MinkowskiSumBlackmorePVC
# ---------------------------------------------------------------------------------------------------------------
def plans_of(g, tag):
    lens, states = g[tag + "plan_lens"], g[tag + "plan_states"]
    off = 0
    for i in range(len(lens)):
        trajs = []
        for L in lens[i]:
            trajs.append(states[off:off + L]); off += L
        yield i, trajs


def check_pvc_block(O, g, tag, okind, dim, p_arg=-1.0):
    w = oracle_world(O, g)
    pv = O.Pvc(w, okind, p_arg)
    want = g[tag + "conflicts"]
    sat = g[tag + "satisfies"]
    n_conf = 0
    for i, trajs in plans_of(g, tag):
        got = pv.validate_plan(trajs, dim)
        exp = [tuple(int(x) for x in r[1:]) for r in want[want[:, 0] == i]]
        assert got == exp, (tag, i)
        n_conf += bool(exp)
        if exp:
            a, other = exp[0][0], exp[0][1]
            steps = [e[2] for e in exp]
            cst = np.array([trajs[other][min(s, len(trajs[other]) - 1)] for s in steps])
            for j in range(sat.shape[1]):
                probe = g[tag + "probe_states"][i, j, :g[tag + "probe_len"][i, j]]
                assert pv.satisfies_constraints(probe, dim, a, [(other, steps, cst)]) == bool(sat[i, j]), (tag, i, j)
    assert n_conf >= 10 and (sat == 1).any() and (sat == 0).any()
    return pv


def test_chisq_pvc_vs_reference(O):
    for g, dim in ((g2(), 2), (g3(), 4)):
        pv = check_pvc_block(O, g, "pcs_", O.PVC_CHISQUARED, dim, 0.9)
        assert abs(pv.sc - g["pcs_info"][1]) <= 2e-16 * pv.sc            # demos/main.cpp:111: hard-coded 0.9


def test_unicycle_minkowski_pvc_vs_reference(O):
    """4-D states: BeliefPVC::getDistribution_ reads entries (0,0),(0,2),(2,0),(2,2) (BeliefPVC.cpp:91-99)."""
    check_pvc_block(O, g3(), "pbm_", O.PVC_BLACKMORE, 4)


# ---------------------------------------------------------------------------------------------------------------
# live reference build (only where /root/reference is mounted): fresh random inputs, not the committed ones
# ---------------------------------------------------------------------------------------------------------------
def _ref():
    from oracle import ref as R
    if not os.path.isdir(R.REFERENCE):
        pytest.skip("/root/reference is not mounted here (the committed ref2_*.npz vectors carry the pin)")
    R.build()
    return R


def test_live_reference_unicycle_and_distance(O):
    R = _ref()
    REF = R.REFERENCE
    w = R.World(f"{REF}/maps/random-32-32-5.map", f"{REF}/scens/random-32-32-5-Rectangles-Uncertain-Unicycle.scen", 4, 0.9)
    from oracle.gen_ref_golden2 import beliefs
    b, u, steps = beliefs(1500, 4, w.x_max, w.y_max, 555)
    s = R.Svc2(w, "unicycle", R.SVC2_BLACKMORE)
    ow = O.World.from_files(f"{REF}/maps/random-32-32-5.map", f"{REF}/scens/random-32-32-5-Rectangles-Uncertain-Unicycle.scen", 4, 0.9)
    osvc = O.Svc(ow, O.SVC_BLACKMORE, 4)
    fin_r, nv_r = s.rollout(b, u, steps)
    fin_o, nv_o, _ = osvc.rollout(b, u, steps)
    assert np.array_equal(nv_r, nv_o) and close(fin_o, fin_r)
    for dim, bb in ((4, b), (2, beliefs(1500, 2, 32.0, 32.0, 556)[0])):
        assert close(O.belief_distance(bb[:700], bb[700:1400], dim), R.distance(dim, bb[:700], bb[700:1400]))


# ---------------------------------------------------------------------------------------------------------------
# GPU: the CUDA path (through the C ABI) against the same reference vectors
# ---------------------------------------------------------------------------------------------------------------
def install_world(K, ctx, g, model, svc_kind):
    """Product-side constant builders (cck_split_risk, cck_build_obstacle_tables, cck_erf_inv, cck_chi2_quantile_2dof)
    on the obstacle list the reference's parser produced; wiring of OmplSetUp.cpp:183-353."""
    dim = 4 if model == K.MODEL_UNICYCLE else 2
    k, M = int(g["k"]), int(g["n_obstacles"])
    ctx.set_model(model)
    rad = K.rect_robot_bounding_radius(float(g["starts"][0, 0]), float(g["starts"][0, 1]))
    assert rad == float(g["bounding_rad"][0])
    p_safe_obs, _ = K.split_risk(float(g["p_safe"]), M, k)
    assert p_safe_obs == float(g["p_safe_obs"])
    A, B, rects = K.build_obstacle_tables(g["centres"], rad)
    low = np.array([-1.0, -1.0, -np.pi, 0.01][:dim])
    high = np.array([float(g["x_max"]), float(g["y_max"]), np.pi, 10.0][:dim])
    c = K.erf_inv(1 - 2 * ((1 - p_safe_obs) / M))
    ctx.set_obstacles(svc_kind, low, high, A, B, g["centres"], rects, erf_inv_result=c, p_collision=1 - p_safe_obs,
                      sc=K.chi2_quantile_2dof(0.9), max_rad=rad)
    return rad, {"A": A, "B": B, "c": c}


class _Tab:          # what helpers.band_mask_halfplane reads
    def __init__(self, A, B):
        self.A, self.B = A, B


@pytest.mark.gpu
def test_cuda_unicycle_vs_reference_vectors(K, O, ctx):
    g = g3()
    rad, tab = install_world(K, ctx, g, K.MODEL_UNICYCLE, K.SVC_BLACKMORE)
    bel, ctrl = K.aos_to_soa(g["beliefs"]), K.aos_to_soa(g["controls"])
    out = K.soa_to_aos(ctx.propagate(bel, ctrl))
    assert close(out, g["propagate"])                     # device sincos differs from libm in the last ulps: 1e-9 relative
    assert close(out[:, 4:], g["propagate"][:, 4:])
    valid = ctx.is_valid(bel)
    bad = valid != g["bm_is_valid"]
    assert not (bad & ~blackmore_band_unicycle(_Tab(tab["A"], tab["B"]), g["beliefs"], tab["c"])).any()
    fin, nval = ctx.propagate_while_valid(bel, ctrl, g["steps"])
    fin = K.soa_to_aos(fin)
    diff = nval != g["bm_rollout_valid"]
    if diff.any():      # a first-violation step may only differ where the deciding state sits inside the +-1e-6 band
        idx = np.nonzero(diff)[0]
        osvc = O.Svc(oracle_world(O, g), O.SVC_BLACKMORE, 4)
        for i in idx:
            t = min(nval[i], g["bm_rollout_valid"][i])
            st = O.propagate_n(g["beliefs"][i], g["controls"][i], t + 1, 4)[-1:]
            assert blackmore_band_unicycle(osvc, st, osvc.erf_inv_result)[0], i
    assert diff.sum() <= 2
    same = ~diff
    assert close(fin[same], g["bm_rollout_final"][same])
    # the planner's regime: 50-step chains from the robots' start states
    cur = K.aos_to_soa(g["chain0"])
    cu = K.aos_to_soa(g["chain_u"])
    for t in range(50):
        cur = ctx.propagate(cur, cu)
        assert close(K.soa_to_aos(cur), g["chain"][t]), t


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["linear", "unicycle"])
def test_cuda_chisq_svc_vs_reference_vectors(K, O, ctx, which):
    g = g2() if which == "linear" else g3()
    model = K.MODEL_LINEAR2D if which == "linear" else K.MODEL_UNICYCLE
    dim = 2 if which == "linear" else 4
    install_world(K, ctx, g, model, K.SVC_CHISQUARED)
    bel, ctrl = K.aos_to_soa(g["beliefs"]), K.aos_to_soa(g["controls"])
    valid = ctx.is_valid(bel)
    bad = valid != g["cs_is_valid"]
    assert not (bad & ~chisq_band(O, oracle_world(O, g), dim, g["beliefs"])).any()
    assert bad.sum() == 0
    n = len(g["cs_rollout_valid"])
    fin, nval = ctx.propagate_while_valid(bel[:, :n].copy(), ctrl[:, :n].copy(), g["steps"][:n])
    assert np.array_equal(nval, g["cs_rollout_valid"])
    assert close(K.soa_to_aos(fin), g["cs_rollout_final"])


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["linear", "unicycle"])
def test_cuda_distance_and_goal_vs_reference_vectors(K, O, ctx, which):
    g = g2() if which == "linear" else g3()
    model = K.MODEL_LINEAR2D if which == "linear" else K.MODEL_UNICYCLE
    dim = 2 if which == "linear" else 4
    d = K.belief_distance(ctx, K.aos_to_soa(g["dist_a"]), K.aos_to_soa(g["dist_b"]), dim)
    assert close(d, g["dist"])
    # goal: one expansion launch with steps = 0 evaluates ChanceConstrainedGoal on the parent state itself
    install_world(K, ctx, g, model, K.SVC_BLACKMORE)
    gb = g["goal_beliefs"]
    n = len(gb)
    gx, gy = g["goal_xy"]
    _, _, _, gdist, flags = ctx.expand_batch(K.aos_to_soa(gb), np.zeros((dim, n)), np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int32),
                                             np.zeros(n), (gx, gy))
    ok = (flags & 4) != 0
    assert np.array_equal(gdist, g["goal_dist"])
    bad = ok != g["goal_ok"]
    assert not (bad & ~goal_band(O, gb, dim, gx, gy)).any()
    assert bad.sum() == 0


def _cuda_pvc_block(K, ctx, g, tag, kind, dim):
    k = int(g["k"])
    rad = K.rect_robot_bounding_radius()
    K.install_pvc(ctx, kind, dim, [rad] * k, float(g["p_safe_agents"]))
    want, sat = g[tag + "conflicts"], g[tag + "satisfies"]
    for i, trajs in plans_of(g, tag):
        exp = want[want[:, 0] == i]
        run = ctx.validate_plan([K.aos_to_soa(t) for t in trajs])
        if len(exp) == 0:
            assert run is None, (tag, i)
            continue
        assert run == (int(exp[0, 1]), int(exp[0, 2]), int(exp[0, 3]), len(exp)), (tag, i, run, exp[0])
        a, other, steps = int(exp[0, 1]), int(exp[0, 2]), [int(x) for x in exp[:, 3]]
        cst = np.array([trajs[other][min(s, len(trajs[other]) - 1)] for s in steps])
        ctx.set_constraints(*K.constraint_entries(k, dim, a, [(other, steps, cst)]))
        soa = np.ascontiguousarray(g[tag + "probe_states"][i].transpose(1, 2, 0))        # [T, W, n]
        ok = ctx.satisfies_constraints(soa, g[tag + "probe_len"][i], np.zeros(sat.shape[1], dtype=np.int32))
        assert np.array_equal(ok, sat[i].astype(bool)), (tag, i)


@pytest.mark.gpu
def test_cuda_chisq_pvc_vs_reference_vectors(K, ctx):
    ctx.set_model(K.MODEL_LINEAR2D)
    _cuda_pvc_block(K, ctx, g2(), "pcs_", K.PVC_CHISQUARED, 2)
    ctx.set_model(K.MODEL_UNICYCLE)
    _cuda_pvc_block(K, ctx, g3(), "pcs_", K.PVC_CHISQUARED, 4)
    _cuda_pvc_block(K, ctx, g3(), "pbm_", K.PVC_BLACKMORE, 4)


# ---------------------------------------------------------------------------------------------------------------
# Blackmore2PVC, CDFGridPVC(n) - the reference's own Blackmore2PVC.cpp / CDFGridPVC.cpp compiled in
# oracle/_ref (EigenSolver / LLT / erf through the stand-ins)
# ---------------------------------------------------------------------------------------------------------------
F4 = [("pb2_", "PVC_BLACKMORE2", 2), ("cdf2_", "PVC_CDFGRID", 2), ("cdf5_", "PVC_CDFGRID", 5)]


def cdfgrid_band(O, g, trajs, dim, n, eps=1e-9):
    """True if some pair test of the plan has its grid probability within eps of p_coll_agnts (the CDF-grid threshold)."""
    rad = float(g["bounding_rad"][0])
    p_coll = 1 - float(g["p_safe_agents"])
    T = max(len(t) for t in trajs)
    k = len(trajs)
    for t in range(T + 1):
        st = [tr[min(t, len(tr) - 1)] for tr in trajs]
        for a in range(k):
            for b in range(a + 1, k):
                pr = O.cdfgrid_prob([st[a][:2]], [st[a][2:6] + st[a][6:10]], [st[b][:2]], [st[b][2:6] + st[b][6:10]], 2 * rad, n)[0]
                if abs(pr - p_coll) < eps:
                    return True
    return False


@pytest.mark.parametrize("tag,kind,n", F4)
def test_f4_plan_checkers_vs_reference(O, tag, kind, n):
    g = g2()
    w = oracle_world(O, g)
    pv = O.Pvc(w, getattr(O, kind), grid_n=n)
    assert pv.p_coll_agnts == g[tag + "info"][0]
    want, sat = g[tag + "conflicts"], g[tag + "satisfies"]
    n_conf = 0
    for i, trajs in plans_of(g, tag):
        got = pv.validate_plan(trajs, 2)
        exp = [tuple(int(x) for x in r[1:]) for r in want[want[:, 0] == i]]
        assert got == exp, (tag, i)
        n_conf += bool(exp)
        if exp:
            a, other = exp[0][0], exp[0][1]
            steps = [e[2] for e in exp]
            cst = np.array([trajs[other][min(s, len(trajs[other]) - 1)] for s in steps])
            for j in range(sat.shape[1]):
                probe = g[tag + "probe_states"][i, j, :g[tag + "probe_len"][i, j]]
                assert pv.satisfies_constraints(probe, 2, a, [(other, steps, cst)]) == bool(sat[i, j]), (tag, i, j)
    assert n_conf >= 10 and (sat == 1).any() and (sat == 0).any()


def test_cdfgrid_probability_is_basis_invariant(O):
    """The closed-form eigenvectors may differ from Eigen's in order and sign: the grid mass must not care (standard normal
    symmetric under axis swaps / reflections), and for n -> large it approaches the exact mass of the box."""
    rng = np.random.default_rng(4)
    n = 300
    mu_a, mu_b = rng.uniform(-1, 1, (n, 2)), rng.uniform(-1, 1, (n, 2))
    L = rng.uniform(0.05, 0.3, (n, 3))
    cov = np.stack([L[:, 0] ** 2, L[:, 0] * L[:, 1], L[:, 0] * L[:, 1], L[:, 1] ** 2 + L[:, 2] ** 2], axis=1)
    p = O.cdfgrid_prob(mu_a, cov, mu_b, cov, 0.35, 5)
    # mirror the whole configuration in x: same probability
    fx = np.array([-1.0, 1.0])
    covm = cov * np.array([1.0, -1.0, -1.0, 1.0])
    pm = O.cdfgrid_prob(mu_a * fx, covm, mu_b * fx, covm, 0.35, 5)
    assert np.allclose(p, pm, rtol=1e-12, atol=1e-15)
    # swap x and y
    sw = cov[:, [3, 2, 1, 0]]
    ps = O.cdfgrid_prob(mu_a[:, ::-1], sw, mu_b[:, ::-1], sw, 0.35, 5)
    assert np.allclose(p, ps, rtol=1e-12, atol=1e-15)
    assert (p >= 0).all() and (p <= 1.0 + 1e-12).all()
    # n = 2 integrates the bounding box of the mapped square: an upper bound of every finer grid's covered mass? not in general
    # (cells are counted whole), but both bound the exact mass of the square from above
    p2 = O.cdfgrid_prob(mu_a, cov, mu_b, cov, 0.35, 2)
    p40 = O.cdfgrid_prob(mu_a, cov, mu_b, cov, 0.35, 40)
    assert (p2 >= p40 - 1e-12).all() and (p >= p40 - 1e-12).all()


@pytest.mark.gpu
@pytest.mark.parametrize("tag,kind,n", F4)
def test_cuda_f4_plan_checkers_vs_reference_vectors(K, O, ctx, tag, kind, n):
    g = g2()
    k = int(g["k"])
    ctx.set_model(K.MODEL_LINEAR2D)
    rad = K.rect_robot_bounding_radius()
    K.install_pvc(ctx, getattr(K, kind), 2, [rad] * k, float(g["p_safe_agents"]), grid_n=n)
    want, sat = g[tag + "conflicts"], g[tag + "satisfies"]
    for i, trajs in plans_of(g, tag):
        exp = want[want[:, 0] == i]
        run = ctx.validate_plan([K.aos_to_soa(t) for t in trajs])
        wanted = None if len(exp) == 0 else (int(exp[0, 1]), int(exp[0, 2]), int(exp[0, 3]), len(exp))
        if run != wanted:      # device erf vs the reference build's: only inside the band
            assert kind == "PVC_CDFGRID" and cdfgrid_band(O, g, trajs, 2, n), (tag, i, run, wanted)
            continue
        if wanted is None:
            continue
        a, other, steps = wanted[0], wanted[1], [int(x) for x in exp[:, 3]]
        cst = np.array([trajs[other][min(s, len(trajs[other]) - 1)] for s in steps])
        ctx.set_constraints(*K.constraint_entries(k, 2, a, [(other, steps, cst)]))
        soa = np.ascontiguousarray(g[tag + "probe_states"][i].transpose(1, 2, 0))
        ok = ctx.satisfies_constraints(soa, g[tag + "probe_len"][i], np.zeros(sat.shape[1], dtype=np.int32))
        assert np.array_equal(ok, sat[i].astype(bool)), (tag, i)


def test_centralized_propagator_vs_reference(O):
    """CentralizedUncertainLinearStatePropagator::propagate (the reference's CentralizedUncertainLinearSP.cpp compiled in
    oracle/_ref): every agent's 2x2 blocks advance exactly like a single-agent linear belief with step size 0.5, whatever sits
    outside the blocks; the rest of the result keeps the allocState default 0.01 I."""
    g = g2()
    k = int(g["central_k"])
    cs, cu, want = g["central_states"], g["central_controls"], g["central_propagate"]
    n, d = len(cs), 2 * k
    S = cs[:, d:d + d * d].reshape(n, d, d); L = cs[:, d + d * d:].reshape(n, d, d)
    Sw = want[:, d:d + d * d].reshape(n, d, d); Lw = want[:, d + d * d:].reshape(n, d, d)
    for a in range(k):
        b = np.concatenate([cs[:, 2 * a:2 * a + 2], S[:, 2 * a:2 * a + 2, 2 * a:2 * a + 2].reshape(n, 4), L[:, 2 * a:2 * a + 2, 2 * a:2 * a + 2].reshape(n, 4)], axis=1)
        got = O.propagate(b, cu[:, 2 * a:2 * a + 2], 2, dt=0.5)
        assert np.array_equal(got[:, :2], want[:, 2 * a:2 * a + 2])
        assert np.array_equal(got[:, 2:6], Sw[:, 2 * a:2 * a + 2, 2 * a:2 * a + 2].reshape(n, 4))
        assert np.array_equal(got[:, 6:10], Lw[:, 2 * a:2 * a + 2, 2 * a:2 * a + 2].reshape(n, 4))
    off = np.ones((d, d), dtype=bool)
    for a in range(k):
        off[2 * a:2 * a + 2, 2 * a:2 * a + 2] = False
    assert (Sw[:, off] == 0).all() and (Lw[:, off] == 0).all()        # result state: allocState default, off-diagonal zeros
