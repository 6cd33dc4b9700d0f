"""GPU parity against the committed golden fixtures (BASELINE.json configs 1-4, generated
from the reference's maps/scens through the oracle) and, for the plan-validity /
constraint / batched-expansion entry points, against the live oracle."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN
from helpers import assert_rollout_mismatches_in_band, svc_band

pytestmark = pytest.mark.gpu

FILES = sorted(glob.glob(os.path.join(GOLDEN, "config*.npz")))
IDS = [os.path.basename(f)[:-4] for f in FILES]
RTOL = 1e-9


def setup_ctx(K, ctx, g, svc_kind):
    """Product-side configuration built with the library's own host builders (no oracle)."""
    dim, M, k = int(g["dim"]), len(g["centres"]), int(g["k"])
    ctx.set_model(K.MODEL_LINEAR2D if dim == 2 else K.MODEL_UNICYCLE)
    rad = K.rect_robot_bounding_radius(float(g["starts"][0, 0]), float(g["starts"][0, 1]), 0.25, 0.25)
    A, B, R = K.build_obstacle_tables(g["centres"], rad)
    p_obs, p_ag = K.split_risk(float(g["p_safe"]), M, k)
    if dim == 2:
        low, high = [-1.0, -1.0], [float(g["x_max"]), float(g["y_max"])]
    else:
        low, high = [-1.0, -1.0, -np.pi, 0.01], [float(g["x_max"]), float(g["y_max"]), np.pi, 10.0]
    accep = 0.9 if svc_kind == K.SVC_CHISQUARED else p_obs          # OmplSetUp.cpp:213-222
    p_coll = 1 - accep
    c = K.erf_inv(1 - 2 * (p_coll / M)) if M > 0 else 0.0
    ctx.set_obstacles(svc_kind, low, high, A, B, g["centres"], R, erf_inv_result=c, p_collision=p_coll,
                      sc=K.chi2_quantile_2dof(accep), max_rad=rad)
    return rad, p_ag


def trajs_of(g):
    lens = g["traj_lens"]
    offs = np.concatenate([[0], np.cumsum(lens)])
    return [g["traj_states"][offs[i]:offs[i + 1]] for i in range(len(lens))]


@pytest.mark.parametrize("path", FILES, ids=IDS)
def test_golden_propagate_valid_rollout(K, O, ctx, path):
    g = np.load(path)
    dim = int(g["dim"])
    oworld = O.World.synthetic(g["centres"], float(g["x_max"]), float(g["y_max"]), int(g["k"]), float(g["p_safe"]))
    bel, ctrl, steps = K.aos_to_soa(g["beliefs"]), K.aos_to_soa(g["controls"]), g["steps"]
    for kind, tag in ((K.SVC_BLACKMORE, "bm"), (K.SVC_ADAPTIVE, "ad"), (K.SVC_CHISQUARED, "cs")):
        setup_ctx(K, ctx, g, kind)
        if kind == K.SVC_BLACKMORE:
            out1 = K.soa_to_aos(ctx.propagate(bel, ctrl))
            if dim == 2:
                assert np.array_equal(out1, g["prop1"])
            else:
                assert np.allclose(out1, g["prop1"], rtol=RTOL, atol=1e-13)
        ok = ctx.is_valid(bel)
        fin, valid = ctx.propagate_while_valid(bel, ctrl, steps)
        fin = K.soa_to_aos(fin)
        exact = dim == 2 and kind != K.SVC_ADAPTIVE
        if exact:   # bit-exact: values, verdicts, first-violation step
            assert np.array_equal(ok, g[f"{tag}_isvalid"]), tag
            assert np.array_equal(valid, g[f"{tag}_valid"]), tag
            assert np.array_equal(fin, g[f"{tag}_final"]), tag
        else:       # device sincos / erfinv: exact outside the +-1e-6 band (every mismatch is SHOWN to lie inside it), values to 1e-9
            osvc = O.Svc(oworld, kind, dim)
            bad = ok != g[f"{tag}_isvalid"]
            assert bad.sum() <= 2 and svc_band(O, osvc, g["beliefs"][bad]).all(), tag
            agree = assert_rollout_mismatches_in_band(O, osvc, g["beliefs"], g["controls"], valid, g[f"{tag}_valid"], max_count=3)
            assert np.allclose(fin[agree], g[f"{tag}_final"][agree], rtol=RTOL, atol=1e-12), tag


@pytest.mark.parametrize("path", FILES, ids=IDS)
def test_golden_validate_plan(K, ctx, path):
    g = np.load(path)
    dim, k = int(g["dim"]), int(g["k"])
    rad, p_ag = setup_ctx(K, ctx, g, K.SVC_BLACKMORE)
    trajs = [K.aos_to_soa(t) for t in trajs_of(g)]
    for kind, tag in ((K.PVC_BLACKMORE, "pbm"), (K.PVC_BOUNDINGBOX, "pbb"), (K.PVC_CHISQUARED, "pcs"), (K.PVC_ADAPTIVE_BLACKMORE, "pad")):
        p_arg = 0.9 if kind == K.PVC_CHISQUARED else p_ag          # demos/main.cpp:105-138
        K.install_pvc(ctx, kind, dim, [rad] * k, p_arg)
        ref = g[f"{tag}_conflicts"]
        run = ctx.validate_plan(trajs)
        assert run is not None, tag
        a, b, s0, n = run
        assert (a, b, s0) == tuple(int(x) for x in ref[0]), (tag, run, ref[0])
        assert n == len(ref), (tag, n, len(ref))
        assert np.array_equal(ref[:, 2], np.arange(s0, s0 + n))


def test_validate_plan_semantics_vs_oracle(K, O, ctx):
    """Earliest step, std::map pair order (k=12: "Robot 10" < "Robot 2"), waiting robots,
    run that reaches the last step, conflict-free plans."""
    rng = np.random.default_rng(5)
    k, dim = 12, 2
    world = O.World.synthetic(np.zeros((0, 2)), 40.0, 40.0, k, 0.9)
    rad = K.rect_robot_bounding_radius()
    ctx.set_model(K.MODEL_LINEAR2D)
    for trial in range(6):
        trajs = []
        for r in range(k):
            b0 = O.default_beliefs(1, dim)[0]
            ang = rng.uniform(0, 2 * np.pi)
            b0[:2] = [20 + 9 * np.cos(ang), 20 + 9 * np.sin(ang)]
            u = -np.array([np.cos(ang), np.sin(ang)]) * rng.uniform(0.6, 1.0)
            n = int(rng.integers(5, 70)) if trial < 5 else 3
            trajs.append(np.concatenate([b0[None, :], O.propagate_n(b0, u, n, dim)]))
        for kind in (K.PVC_BLACKMORE, K.PVC_CHISQUARED, K.PVC_ADAPTIVE_BOUNDINGBOX):
            opvc = O.Pvc(world, kind)
            K.install_pvc(ctx, kind, dim, [rad] * k, 0.9 if kind == K.PVC_CHISQUARED else world.p_safe_agents)
            ref = opvc.validate_plan(trajs, dim)
            run = ctx.validate_plan([K.aos_to_soa(t) for t in trajs])
            if not ref:
                assert run is None
            else:
                assert run == (ref[0][0], ref[0][1], ref[0][2], len(ref)), (trial, kind, run, ref[0], len(ref))


def test_validate_plan_sharded_over_step_ranges(K, O, ctx):
    """the conflict search shards over time steps - every rank scans its own range of the replicated plan
    (cck_validate_plan_range) and the record with the smallest (step, pair-order) key is validatePlan's answer, follow-up run
    included (it is always taken over the whole plan)."""
    from cck_b200 import sharding
    rng = np.random.default_rng(11)
    k, dim = 12, 2
    world = O.World.synthetic(np.zeros((0, 2)), 40.0, 40.0, k, 0.9)
    rad = K.rect_robot_bounding_radius()
    ctx.set_model(K.MODEL_LINEAR2D)
    seen_conflict = seen_free = False
    for trial in range(5):
        trajs = []
        for r in range(k):
            b0 = O.default_beliefs(1, dim)[0]
            ang = rng.uniform(0, 2 * np.pi)
            far = 9.0 if trial < 4 else 1000.0 * (r + 1)           # last trial: nobody ever meets
            b0[:2] = [20 + far * np.cos(ang), 20 + far * np.sin(ang)]
            u = -np.array([np.cos(ang), np.sin(ang)]) * rng.uniform(0.6, 1.0)
            trajs.append(np.concatenate([b0[None, :], O.propagate_n(b0, u, int(rng.integers(5, 70)), dim)]))
        soa = [K.aos_to_soa(t) for t in trajs]
        T = max(len(t) for t in trajs)
        for kind in (K.PVC_BLACKMORE, K.PVC_CHISQUARED):
            K.install_pvc(ctx, kind, dim, [rad] * k, 0.9 if kind == K.PVC_CHISQUARED else world.p_safe_agents)
            ref = O.Pvc(world, kind).validate_plan(trajs, dim)
            want = None if not ref else (ref[0][0], ref[0][1], ref[0][2], len(ref))
            assert ctx.validate_plan(soa) == want
            for n_ranks in (1, 2, 3, 8):
                recs = [ctx.validate_plan_range(soa, *sharding.shard_bounds(T, r, n_ranks)) for r in range(n_ranks)]
                assert sharding.pick_conflict([x.tolist() for x in recs]) == want, (trial, kind, n_ranks)
                if want is not None:      # ranges that end before the first conflict report none; the key carries the step
                    for r, x in enumerate(recs):
                        lo, hi = sharding.shard_bounds(T, r, n_ranks)
                        assert (x[0] < 0) if hi <= want[2] else True
                        assert x[0] < 0 or (lo <= (x[0] >> 32) < hi and x[3] == (x[0] >> 32))
            seen_conflict |= want is not None
            seen_free |= want is None
    assert seen_conflict and seen_free


def test_satisfies_constraints_vs_oracle(K, O, ctx):
    rng = np.random.default_rng(9)
    k, dim = 3, 2
    world = O.World.synthetic(np.zeros((0, 2)), 20.0, 20.0, k, 0.9)
    rad = K.rect_robot_bounding_radius()
    ctx.set_model(K.MODEL_LINEAR2D)
    # constraining robots 1 and 2 cross the corridor x in [6, 14], y = 10 at different times
    cons = []
    for other, (t0, n) in ((1, (10, 8)), (2, (25, 6))):
        b0 = O.default_beliefs(1, dim)[0]
        b0[:2] = [10.0, 4.0 + other]
        tr = np.concatenate([b0[None, :], O.propagate_n(b0, np.array([0.0, 0.9]), 60, dim)])
        steps = list(range(t0, t0 + n))
        cons.append((other, steps, tr[steps]))
    n_paths, T = 300, 45
    paths = np.zeros((n_paths, T, 10))
    plen = rng.integers(2, T + 1, size=n_paths).astype(np.int32)
    for i in range(n_paths):
        b0 = O.default_beliefs(1, dim)[0]
        b0[:2] = [rng.uniform(6, 14), rng.uniform(6, 14)]
        paths[i, 0] = b0
        paths[i, 1:] = O.propagate_n(b0, rng.uniform(-1, 1, 2), T - 1, dim)
    for kind in (K.PVC_BLACKMORE, K.PVC_BOUNDINGBOX, K.PVC_ADAPTIVE_BLACKMORE, K.PVC_CHISQUARED):
        opvc = O.Pvc(world, kind)
        K.install_pvc(ctx, kind, dim, [rad] * k, 0.9 if kind == K.PVC_CHISQUARED else world.p_safe_agents)
        ctx.set_constraints(*K.constraint_entries(k, dim, 0, cons))
        soa = np.ascontiguousarray(paths.transpose(1, 2, 0))           # [T, W, n]
        ok = ctx.satisfies_constraints(soa, plen, np.zeros(n_paths, dtype=np.int32))
        ref = np.array([opvc.satisfies_constraints(paths[i, :plen[i]], dim, 0, cons) for i in range(n_paths)])
        assert np.array_equal(ok, ref), kind
        # every kind runs its own isSafe_ at the constraint times (ChiSquaredBoundaryPVC.cpp:54-87 included)
        assert 0.05 < ref.mean() < 0.995, (kind, ref.mean())
    # no constraints installed: everything passes
    ctx.set_constraints([], [], np.zeros((0, 2)), np.zeros((0, 4)))
    assert ctx.satisfies_constraints(soa, plen, np.zeros(n_paths, dtype=np.int32)).all()


@pytest.mark.parametrize("model", ["linear", "unicycle"])
def test_expand_batch_vs_oracle(K, O, W, ctx, model):
    """One launch = propagateWhileValid + constraint check on the new segment + path-length
    cost + ChanceConstrainedGoal, for K candidates (ConstraintRespectingBSST.cpp:245-326)."""
    dim = 2 if model == "linear" else 4
    rng = np.random.default_rng(3)
    k = 2
    centres = W.synthetic_obstacles(40, seed=77)[:, :] % 20
    world = O.World.synthetic(centres, 20.0, 20.0, k, 0.9)
    svc = O.Svc(world, O.SVC_BLACKMORE, dim)
    opvc = O.Pvc(world, O.PVC_BLACKMORE)
    ctx.set_model(K.MODEL_LINEAR2D if dim == 2 else K.MODEL_UNICYCLE)
    ctx.set_obstacles(K.SVC_BLACKMORE, svc.low, svc.high, svc.A, svc.B, svc.centres, svc.rects, erf_inv_result=svc.erf_inv_result)
    rad = K.rect_robot_bounding_radius()
    K.install_pvc(ctx, K.PVC_BLACKMORE, dim, [rad] * k, world.p_safe_agents)
    n = 3000
    bel, ctrl = W.synthetic_beliefs(n, model, seed=13)
    bel[0] = rng.uniform(0, 19, n); bel[1] = rng.uniform(0, 19, n)
    bel[0, ::2] = rng.uniform(8.5, 11.5, (n + 1) // 2); bel[1, ::2] = rng.uniform(8.5, 11.5, (n + 1) // 2)   # near the constraint
    bel[dim:] *= 0.02
    if dim == 4:
        ctrl[0] = rng.uniform(-1, 20, n); ctrl[1] = rng.uniform(-1, 20, n)
    steps = rng.integers(1, 11, size=n).astype(np.int32)
    pstep = rng.integers(0, 30, size=n).astype(np.int32)
    pcost = rng.uniform(0, 5, n)
    # a constraint: the other robot sits near (10, 10) during steps 8..24
    other = O.default_beliefs(1, dim)[0]; other[:2] = [10.0, 10.0]
    if dim == 4:
        other[3] = 0.1
    csteps = list(range(8, 25))
    cons = [(1, csteps, np.repeat(other[None, :], len(csteps), axis=0))]
    ctx.set_constraints(*K.constraint_entries(k, dim, 0, cons))
    goal = (12.0, 12.0)
    out, valid, cost, gd, fl = ctx.expand_batch(bel, ctrl, steps, pstep, pcost, goal)
    aos_b, aos_c = K.soa_to_aos(bel), K.soa_to_aos(ctrl)
    ref_out, ref_valid, _, traj = svc.rollout(aos_b, aos_c, steps, traj_stride=10)
    if dim == 2:
        assert np.array_equal(valid, ref_valid)
    good = assert_rollout_mismatches_in_band(O, svc, aos_b, aos_c, valid, ref_valid)
    if dim == 2:
        assert np.array_equal(K.soa_to_aos(out), ref_out)
    else:
        assert np.allclose(K.soa_to_aos(out)[good], ref_out[good], rtol=RTOL, atol=1e-12)
    ref_cost = pcost + np.sqrt(((aos_b[:, :dim] - ref_out[:, :dim]) ** 2).sum(axis=1))
    assert np.allclose(cost[good], ref_cost[good], rtol=1e-12)
    g_ok, g_d = O.goal(ref_out, dim, goal[0], goal[1])
    assert np.allclose(gd[good], g_d[good], rtol=1e-12, atol=1e-12)
    # goal flags: exact, except where the goal test itself is inside the band (the goal decagon's radius moved by +-1e-6
    # flips the oracle's verdict) - the unicycle's final mean carries the device sincos ulps
    g_bad = (((fl & K.FLAG_GOAL) != 0) != g_ok) & good
    if g_bad.any():
        flips = np.zeros(n, dtype=bool)
        for thr in (2.0 - 1e-6, 2.0 + 1e-6):
            flips |= O.goal(ref_out, dim, goal[0], goal[1], threshold=thr)[0] != g_ok
        assert not (g_bad & ~flips).any()
    assert g_ok.any() and (dim == 4 or not g_bad.any())
    assert np.array_equal(((fl & K.FLAG_FULL) != 0)[good], (ref_valid == steps)[good])
    # constraints on the new segment: states at absolute steps pstep+1 .. pstep+valid
    ref_c = np.ones(n, bool)
    for i in np.nonzero(good)[0]:
        v = ref_valid[i]
        if v == 0:
            continue
        # a path whose states carry absolute steps: pad the front with far-away states
        ref_c[i] = _segment_ok(opvc, traj[i, :v], int(pstep[i]) + 1, dim, cons)
    got_c = (fl & K.FLAG_CONS_OK) != 0
    assert np.array_equal(got_c[good], ref_c[good])
    assert 0.01 < (~ref_c).mean() < 0.9


def _segment_ok(opvc, seg, first_step, dim, cons):
    far = seg[:1].copy(); far[0, :2] = [-500.0, -500.0]
    path = np.concatenate([np.repeat(far, first_step, axis=0), seg])
    return opvc.satisfies_constraints(path, dim, 0, cons)


@pytest.mark.parametrize("model", ["linear", "unicycle"])
def test_expand_batch_cov_table_is_bit_identical(K, O, W, ctx, model):
    """cck_set_cov_table: a candidate whose parent covariance is the tree's covariance at its depth reads its next
    covariances from the table instead of running the Kalman step - same bits as computing them (the
    recursion is control-independent, UncertainLinearSP.cpp:98-104); any other candidate is computed as usual."""
    dim = 2 if model == "linear" else 4
    Wd = K.width(dim)
    rng = np.random.default_rng(8)
    centres = W.synthetic_obstacles(40, seed=77)[:, :] % 20
    world = O.World.synthetic(centres, 20.0, 20.0, 2, 0.9)
    svc = O.Svc(world, O.SVC_BLACKMORE, dim)
    ctx.set_model(K.MODEL_LINEAR2D if dim == 2 else K.MODEL_UNICYCLE)
    ctx.set_obstacles(K.SVC_BLACKMORE, svc.low, svc.high, svc.A, svc.B, svc.centres, svc.rects, erf_inv_result=svc.erf_inv_result)
    # the tree's covariances by depth, from the DEVICE's own single-step propagator (what a planner's nodes carry)
    depth_max = 40
    cur = K.aos_to_soa(O.default_beliefs(1, dim))
    if dim == 4:
        cur[3] = 0.1
    zero_u = np.zeros((dim, 1))
    if dim == 4:
        zero_u[:, 0] = [5.0, 5.0, 0.0, 0.5]
    covs = [cur[dim:, 0].copy()]
    for _ in range(depth_max + 12):
        cur = ctx.propagate(cur, zero_u)
        covs.append(cur[dim:, 0].copy())
    n = 2000
    bel, ctrl = W.synthetic_beliefs(n, model, seed=14)
    bel[0] = rng.uniform(1, 18, n); bel[1] = rng.uniform(1, 18, n)
    pstep = rng.integers(0, depth_max, size=n).astype(np.int32)
    on_tree = rng.uniform(size=n) < 0.7
    for i in range(n):
        if on_tree[i]:
            bel[dim:, i] = covs[pstep[i]]
        else:
            bel[dim:, i] *= 0.02
    if dim == 4:
        ctrl[0] = rng.uniform(-1, 20, n); ctrl[1] = rng.uniform(-1, 20, n)
    steps = rng.integers(0, 11, size=n).astype(np.int32)
    pcost = rng.uniform(0, 5, n)
    goal = (12.0, 12.0)
    ref = ctx.expand_batch(bel, ctrl, steps, pstep, pcost, goal)
    assert not (ref[4] & K.FLAG_COV_TABLE).any()
    ctx.set_cov_table(covs[0], depth_max + 11)
    got = ctx.expand_batch(bel, ctrl, steps, pstep, pcost, goal)
    used = (got[4] & K.FLAG_COV_TABLE) != 0
    assert np.array_equal(used, on_tree & (steps > 0)), "the table must be used exactly for the on-tree candidates"
    for a_, b_ in zip(ref[:4], got[:4]):
        assert np.array_equal(a_.view(np.uint64) if a_.dtype == np.float64 else a_, b_.view(np.uint64) if b_.dtype == np.float64 else b_)
    assert np.array_equal(ref[4], got[4] & 7)
    # a segment that would run past the end of the table is computed, not read
    ctx.set_cov_table(covs[0], 5)
    short = ctx.expand_batch(bel, ctrl, steps, pstep, pcost, goal)
    assert np.array_equal((short[4] & K.FLAG_COV_TABLE) != 0, on_tree & (steps > 0) & (pstep + steps < 5))
    assert np.array_equal(short[0].view(np.uint64), ref[0].view(np.uint64))
    ctx.set_cov_table(None, 0)
    assert not (ctx.expand_batch(bel, ctrl, steps, pstep, pcost, goal)[4] & K.FLAG_COV_TABLE).any()


def test_chisq_pairs_disjoint_vs_oracle(K, O, ctx):
    """CentralizedChiSquaredBoundarySVC's agent-agent decagon test (CentralizedChiSquaredBoundarySVC.cpp:70-99): bit-exact
    against the oracle, including pairs a few ulps from tangency."""
    rng = np.random.default_rng(21)
    n = 20000
    mu_a = rng.uniform(0, 8, (n, 2))
    L = rng.uniform(0.02, 0.3, (2, n, 3))
    cov = [np.stack([l[:, 0] ** 2, l[:, 0] * l[:, 1], l[:, 0] * l[:, 1], l[:, 1] ** 2 + l[:, 2] ** 2], axis=1) for l in L]
    rad = rng.uniform(0.1, 0.3, (2, n))
    sc = K.chi2_quantile_2dof(0.9)
    lam = [0.5 * (c[:, 0] + c[:, 3]) + np.sqrt((0.5 * (c[:, 0] + c[:, 3])) ** 2 - (c[:, 0] * c[:, 3] - c[:, 1] * c[:, 2])) for c in cov]
    ra, rb = lam[0] * sc + rad[0], lam[1] * sc + rad[1]
    ang = rng.uniform(0, 2 * np.pi, n)
    # distances spread around the tangency range, a third right on a vertex-vertex contact
    d = (ra + rb) * rng.uniform(0.85, 1.1, n)
    k = np.arange(0, n, 3)
    ang[k] = rng.integers(0, 10, len(k)) * (2 * np.pi / 10)
    d[k] = (ra + rb)[k] * (1 + rng.uniform(-3e-15, 3e-15, len(k)))
    mu_b = mu_a + np.stack([d * np.cos(ang), d * np.sin(ang)], axis=1)
    # degenerate covariances (NaN / inf / hugely negative entries): non-finite decagons; the oracle's separating-axis test never
    # takes a NaN projection, the kernel's must not either
    bad = np.arange(5, n, 7)
    for j, i in enumerate(bad):
        cov[j % 2][i, j % 4] = [np.nan, np.inf, -np.inf, -1e300, 1e300][j % 5]
    got = ctx.chisq_pairs_disjoint(mu_a.T, cov[0].T, rad[0], mu_b.T, cov[1].T, rad[1], sc)
    ref = O.chisq_pairs_disjoint(mu_a, cov[0], rad[0], mu_b, cov[1], rad[1], sc)
    assert np.array_equal(got, ref)
    assert 0.2 < ref.mean() < 0.8
