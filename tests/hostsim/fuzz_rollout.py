#!/usr/bin/env python
"""TEST INFRASTRUCTURE: randomised parity fuzzing of the linear + PCCBlackmore rollout kernel (clearance field, fp32 broad phase,
exact decision, lane refill) against the oracle, on the CPU build of the CUDA sources (tests/hostsim) - or on a GPU when one is
there (leave CCK_SO unset).  Every case draws a world (size, obstacle count, map-like integer cells or arbitrary / rotated /
degenerate half-plane tables, risk level, state bounds), beliefs in several regimes (uniform; means placed within 1e-12 .. 1e-3
of an obstacle's decision threshold; tiny / huge / degenerate covariances; means at and outside the bounds) and per-belief step
counts, and demands bit-identical isValid verdicts, first-violation steps and final beliefs.

  python tests/hostsim/fuzz_rollout.py [--minutes 10] [--seed 1] [--n 4096]
Exit code 1 and a reproducer line on the first mismatch.
"""
import argparse
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def random_tables(rng, M, mode, world):
    A, B = np.zeros((M, 4, 2)), np.zeros((M, 4))
    for o in range(M):
        cx, cy = rng.uniform(0, world - 1, 2)
        hw, hh = rng.uniform(0.2, 1.5, 2)
        th = 0.0 if mode == "axis" or (mode == "mixed" and o % 2 == 0) else rng.uniform(0, np.pi)
        w = rng.uniform(0.5, 3.0)
        ct, st = (1.0, 0.0) if th == 0.0 else (np.cos(th), np.sin(th))
        normals = [(st, -ct), (ct, st), (-st, ct), (-ct, -st)]
        ext = [hh, hw, hh, hw]
        for i, ((nx, ny), e) in enumerate(zip(normals, ext)):
            A[o, i] = (w * nx, w * ny)
            B[o, i] = w * (nx * cx + ny * cy + e)
    return A, B


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--minutes", type=float, default=10.0)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--n", type=int, default=4096)
    ap.add_argument("--cases", type=int, default=0, help="stop after this many cases (0: by time)")
    ap.add_argument("--generic", action="store_true", help="fuzz the other kernels instead: unicycle + Blackmore / chi-squared, linear + "
                    "chi-squared (k_rollout_uni2, k_rollout_lin<CHISQ>); bit-identical verdicts and steps, values bit-identical on the CPU "
                    "build (same libm) and 1e-9 relative on a GPU (device sincos)")
    ap.add_argument("--pvc", action="store_true", help="fuzz the plan checkers instead: validatePlan and satisfiesConstraints of all seven "
                    "kinds on random ragged plans with degenerate states mixed in")
    ap.add_argument("--tree", action="store_true", help="fuzz the belief-space distance, selectNode and nearest-witness queries instead "
                    "(non-symmetric, near-singular, negative-definite and non-finite covariances mixed in)")
    ap.add_argument("--expand", action="store_true", help="fuzz cck_expand_batch instead: rollout + constraint check on the new segment + cost + "
                    "goal test in one launch, both models, constraints of the exact plan-checker kinds, with and without the covariance table")
    ap.add_argument("--witness", action="store_true", help="fuzz cck_tree_witness_batch instead: the witness rules of one launch (nearest "
                    "existing witness, which survivors become witnesses, nearest same-launch witness) against a sequential replay")
    a = ap.parse_args()
    if os.environ.get("CCK_HOSTSIM", "") == "1" and not os.environ.get("CCK_SO"):
        sys.path.insert(0, HERE)
        import build as hostsim_build
        os.environ["CCK_SO"] = hostsim_build.build()
    import cck_b200 as K
    from oracle import oracle as O
    O.build()
    ctx = K.Context(0)
    t_end = time.time() + 60.0 * a.minutes
    if a.generic:
        return fuzz_generic(a, K, O, ctx, t_end)
    if a.pvc:
        return fuzz_pvc(a, K, O, ctx, t_end)
    if a.tree:
        return fuzz_tree(a, K, O, ctx, t_end)
    if a.expand:
        return fuzz_expand(a, K, O, ctx, t_end)
    if a.witness:
        return fuzz_witness(a, K, O, ctx, t_end)
    case = 0
    total_beliefs = total_steps = died = 0
    while (a.cases and case < a.cases) or (not a.cases and time.time() < t_end):
        seed = a.seed * 1000003 + case
        rng = np.random.default_rng(seed)
        world = int(rng.choice([8, 16, 32, 64, 200]))
        M = int(rng.choice([0, 1, 2, 7, 40, 102, 256, 600]))
        M = min(M, world * world // 2)
        k_agents = int(rng.integers(1, 9))
        p_safe = float(rng.choice([0.6, 0.8, 0.9, 0.95, 0.99, 0.999]))
        cells = rng.choice(world * world, size=M, replace=False)
        centres = np.stack([cells % world, cells // world], axis=1).astype(np.float64)
        wld = O.World.synthetic(centres, float(world), float(world), k_agents, p_safe)
        svc = O.Svc(wld, O.SVC_BLACKMORE, 2)
        mode = str(rng.choice(["map", "map", "map", "axis", "rotated", "mixed"])) if M > 0 else "map"
        if mode != "map":
            A, B = random_tables(rng, M, mode, world)
            if rng.uniform() < 0.3 and M >= 8:
                B[0, 1] = np.inf; B[1, 0] = -np.inf; B[2, 2] = np.nan; A[3, 3] = [0.0, 0.0]
                A[4] *= 1e150; B[4] *= 1e150
                A[5:7], B[5:7] = A[6:8].copy(), B[6:8].copy()
            svc.override_tables(A, B, float(rng.choice([svc.erf_inv_result, 0.5, 2.2, 3.5])))
        ctx.set_model(K.MODEL_LINEAR2D)
        ctx.set_obstacles(svc.kind, svc.low, svc.high, svc.A, svc.B, svc.centres, svc.rects, erf_inv_result=svc.erf_inv_result,
                          p_collision=svc.p_collision, sc=svc.sc, max_rad=svc.max_rad)
        n = a.n
        aos = O.default_beliefs(n, 2)
        aos[:, 0] = rng.uniform(-1.5, world + 0.5, n)
        aos[:, 1] = rng.uniform(-1.5, world + 0.5, n)
        scale = 10.0 ** rng.uniform(-4, 0.3, n)
        for blk in (2, 6):
            L = np.zeros((n, 2, 2))
            L[:, 0, 0] = rng.uniform(0.2, 1.0, n) * scale; L[:, 1, 1] = rng.uniform(0.2, 1.0, n) * scale
            L[:, 1, 0] = rng.uniform(-0.5, 0.5, n) * scale
            aos[:, blk:blk + 4] = np.einsum("nij,nkj->nik", L, L).reshape(n, 4)
        ctrl = rng.uniform(-1.0, 1.0, size=(n, 2)) * rng.choice([0.0, 0.01, 0.3, 1.0], size=(n, 1))
        steps = rng.integers(0, 51, n).astype(np.int32)
        # regime: means at an obstacle's threshold distance +- eps for the covariance the checker sees after ONE step
        if M > 0:
            m = n // 2
            post = O.propagate(aos[:m], ctrl[:m], 2)
            P = post[:, 2:6] + post[:, 6:10]
            ob = rng.integers(0, len(svc.A), m)
            pl = rng.integers(0, 4, m)
            a0, a1 = svc.A[ob, pl, 0], svc.A[ob, pl, 1]
            with np.errstate(all="ignore"):
                q = a0 * (a0 * P[:, 0] + a1 * P[:, 2]) + a1 * (a0 * P[:, 1] + a1 * P[:, 3])
                rhs = svc.B[ob, pl] + np.sqrt(2.0) * np.sqrt(q) * svc.erf_inv_result
                nn = a0 * a0 + a1 * a1
                eps = rng.choice([0.0, 1e-12, 1e-9, 1e-6, 1e-3], m) * rng.choice([-1.0, 1.0], m)
                # a point with a . mu = rhs + eps, displaced along the plane by a random amount, minus the step's own motion
                tpar = rng.uniform(-1.0, 1.0, m)
                px = a0 * (rhs + eps) / nn - a1 * tpar
                py = a1 * (rhs + eps) / nn + a0 * tpar
            ok = np.isfinite(px) & np.isfinite(py)
            aos[:m][ok, 0] = (px - 0.2 * ctrl[:m, 0])[ok]
            aos[:m][ok, 1] = (py - 0.2 * ctrl[:m, 1])[ok]
        # regime: degenerate / non-finite beliefs
        bad = rng.choice(n, size=max(1, n // 64), replace=False)
        vals = np.array([np.nan, np.inf, -np.inf, 0.0, -0.0, 1e300, -1e300, 1e-310, 5e-324])
        for i in bad:
            aos[i, int(rng.integers(0, 10))] = rng.choice(vals)
        # regime: means exactly on the state bounds
        onb = rng.choice(n, size=max(1, n // 64), replace=False)
        aos[onb, 0] = rng.choice([svc.low[0], svc.high[0], np.nextafter(svc.low[0], 1e9), np.nextafter(svc.high[0], -1e9)], len(onb))
        bel, ctl = K.aos_to_soa(aos), K.aos_to_soa(ctrl)
        with np.errstate(all="ignore"):
            v_got = ctx.is_valid(bel)
            v_ref = svc.is_valid(aos)[0]
            out, valid = ctx.propagate_while_valid(bel, ctl, steps)
            ref_out, ref_valid, _ = svc.rollout(aos, ctrl, steps, nthreads=8)
        got = K.soa_to_aos(out)
        same = np.array_equal(v_got, v_ref) and np.array_equal(valid, ref_valid) and np.array_equal(got, ref_out, equal_nan=True)
        if same and case % 4 == 0:      # trajectory mode (TRAJ kernels): every valid intermediate state, same rollout
            with np.errstate(all="ignore"):
                out2, valid2, tr = ctx.propagate_while_valid(bel, ctl, steps, traj_T=50)
                _, _, _, rtr = svc.rollout(aos, ctrl, steps, traj_stride=50)
            same = np.array_equal(valid2, ref_valid) and np.array_equal(K.soa_to_aos(out2), ref_out, equal_nan=True)
            tr = tr.transpose(2, 0, 1)                                   # [n, T, W]
            mask = np.arange(50)[None, :] < ref_valid[:, None]           # only the valid steps are defined
            same = same and np.array_equal(tr[mask], rtr[mask], equal_nan=True)
            if not same:
                print(f"MISMATCH (trajectory mode) case {case} seed {seed}")
        if not same:
            b1 = np.flatnonzero(v_got != v_ref)
            b2 = np.flatnonzero(valid != ref_valid)
            b3 = np.flatnonzero(~np.all((got == ref_out) | (np.isnan(got) & np.isnan(ref_out)), axis=1))
            print(f"MISMATCH case {case} seed {seed} world {world} M {M} mode {mode} p_safe {p_safe} k {k_agents}: "
                  f"is_valid {b1[:5]}, valid_steps {b2[:5]} (got {valid[b2[:5]]} ref {ref_valid[b2[:5]]}), finals {b3[:5]}")
            for i in list(b2[:3]) + list(b1[:3]):
                print("  belief", i, repr(aos[i].tolist()), "ctrl", repr(ctrl[i].tolist()), "steps", int(steps[i]))
            sys.exit(1)
        total_beliefs += n
        total_steps += int(ref_valid.sum())
        died += int((ref_valid < steps).sum())
        case += 1
        if case % 20 == 0:
            print(f"{case} cases, {total_beliefs} beliefs, {total_steps} valid belief-steps, {died} rollouts ended early: all bit-identical", flush=True)
    print(f"fuzz ok: {case} cases, {total_beliefs} beliefs, {total_steps} valid belief-steps, {died} rollouts ended early, 0 mismatches "
          f"(seed {a.seed}, library {K.capi._SO})")
    ctx.close()


def fuzz_witness(a, K, O, ctx, t_end):
    def beliefs(rng, n, dim, spread):
        b = np.zeros((n, K.width(dim)))
        b[:, :dim] = rng.uniform(0, spread, (n, dim))
        for off in (dim, dim + dim * dim):
            L = np.zeros((n, dim, dim))
            idx = np.tril_indices(dim)
            L[:, idx[0], idx[1]] = rng.uniform(0.05, 0.3, (n, len(idx[0])))
            b[:, off:off + dim * dim] = np.einsum("nij,nkj->nik", L, L).reshape(n, dim * dim)
        return b

    case = surv = new = adopt = 0
    while (a.cases and case < a.cases) or (not a.cases and time.time() < t_end):
        seed = a.seed * 1000003 + case
        rng = np.random.default_rng(seed)
        dim = int(rng.choice([2, 4]))
        spread = float(rng.choice([3.0, 10.0, 25.0]))
        pr = float(rng.choice([0.1, 0.5, 2.0]))
        nw = int(rng.choice([0, 1, 60, 400]))
        n = int(rng.choice([1, 33, 200, 700, 1024]))       # CCK_WIT_MAX = 1024 survivors per call
        wit = beliefs(rng, max(nw, 1), dim, spread)[:nw]
        b = beliefs(rng, n, dim, spread)
        if n > 50:
            b[10:30] = b[9]                                                       # identical survivors
            b[30:50, :dim] = b[29, :dim] + rng.normal(0, 0.3 * pr, (20, dim))    # a cluster around the pruning radius
        elig = (rng.random(n) < rng.uniform(0.2, 1.0)) if rng.uniform() < 0.5 else None
        tree = K.Tree(ctx, dim)
        if nw:
            tree.append(K.aos_to_soa(wit), np.zeros(nw))
        widx, wdist, near, ndist, fresh = tree.witness_batch(K.aos_to_soa(b), pr, elig)
        if nw:
            ridx, rdist = O.nearest_witness(wit, b, dim)
        else:
            ridx, rdist = np.full(n, -1, dtype=np.int32), np.full(n, np.inf)
        ok = np.array_equal(widx, ridx) and np.array_equal(wdist, rdist)
        created, why = [], ""
        for s_ in range(n):
            if not ok:
                why = "nearest existing witness"
                break
            cd, best = rdist[s_], -1
            if created:
                d = O.belief_distance(b[created], np.repeat(b[s_:s_ + 1], len(created), axis=0), dim)
                for f, df in zip(created, d):
                    if df < cd:
                        cd, best = df, f
            keeps = (cd > pr) and (elig is None or bool(elig[s_]))
            if bool(fresh[s_]) != keeps:
                ok, why = False, f"fresh[{s_}] got {bool(fresh[s_])} want {keeps}"
            elif best >= 0 and cd <= pr:
                if not (near[s_] == best and ndist[s_] == cd):
                    ok, why = False, f"near[{s_}] got {near[s_]} ({ndist[s_]}) want {best} ({cd})"
                adopt += 1
            elif near[s_] != -1:
                ok, why = False, f"near[{s_}] got {near[s_]} want -1"
            if keeps:
                created.append(s_)
        tree.close()
        if not ok:
            print(f"MISMATCH witness case {case} seed {seed} dim {dim} existing {nw} survivors {n} radius {pr} eligibility {elig is not None}: {why}")
            sys.exit(1)
        surv += n
        new += len(created)
        case += 1
        if case % 50 == 0:
            print(f"{case} cases, {surv} survivors, {new} new witnesses, {adopt} same-launch adoptions: all identical", flush=True)
    print(f"fuzz ok (witness batch): {case} cases, {surv} survivors, {new} new witnesses, {adopt} same-launch adoptions, 0 mismatches "
          f"(seed {a.seed}, library {K.capi._SO})")
    ctx.close()


def fuzz_expand(a, K, O, ctx, t_end):
    from cck_b200 import workloads as W
    sim = os.environ.get("CCK_HOSTSIM", "") == "1"
    rad = K.rect_robot_bounding_radius()
    kinds = [K.PVC_BLACKMORE, K.PVC_BOUNDINGBOX, K.PVC_CHISQUARED, K.PVC_BLACKMORE2] + ([K.PVC_CDFGRID] if sim else [])
    okind = {K.PVC_BLACKMORE: O.PVC_BLACKMORE, K.PVC_BOUNDINGBOX: O.PVC_BOUNDINGBOX, K.PVC_CHISQUARED: O.PVC_CHISQUARED,
             K.PVC_BLACKMORE2: O.PVC_BLACKMORE2, K.PVC_CDFGRID: O.PVC_CDFGRID}
    case = cands = 0
    while (a.cases and case < a.cases) or (not a.cases and time.time() < t_end):
        seed = a.seed * 1000003 + case
        rng = np.random.default_rng(seed)
        model = "linear" if case % 2 == 0 else "unicycle"
        dim = 2 if model == "linear" else 4
        if dim == 4 and not sim:      # device sincos: needs the band logic of the GPU tests, not a bit-exact fuzz
            case += 1
            continue
        k = int(rng.choice([2, 3, 5]))
        M = int(rng.choice([0, 7, 40]))
        cells = rng.choice(20 * 20, size=M, replace=False)
        centres = np.stack([cells % 20, cells // 20], axis=1).astype(np.float64)
        wld = O.World.synthetic(centres, 20.0, 20.0, k, float(rng.choice([0.8, 0.9, 0.99])))
        svc = O.Svc(wld, O.SVC_BLACKMORE, dim)
        kind = kinds[int(rng.integers(len(kinds)))]
        grid_n = int(rng.choice([2, 3, 5]))
        opvc = O.Pvc(wld, okind[kind], grid_n=grid_n) if kind == K.PVC_CDFGRID else O.Pvc(wld, okind[kind])
        ctx.set_model(K.MODEL_LINEAR2D if dim == 2 else K.MODEL_UNICYCLE)
        ctx.set_obstacles(K.SVC_BLACKMORE, svc.low, svc.high, svc.A, svc.B, svc.centres, svc.rects, erf_inv_result=svc.erf_inv_result)
        K.install_pvc(ctx, kind, dim, [rad] * k, 0.9 if kind == K.PVC_CHISQUARED else wld.p_safe_agents, grid_n=grid_n)
        n = 256
        bel, ctrl = W.synthetic_beliefs(n, model, seed=seed % 100000)
        bel[0] = rng.uniform(0, 19, n); bel[1] = rng.uniform(0, 19, n)
        bel[0, ::2] = rng.uniform(8.0, 12.0, (n + 1) // 2); bel[1, ::2] = rng.uniform(8.0, 12.0, (n + 1) // 2)
        bel[dim:] *= 10.0 ** rng.uniform(-3, -1)
        if dim == 4:
            ctrl[0] = rng.uniform(-1, 20, n); ctrl[1] = rng.uniform(-1, 20, n)
        steps = rng.integers(0, 11, size=n).astype(np.int32)
        pstep = rng.integers(0, 30, size=n).astype(np.int32)
        pcost = rng.uniform(0, 5, n)
        cons = []
        for other in range(1, k):
            ob = O.default_beliefs(1, dim)[0]
            ob[:2] = rng.uniform(8.5, 11.5, 2)
            if dim == 4:
                ob[3] = 0.1
            ob[dim:] *= 10.0 ** rng.uniform(-1, 0.5)
            st = sorted(set(int(x) for x in rng.integers(0, 40, 12)))
            cons.append((other, st, np.repeat(ob[None, :], len(st), axis=0)))
        ctx.set_constraints(*K.constraint_entries(k, dim, 0, cons))
        goal = (float(rng.uniform(5, 15)), float(rng.uniform(5, 15)))
        # half of the cases with the covariance table of a tree rooted at the default start covariance: candidates whose
        # parent covariance is the tree's get it (same bits), all others are computed
        use_tab = case % 4 >= 2
        if use_tab:
            cur = K.aos_to_soa(O.default_beliefs(1, dim))
            if dim == 4:
                cur[3] = 0.1
            zero_u = np.zeros((dim, 1))
            covs = [cur[dim:, 0].copy()]
            for _ in range(45):
                cur = ctx.propagate(cur, zero_u)
                covs.append(cur[dim:, 0].copy())
            for i in range(0, n, 2):
                bel[dim:, i] = covs[pstep[i]]
            ctx.set_cov_table(covs[0], 44)
        else:
            ctx.set_cov_table(None, 0)
        aos_b, aos_c = K.soa_to_aos(bel), K.soa_to_aos(ctrl)
        with np.errstate(all="ignore"):
            out, valid, cost, gd, fl = ctx.expand_batch(bel, ctrl, steps, pstep, pcost, goal)
            ref_out, ref_valid, _, traj = svc.rollout(aos_b, aos_c, steps, traj_stride=10)
            g_ok, g_d = O.goal(ref_out, dim, goal[0], goal[1])
        ref_cost = pcost + np.sqrt(((aos_b[:, :dim] - ref_out[:, :dim]) ** 2).sum(axis=1))
        ref_c = np.ones(n, bool)
        for i in range(n):
            v = int(ref_valid[i])
            if v:
                far = traj[i, :1].copy(); far[0, :2] = [-500.0, -500.0]
                path = np.concatenate([np.repeat(far, int(pstep[i]) + 1, axis=0), traj[i, :v]])
                ref_c[i] = opvc.satisfies_constraints(path, dim, 0, cons)
        problems = []
        if not np.array_equal(valid, ref_valid): problems.append(("valid_steps", np.flatnonzero(valid != ref_valid)))
        if not np.array_equal(K.soa_to_aos(out), ref_out, equal_nan=True): problems.append(("finals", np.flatnonzero((K.soa_to_aos(out) != ref_out).any(axis=1))))
        if not np.allclose(cost, ref_cost, rtol=1e-15, atol=0): problems.append(("cost", np.flatnonzero(cost != ref_cost)))
        if not np.array_equal(gd, g_d): problems.append(("goal_dist", np.flatnonzero(gd != g_d)))
        if not np.array_equal((fl & K.FLAG_GOAL) != 0, g_ok): problems.append(("goal flag", np.flatnonzero(((fl & K.FLAG_GOAL) != 0) != g_ok)))
        if not np.array_equal((fl & K.FLAG_FULL) != 0, ref_valid == steps): problems.append(("full flag", np.flatnonzero(((fl & K.FLAG_FULL) != 0) != (ref_valid == steps))))
        if not np.array_equal((fl & K.FLAG_CONS_OK) != 0, ref_c): problems.append(("constraint flag", np.flatnonzero(((fl & K.FLAG_CONS_OK) != 0) != ref_c)))
        if use_tab:
            want_tab = (np.arange(n) % 2 == 0) & (steps > 0) & (pstep + steps <= 44)
            if not np.array_equal((fl & K.FLAG_COV_TABLE) != 0, want_tab): problems.append(("cov-table flag", np.flatnonzero(((fl & K.FLAG_COV_TABLE) != 0) != want_tab)))
        if problems:
            print(f"MISMATCH expand case {case} seed {seed} {model} pvc kind {kind} grid_n {grid_n} k {k} M {M} table {use_tab}: "
                  + "; ".join(f"{name} at {idx[:5]}" for name, idx in problems))
            sys.exit(1)
        cands += n
        case += 1
        if case % 20 == 0:
            print(f"{case} cases, {cands} candidates: all identical", flush=True)
    print(f"fuzz ok (batched expansion): {case} cases, {cands} candidates, 0 mismatches (seed {a.seed}, library {K.capi._SO})")
    ctx.set_cov_table(None, 0)
    ctx.close()


def fuzz_tree(a, K, O, ctx, t_end):
    def beliefs(rng, n, dim, spread, nan_ok=True):
        b = np.zeros((n, K.width(dim)))
        b[:, :dim] = rng.uniform(0, spread, (n, dim))
        for off in (dim, dim + dim * dim):
            L = np.zeros((n, dim, dim))
            idx = np.tril_indices(dim)
            L[:, idx[0], idx[1]] = rng.uniform(0.02, 0.4, (n, len(idx[0])))
            b[:, off:off + dim * dim] = np.einsum("nij,nkj->nik", L, L).reshape(n, dim * dim)
        # degenerate covariances: zero, rank one, non-symmetric (the unicycle's), negative definite (operatorSqrt -> NaN), non-finite
        k = rng.choice(n, size=max(4, n // 8), replace=False)
        for j, i in enumerate(k):
            what = j % 6 if nan_ok else j % 3
            S = b[i, dim:dim + dim * dim].reshape(dim, dim)
            if what == 0: S[:] = 0.0
            elif what == 1: S[:] = np.outer(S[0], S[0])
            elif what == 2: S[0, 1] += 0.05
            elif what == 3: S[:] = -S
            elif what == 4: S[int(rng.integers(dim)), int(rng.integers(dim))] = rng.choice([np.nan, np.inf, -np.inf, 1e300, -1e300])
            else: b[i, int(rng.integers(0, dim))] = rng.choice([np.nan, np.inf, 1e300])
        return b

    case = pairs = queries = skipped = 0
    while (a.cases and case < a.cases) or (not a.cases and time.time() < t_end):
        seed = a.seed * 1000003 + case
        rng = np.random.default_rng(seed)
        dim = int(rng.choice([2, 4]))
        n = int(rng.choice([64, 300, 1500]))
        x, y = beliefs(rng, n, dim, 6.0), beliefs(rng, n, dim, 6.0)
        y[: n // 16, :dim] = x[: n // 16, :dim]
        with np.errstate(all="ignore"):
            got = K.belief_distance(ctx, K.aos_to_soa(x), K.aos_to_soa(y), dim)
            ref = O.belief_distance(x, y, dim)
        if not np.array_equal(got, ref, equal_nan=True):
            bad = np.flatnonzero(~((got == ref) | (np.isnan(got) & np.isnan(ref))))
            print(f"MISMATCH tree case {case} seed {seed} dim {dim}: distance at {bad[:5]} got {got[bad[:5]]} want {ref[bad[:5]]}")
            for i in bad[:2]:
                print("  a", repr(x[i].tolist()), "b", repr(y[i].tolist()))
            sys.exit(1)
        pairs += n
        # selectNode / nearest witness: a NaN distance makes a sequential scan order-dependent (the first item wins and is never
        # replaced) and is undefined in the reference's GNAT, so the tree holds the degenerate-but-comparable kinds only
        x = beliefs(rng, n, dim, 6.0, nan_ok=False)
        tree = K.Tree(ctx, dim)
        cost = np.round(rng.uniform(0, 20, n), 1)
        tree.append(K.aos_to_soa(x), cost)
        active = (rng.uniform(size=n) > 0.3).astype(np.uint8)
        off = np.nonzero(active == 0)[0].astype(np.int32)
        if len(off):
            tree.update(off, active=np.zeros(len(off), dtype=np.uint8))
        q = beliefs(rng, 64, dim, 6.0, nan_ok=False)
        q[:8, :dim] = x[:8, :dim]

        def undefined(j, select):
            """A query with a NaN distance to some (active) item: a sequential scan keeps the FIRST item when its distance is
            NaN (nothing compares below NaN) and the reference's GNAT is undefined - there is no answer to agree on.  Rank-
            deficient covariances get there through a slightly negative eigenvalue under operatorSqrt."""
            nonlocal skipped
            with np.errstate(all="ignore"):
                d = O.belief_distance(np.repeat(q[j:j + 1], n, axis=0), x, dim) if select else O.belief_distance(x, np.repeat(q[j:j + 1], n, axis=0), dim)
            u = bool(np.isnan(d[active == 1] if select else d).any())
            skipped += u
            return u
        for radius in (0.2, 1.5):
            with np.errstate(all="ignore"):
                idx, dist = tree.select(K.aos_to_soa(q), radius)
                ridx, rdist = O.select_nodes(x, cost, active, q, dim, radius)
            bad = [j for j in np.flatnonzero((idx != ridx) | ~((dist == rdist) | (np.isnan(dist) & np.isnan(rdist)))) if not undefined(j, True)]
            if bad:
                print(f"MISMATCH tree case {case} seed {seed} dim {dim} radius {radius}: select at {bad[:5]} got {idx[bad[:5]]} want {ridx[bad[:5]]}")
                sys.exit(1)
        with np.errstate(all="ignore"):
            idx, dist = tree.nearest(K.aos_to_soa(q))
            ridx, rdist = O.nearest_witness(x, q, dim)
        bad = [j for j in np.flatnonzero((idx != ridx) | ~((dist == rdist) | (np.isnan(dist) & np.isnan(rdist)))) if not undefined(j, False)]
        if bad:
            print(f"MISMATCH tree case {case} seed {seed} dim {dim}: nearest at {bad[:5]} got {idx[bad[:5]]} want {ridx[bad[:5]]}")
            sys.exit(1)
        queries += 3 * 64
        tree.close()
        case += 1
        if case % 50 == 0:
            print(f"{case} cases, {pairs} distances, {queries} tree queries ({skipped} differing answers on queries with NaN distances, undefined): all others identical", flush=True)
    print(f"fuzz ok (distance / tree): {case} cases, {pairs} distances (bit-identical, NaN included), {queries} tree queries, 0 mismatches; "
          f"{skipped} differing answers on queries with NaN distances (undefined) (seed {a.seed}, library {K.capi._SO})")
    ctx.close()


def fuzz_pvc(a, K, O, ctx, t_end):
    sim = os.environ.get("CCK_HOSTSIM", "") == "1"
    # adaptive kinds call erfinv with data-dependent arguments: the CPU build's erfinv and a GPU's differ from the oracle's
    # Boost restatement in the last ulps, so only a GPU run / band test could judge them; CDFGrid uses erf (same libm on the CPU build)
    kinds = [K.PVC_BLACKMORE, K.PVC_BOUNDINGBOX, K.PVC_CHISQUARED, K.PVC_BLACKMORE2] + ([K.PVC_CDFGRID] if sim else [])
    okind = {K.PVC_BLACKMORE: O.PVC_BLACKMORE, K.PVC_BOUNDINGBOX: O.PVC_BOUNDINGBOX, K.PVC_CHISQUARED: O.PVC_CHISQUARED,
             K.PVC_BLACKMORE2: O.PVC_BLACKMORE2, K.PVC_CDFGRID: O.PVC_CDFGRID}
    rad = K.rect_robot_bounding_radius()
    case = plans = cons_checked = conflicts = 0
    while (a.cases and case < a.cases) or (not a.cases and time.time() < t_end):
        seed = a.seed * 1000003 + case
        rng = np.random.default_rng(seed)
        dim = int(rng.choice([2, 2, 4]))
        k = int(rng.choice([2, 3, 5, 8, 12]))
        wld = O.World.synthetic(np.zeros((0, 2)), 40.0, 40.0, k, float(rng.choice([0.8, 0.9, 0.99])))
        ctx.set_model(K.MODEL_LINEAR2D if dim == 2 else K.MODEL_UNICYCLE)
        trajs = []
        spread = float(rng.choice([3.0, 9.0, 14.0]))
        for r in range(k):
            b0 = O.default_beliefs(1, dim)[0]
            ang = rng.uniform(0, 2 * np.pi)
            b0[:2] = [20 + spread * np.cos(ang), 20 + spread * np.sin(ang)]
            if dim == 4:
                b0[2], b0[3] = ang + np.pi, 0.5
                u = np.array([20.0, 20.0, ang + np.pi, rng.uniform(0.3, 2.0)])
            else:
                u = -np.array([np.cos(ang), np.sin(ang)]) * rng.uniform(0.3, 1.0)
            b0[dim:] *= 10.0 ** rng.uniform(-2, 0.5)
            t = np.concatenate([b0[None, :], O.propagate_n(b0, u, int(rng.integers(1, 60)), dim)])
            if rng.uniform() < 0.15:      # a degenerate state somewhere in the plan
                t[int(rng.integers(0, len(t))), int(rng.integers(0, t.shape[1]))] = rng.choice([np.nan, np.inf, -np.inf, 1e300, -1e300, 0.0])
            trajs.append(t)
        soa = [K.aos_to_soa(t) for t in trajs]
        for kind in kinds:
            grid_n = int(rng.choice([2, 3, 5, 9]))
            opvc = O.Pvc(wld, okind[kind], grid_n=grid_n) if kind == K.PVC_CDFGRID else O.Pvc(wld, okind[kind])
            K.install_pvc(ctx, kind, dim, [rad] * k, 0.9 if kind == K.PVC_CHISQUARED else wld.p_safe_agents, grid_n=grid_n)
            with np.errstate(all="ignore"):
                ref = opvc.validate_plan(trajs, dim)
                run = ctx.validate_plan(soa)
            want = None if not ref else (ref[0][0], ref[0][1], ref[0][2], len(ref))
            if run != want:
                print(f"MISMATCH pvc case {case} seed {seed} kind {kind} dim {dim} k {k} grid_n {grid_n}: validate_plan got {run} want {want}")
                sys.exit(1)
            conflicts += want is not None
            plans += 1
            # constraints: robot 0 against the states of the others at random steps; candidate paths = prefixes of robot 0's plan
            cons = []
            for other in range(1, min(k, 4)):
                st = sorted(set(int(x) for x in rng.integers(0, len(trajs[other]), 6)))
                cons.append((other, st, trajs[other][st]))
            ctx.set_constraints(*K.constraint_entries(k, dim, 0, cons))
            T = len(trajs[0])
            npth = min(T, 12)
            plen = rng.integers(1, T + 1, npth).astype(np.int32)
            paths = np.zeros((npth, T, trajs[0].shape[1]))
            for i in range(npth):
                paths[i, :plen[i]] = trajs[0][:plen[i]]
            with np.errstate(all="ignore"):
                ok = ctx.satisfies_constraints(np.ascontiguousarray(paths.transpose(1, 2, 0)), plen, np.zeros(npth, dtype=np.int32))
                refc = np.array([opvc.satisfies_constraints(paths[i, :plen[i]], dim, 0, cons) for i in range(npth)])
            if not np.array_equal(ok, refc):
                print(f"MISMATCH pvc case {case} seed {seed} kind {kind} dim {dim} k {k} grid_n {grid_n}: satisfies_constraints got {ok.astype(int)} want {refc.astype(int)}")
                sys.exit(1)
            cons_checked += npth
        case += 1
        if case % 50 == 0:
            print(f"{case} cases, {plans} plans validated ({conflicts} with a conflict), {cons_checked} constraint verdicts: all identical", flush=True)
    print(f"fuzz ok (plan checkers): {case} cases, {plans} plans ({conflicts} with a conflict), {cons_checked} constraint verdicts, 0 mismatches "
          f"(seed {a.seed}, library {K.capi._SO})")
    ctx.close()


def fuzz_generic(a, K, O, ctx, t_end):
    from cck_b200 import workloads as W
    sim = os.environ.get("CCK_HOSTSIM", "") == "1"
    case = total = 0
    while (a.cases and case < a.cases) or (not a.cases and time.time() < t_end):
        seed = a.seed * 1000003 + case
        rng = np.random.default_rng(seed)
        model, kind = [("unicycle", O.SVC_BLACKMORE), ("unicycle", O.SVC_CHISQUARED), ("linear", O.SVC_CHISQUARED),
                       ("linear", O.SVC_ADAPTIVE), ("unicycle", O.SVC_ADAPTIVE)][case % 5]
        dim = 2 if model == "linear" else 4
        world = int(rng.choice([8, 32, 64]))
        M = min(int(rng.choice([0, 7, 40, 102, 256])), world * world // 2)
        cells = rng.choice(world * world, size=M, replace=False)
        centres = np.stack([cells % world, cells // world], axis=1).astype(np.float64)
        wld = O.World.synthetic(centres, float(world), float(world), int(rng.integers(1, 9)), float(rng.choice([0.8, 0.9, 0.99])))
        svc = O.Svc(wld, kind, dim)
        ctx.set_model(K.MODEL_LINEAR2D if dim == 2 else K.MODEL_UNICYCLE)
        ctx.set_obstacles(svc.kind, svc.low, svc.high, svc.A, svc.B, svc.centres, svc.rects, erf_inv_result=svc.erf_inv_result,
                          p_collision=svc.p_collision, sc=svc.sc, max_rad=svc.max_rad)
        n = max(256, a.n // 4)
        bel, ctl = W.synthetic_beliefs(n, model, seed=seed % 100000, variant="realistic")
        bel[0] *= world / W.WORLD; bel[1] *= world / W.WORLD
        bel[dim:] *= 10.0 ** rng.uniform(-3, 0.5, n)
        if dim == 4:
            ctl[0] *= world / W.WORLD; ctl[1] *= world / W.WORLD
        bad = rng.choice(n, size=max(1, n // 64), replace=False)
        for i in bad:
            bel[int(rng.integers(0, bel.shape[0])), i] = rng.choice([np.nan, np.inf, -np.inf, 0.0, 1e300, -1e300])
        steps = rng.integers(0, 51, n).astype(np.int32)
        aos, ctrl = K.soa_to_aos(bel), K.soa_to_aos(ctl)
        with np.errstate(all="ignore"):
            out, valid = ctx.propagate_while_valid(bel, ctl, steps)
            ref_out, ref_valid, _ = svc.rollout(aos, ctrl, steps, nthreads=8)
        got = K.soa_to_aos(out)
        if kind == O.SVC_ADAPTIVE:
            # erf_inv at data-dependent arguments (Boost restatement in the oracle, erfinv in the kernels): a first-violation
            # step may differ only where the deciding state is inside the +-1e-6 band; everything else bit-identical
            import helpers
            with np.errstate(all="ignore"):
                try:
                    agree = helpers.assert_rollout_mismatches_in_band(O, svc, aos, ctrl, valid, ref_valid)
                    same = np.array_equal(got[agree], ref_out[agree], equal_nan=True) if (sim or dim == 2) else \
                        np.allclose(got[agree], ref_out[agree], rtol=1e-9, atol=1e-12, equal_nan=True)
                    same = same and agree.mean() > 0.995
                except AssertionError as e:
                    print("  ", e)
                    same = False
        elif sim or dim == 2:
            same = np.array_equal(valid, ref_valid) and np.array_equal(got, ref_out, equal_nan=True)
        else:
            agree = valid == ref_valid
            same = agree.mean() > 0.999 and np.allclose(got[agree], ref_out[agree], rtol=1e-9, atol=1e-12, equal_nan=True)
        if not same:
            b2 = np.flatnonzero(valid != ref_valid)
            b3 = np.flatnonzero(~np.all((got == ref_out) | (np.isnan(got) & np.isnan(ref_out)), axis=1))
            print(f"MISMATCH generic case {case} seed {seed} {model} kind {kind} world {world} M {M}: valid_steps {b2[:5]} "
                  f"(got {valid[b2[:5]]} ref {ref_valid[b2[:5]]}), finals {b3[:5]}")
            for i in list(b2[:3]) + list(b3[:2]):
                print("  belief", i, repr(aos[i].tolist()), "ctrl", repr(ctrl[i].tolist()), "steps", int(steps[i]))
            sys.exit(1)
        total += n
        case += 1
        if case % 30 == 0:
            print(f"{case} cases, {total} beliefs: all identical", flush=True)
    print(f"fuzz ok (generic kernels): {case} cases, {total} beliefs, 0 mismatches (seed {a.seed}, library {K.capi._SO})")
    ctx.close()


if __name__ == "__main__":
    main()
