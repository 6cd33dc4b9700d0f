#!/usr/bin/env python
"""Generate tests/golden/ref2_*.npz from oracle/_ref: the reference's OWN sources for the rows that round 1 left
unpinned - DynUnicycleControlSpace (UncertainUnicycleSP.cpp), RealVectorBeliefSpace::distance, ChanceConstrainedGoal,
ChiSquaredBoundarySVC, ChiSquaredBoundaryPVC, CDFGridPVC, Blackmore2PVC - compiled where they lie under /root/reference
against oracle/shim/ (see oracle/ref_capi2.cpp for what such a pin can and cannot prove).  Run where /root/reference is
mounted:

    python oracle/gen_ref_golden2.py

Seeded inputs (numpy Philox) and the reference's outputs are stored together; the CPU tests check the oracle against
them everywhere, the GPU tests check the CUDA path against them on the B200.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref as R  # noqa: E402


def rng_of(seed):
    return np.random.Generator(np.random.Philox(key=seed))


def spd(rng, n, dim, lo, hi):
    """L L^T with L lower triangular, entries U(lo, hi): SPD, non-isotropic, dense."""
    L = np.tril(rng.uniform(lo, hi, (n, dim, dim)))
    return (L @ np.transpose(L, (0, 2, 1))).reshape(n, dim * dim)


def beliefs(n, dim, x_max, y_max, seed, margin=1.5):
    """States over (and slightly beyond) the world; every 5th record is the allocState default 0.01 I, every 7th has a
    NON-symmetric Sigma (what the unicycle recursion produces, quirk 2)."""
    rng = rng_of(seed)
    W = dim + 2 * dim * dim
    b = np.zeros((n, W))
    b[:, 0] = rng.uniform(-margin, x_max + margin - 1.0, n)
    b[:, 1] = rng.uniform(-margin, y_max + margin - 1.0, n)
    if dim == 4:
        b[:, 2] = rng.uniform(-np.pi, np.pi, n)
        b[:, 3] = rng.uniform(0.02, 3.0, n)
    b[:, dim:dim + dim * dim] = spd(rng, n, dim, 0.01, 0.15)
    b[:, dim + dim * dim:] = spd(rng, n, dim, 0.01, 0.15)
    eye = 0.01 * np.eye(dim).reshape(-1)
    b[::5, dim:dim + dim * dim] = eye
    b[::5, dim + dim * dim:] = eye
    if dim == 4:
        b[3::7, dim + 1] += 0.002          # Sigma(0,1) != Sigma(1,0)
        b[3::7, dim + 2 * 4 + 3] += 0.001  # Sigma(2,3) != Sigma(3,2)
    if dim == 4:
        u = np.stack([rng.uniform(-1, x_max, n), rng.uniform(-1, y_max, n), rng.uniform(-np.pi, np.pi, n), rng.uniform(0.01, 10.0, n)], axis=1)
    else:
        u = rng.uniform(-1.0, 1.0, (n, 2))
    steps = rng.integers(0, 11, n).astype(np.int32)
    return b, u, steps


def random_plans(k, n_plans, dim, seed, arena=2.5, sig=(0.001, 0.03)):
    """Ragged random-walk trajectories in a small arena (conflicts are common)."""
    rng = rng_of(seed)
    W = dim + 2 * dim * dim
    lens, states = [], []
    for _ in range(n_plans):
        for _r in range(k):
            T = int(rng.integers(3, 40))
            t = np.zeros((T, W))
            t[:, :2] = np.cumsum(rng.uniform(-0.2, 0.2, (T, 2)), axis=0) + rng.uniform(0, arena, 2)
            if dim == 4:
                t[:, 2] = rng.uniform(-3, 3); t[:, 3] = rng.uniform(0.05, 1.0)
            S = spd(rng, 1, dim, np.sqrt(sig[0]), np.sqrt(sig[1]))[0]
            Lm = spd(rng, 1, dim, np.sqrt(sig[0]), np.sqrt(sig[1]))[0]
            t[:, dim:dim + dim * dim] = S
            t[:, dim + dim * dim:] = Lm
            lens.append(T); states.append(t)
    return np.array(lens, dtype=np.int32).reshape(n_plans, k), np.concatenate(states, axis=0)


def split(lens_row, states, off):
    trajs = []
    for L in lens_row:
        trajs.append(states[off:off + L]); off += L
    return trajs, off


def pvc_block(world, kind, dim, k, seed, n_plans, n_steps=2, tag=""):
    """validatePlan + satisfiesConstraints of one reference plan checker on random plans."""
    lens, states = random_plans(k, n_plans, dim, seed)
    W = dim + 2 * dim * dim
    rng = rng_of(seed + 1)
    n_probe = 3
    probe_len = rng.integers(2, 45, (n_plans, n_probe)).astype(np.int32)
    probes = np.zeros((n_plans, n_probe, 45, W))
    probes[..., :2] = np.cumsum(rng.uniform(-0.2, 0.2, (n_plans, n_probe, 45, 2)), axis=2) + rng.uniform(0, 2.5, (n_plans, n_probe, 1, 2))
    eye = 0.01 * np.eye(dim).reshape(-1)
    probes[..., dim:dim + dim * dim] = eye
    probes[..., dim + dim * dim:] = eye
    pv = R.Pvc2(world, kind, dim, n_steps=n_steps)
    confs, sat = [], np.full((n_plans, n_probe), -1, dtype=np.int8)
    off = 0
    for i in range(n_plans):
        trajs, off = split(lens[i], states, off)
        c = pv.validate_plan(trajs)
        confs.append(np.array([[i, a, b, s] for a, b, s in c], dtype=np.int32).reshape(-1, 4))
        if c:
            a, other = c[0][0], c[0][1]
            steps = [x[2] for x in c]
            cst = np.array([trajs[other][min(s, len(trajs[other]) - 1)] for s in steps])      # BeliefPVC::createConstraint (quirk 8)
            for j in range(n_probe):
                sat[i, j] = pv.satisfies_constraints(probes[i, j, :probe_len[i, j]], a, [(other, steps, cst)])
    return {f"{tag}plan_lens": lens, f"{tag}plan_states": states, f"{tag}probe_len": probe_len, f"{tag}probe_states": probes,
            f"{tag}conflicts": np.concatenate(confs, axis=0), f"{tag}satisfies": sat,
            f"{tag}info": np.array([pv.p_coll_agnts, pv.sc, pv.p_coll_dist])}


def world_fields(w, mp, sc, k):
    return dict(map=mp, scen=sc, k=k, p_safe=0.9, n_obstacles=w.n_obstacles, x_max=w.x_max, y_max=w.y_max, centres=w.centres,
                starts=w.starts, goals=w.goals, bounding_rad=w.bounding_rad, p_safe_obs=w.p_safe_obs, p_safe_agents=w.p_safe_agents)


def main():
    assert R.build(force=True), "needs /root/reference"
    out_dir = os.path.join(ROOT, "tests", "golden")
    REF = R.REFERENCE

    # ---- config 3: random-32-32-5, Uncertain-Unicycle, k = 4 ----
    mp, sc, k = "random-32-32-5", "random-32-32-5-Rectangles-Uncertain-Unicycle", 4
    w = R.World(f"{REF}/maps/{mp}.map", f"{REF}/scens/{sc}.scen", k, 0.9)
    out = world_fields(w, mp, sc, k)
    s_bm = R.Svc2(w, "unicycle", R.SVC2_BLACKMORE)
    F, A, Q, Rm = s_bm.uni_params()
    out.update(uni_F=F, uni_Acl=A, uni_Q=Q, uni_R=Rm)
    b, u, steps = beliefs(2500, 4, w.x_max, w.y_max, 20240503)
    out.update(beliefs=b, controls=u, steps=steps, propagate=s_bm.propagate(b, u), propagate_d01=s_bm.propagate(b[:300], u[:300], duration=0.1),
               bm_is_valid=s_bm.is_valid(b))
    fin, nval = s_bm.rollout(b, u, steps)
    out.update(bm_rollout_final=fin, bm_rollout_valid=nval)
    # 50-step chains from the start state of every robot (the planner's regime: Sigma0 = Lambda0 = 0.01 I)
    rng = rng_of(99)
    chain0 = np.zeros((k, 36)); chain0[:, :2] = w.starts; chain0[:, 2] = 0.0; chain0[:, 3] = 0.1
    chain0[:, 4:20] = 0.01 * np.eye(4).reshape(-1); chain0[:, 20:] = 0.01 * np.eye(4).reshape(-1)
    chain_u = np.stack([rng.uniform(0, 31, k), rng.uniform(0, 31, k), rng.uniform(-3, 3, k), rng.uniform(0.1, 2, k)], axis=1)
    chain, cur = [], chain0
    for _ in range(50):
        cur = s_bm.propagate(cur, chain_u); chain.append(cur)
    out.update(chain0=chain0, chain_u=chain_u, chain=np.stack(chain, axis=0))
    s_cs = R.Svc2(w, "unicycle", R.SVC2_CHISQUARED)
    ci = s_cs.chisq_info()
    out.update(cs_is_valid=s_cs.is_valid(b), cs_info=np.array([ci["sc"], ci["max_rad"], ci["n_obstacles"]]))
    fin, nval = s_cs.rollout(b[:800], u[:800], steps[:800])
    out.update(cs_rollout_final=fin, cs_rollout_valid=nval)
    s_ad = R.Svc2(w, "unicycle", R.SVC2_ADAPTIVE)
    out.update(ad_is_valid=s_ad.is_valid(b[:500]))
    # belief-space distance, dim 4 (pairs of the same batch, shifted) and goal test around robot 0's goal
    out.update(dist_a=b[:600], dist_b=b[600:1200], dist=R.distance(4, b[:600], b[600:1200]))
    gb, _, _ = beliefs(1500, 4, 3.0, 3.0, 4242, margin=0.0)
    gx, gy = w.goals[0]
    gb[:, 0] = gx + (gb[:, 0] - 1.0) * 1.6; gb[:, 1] = gy + (gb[:, 1] - 1.0) * 1.6
    gb[:, 4:] *= rng_of(7).uniform(0.05, 3.0, (1500, 1))
    ok, gd = R.goal(4, gb, gx, gy)
    out.update(goal_beliefs=gb, goal_xy=np.array([gx, gy]), goal_ok=ok, goal_dist=gd)
    out.update(pvc_block(w, R.PVC2_CHISQUARED, 4, k, 900, 40, tag="pcs_"))
    out.update(pvc_block(w, R.PVC2_MINKOWSKI, 4, k, 901, 40, tag="pbm_"))
    np.savez_compressed(os.path.join(out_dir, "ref2_config3_random32_5_unicycle_k4.npz"), **out)
    print("config3: M", w.n_obstacles, "bm valid %.2f" % out["bm_is_valid"].mean(), "bm full %.2f" % (out["bm_rollout_valid"] == steps).mean(),
          "cs valid %.2f" % out["cs_is_valid"].mean(), "goal ok %.2f" % ok.mean(), "pcs conflicts", len(out["pcs_conflicts"]), "pbm conflicts", len(out["pbm_conflicts"]))

    # ---- config 2 (linear, M = 102, k = 4): chi-squared SVC / PVC, distance, goal, CDFGrid, Blackmore2 ----
    mp, sc, k = "random-32-32-10", "random-32-32-10-Rectangles-2DUncertainLinear", 4
    w = R.World(f"{REF}/maps/{mp}.map", f"{REF}/scens/{sc}.scen", k, 0.9)
    out = world_fields(w, mp, sc, k)
    b, u, steps = beliefs(3000, 2, w.x_max, w.y_max, 20240504)
    s_cs = R.Svc2(w, "linear", R.SVC2_CHISQUARED)
    ci = s_cs.chisq_info()
    out.update(beliefs=b, controls=u, steps=steps, cs_is_valid=s_cs.is_valid(b), cs_info=np.array([ci["sc"], ci["max_rad"], ci["n_obstacles"]]))
    fin, nval = s_cs.rollout(b, u, steps)
    out.update(cs_rollout_final=fin, cs_rollout_valid=nval)
    out.update(dist_a=b[:1000], dist_b=b[1000:2000], dist=R.distance(2, b[:1000], b[1000:2000]))
    gb, _, _ = beliefs(2000, 2, 3.0, 3.0, 4243, margin=0.0)
    gx, gy = w.goals[1]
    gb[:, 0] = gx + (gb[:, 0] - 1.0) * 1.6; gb[:, 1] = gy + (gb[:, 1] - 1.0) * 1.6
    gb[:, 2:] *= rng_of(8).uniform(0.05, 3.0, (2000, 1))
    ok, gd = R.goal(2, gb, gx, gy)
    out.update(goal_beliefs=gb, goal_xy=np.array([gx, gy]), goal_ok=ok, goal_dist=gd)
    out.update(pvc_block(w, R.PVC2_CHISQUARED, 2, k, 910, 60, tag="pcs_"))
    out.update(pvc_block(w, R.PVC2_BLACKMORE2, 2, k, 911, 60, tag="pb2_"))
    for n_steps in (2, 5):
        out.update(pvc_block(w, R.PVC2_CDFGRID, 2, k, 912 + n_steps, 40, n_steps=n_steps, tag=f"cdf{n_steps}_"))
    # CentralizedUncertainLinearStatePropagator: composite states of k = 3 agents, step size 0.5 (OmplSetUp.cpp:357)
    kc, nc = 3, 400
    rngc = rng_of(31)
    Wc = 2 * kc + 2 * (2 * kc) ** 2
    cs = np.zeros((nc, Wc))
    cs[:, :2 * kc] = rngc.uniform(0, 30, (nc, 2 * kc))
    full = rngc.uniform(-0.01, 0.01, (nc, 2, 2 * kc, 2 * kc))              # arbitrary off-block entries: the propagator must ignore them
    for a_ in range(kc):
        for which in range(2):
            full[:, which, 2 * a_:2 * a_ + 2, 2 * a_:2 * a_ + 2] = spd(rngc, nc, 2, 0.03, 0.2).reshape(nc, 2, 2)
    cs[:, 2 * kc:] = full.reshape(nc, -1)
    cu = rngc.uniform(-1, 1, (nc, 2 * kc))
    out.update(central_k=kc, central_states=cs, central_controls=cu, central_propagate=R.central_propagate(kc, cs, cu, 0.5))
    np.savez_compressed(os.path.join(out_dir, "ref2_config2_random32_10_linear_k4.npz"), **out)
    print("config2: M", w.n_obstacles, "cs valid %.2f" % out["cs_is_valid"].mean(), "cs full %.2f" % (out["cs_rollout_valid"] == steps).mean(),
          "goal ok %.2f" % ok.mean(), "pcs", len(out["pcs_conflicts"]), "pb2", len(out["pb2_conflicts"]), "cdf2", len(out["cdf2_conflicts"]),
          "cdf5", len(out["cdf5_conflicts"]))


if __name__ == "__main__":
    main()
