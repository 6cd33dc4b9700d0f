#!/usr/bin/env python
"""Generate tests/golden/ref_*.npz from oracle/_ref (the REFERENCE'S OWN SOURCES for the headline path compiled
against oracle/shim/, see oracle/ref_capi.cpp): parser outputs, Minkowski half-plane tables, the erf_inv
constant, propagate / isValid / propagateWhileValid results on seeded beliefs, validatePlan / createConstraint /
satisfiesConstraints of the Minkowski-sum and bounding-box Blackmore plan checkers on random plans, for the reference's maps and
scenario files of BASELINE configs 1, 2 and 4.  Run in the container where /root/reference is mounted:

    python oracle/gen_ref_golden.py

The vectors travel (the reference does not): the CPU tests check the oracle against them everywhere, the GPU
tests check the CUDA path against them on the B200.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref as R  # noqa: E402

CASES = [("ref_config1_empty8_k2", "empty-8-8", "empty-8-8-Rectangles-2DUncertainLinear", 2),
         ("ref_config2_random32_10_k4", "random-32-32-10", "random-32-32-10-Rectangles-2DUncertainLinear", 4),
         ("ref_config2_random32_10_k8", "random-32-32-10", "random-32-32-10-Rectangles-2DUncertainLinear", 8),
         ("ref_config2_random32_10_k12", "random-32-32-10", "random-32-32-10-Rectangles-2DUncertainLinear", 12),
         ("ref_config4_narrow_k2", "narrowPassage", "narrowPassage-2DUncertainLinear", 2)]


def beliefs(n, x_max, y_max, seed):
    rng = np.random.Generator(np.random.Philox(key=seed))
    b = np.zeros((n, 10))
    b[:, 0] = rng.uniform(-1.5, x_max + 0.5, n)
    b[:, 1] = rng.uniform(-1.5, y_max + 0.5, n)
    for off in (2, 6):                                   # Sigma, Lambda = L L^T, L lower triangular
        l00, l10, l11 = (rng.uniform(0.01, 0.15, n) for _ in range(3))
        b[:, off] = l00 * l00; b[:, off + 1] = l00 * l10; b[:, off + 2] = l00 * l10; b[:, off + 3] = l10 * l10 + l11 * l11
    b[::7, 2:] = 0.0
    b[3::7, 2:6] = [0.01, 0, 0, 0.01]; b[3::7, 6:10] = [0.01, 0, 0, 0.01]      # allocState default
    u = rng.uniform(-1.0, 1.0, (n, 2))
    steps = rng.integers(0, 11, n).astype(np.int32)
    return b, u, steps


def random_plans(k, n_plans, seed):
    """Ragged random-walk trajectories in a small arena (conflicts are common): flat arrays + offsets."""
    rng = np.random.Generator(np.random.Philox(key=seed))
    lens, states = [], []
    for _ in range(n_plans):
        for r in range(k):
            T = int(rng.integers(3, 40))
            t = np.zeros((T, 10))
            t[:, :2] = np.cumsum(rng.uniform(-0.2, 0.2, (T, 2)), axis=0) + rng.uniform(0, 2.5, 2)
            t[:, 2] = t[:, 5] = rng.uniform(0.001, 0.03)
            t[:, 6] = t[:, 9] = rng.uniform(0.001, 0.03)
            t[:, 3] = t[:, 4] = rng.uniform(-0.0005, 0.0005)
            lens.append(T); states.append(t)
    return np.array(lens, dtype=np.int32).reshape(n_plans, k), np.concatenate(states, axis=0)


def pvc_vectors(world, k, seed, n_plans=60):
    """validatePlan / createConstraint / satisfiesConstraints of the reference's Minkowski-sum and bounding-box
    Blackmore checkers on random plans."""
    lens, states = random_plans(k, n_plans, seed)
    out = {"plan_lens": lens, "plan_states": states}
    rng = np.random.Generator(np.random.Philox(key=seed + 1))
    n_probe = 4
    probe_len = rng.integers(2, 45, (n_plans, n_probe)).astype(np.int32)
    probes = np.zeros((n_plans, n_probe, 45, 10))
    probes[..., :2] = np.cumsum(rng.uniform(-0.2, 0.2, (n_plans, n_probe, 45, 2)), axis=2) + rng.uniform(0, 2.5, (n_plans, n_probe, 1, 2))
    probes[..., 2] = probes[..., 5] = 0.01; probes[..., 6] = probes[..., 9] = 0.01
    out["probe_len"], out["probe_states"] = probe_len, probes
    for kind, tag in ((0, "mink"), (1, "bbox")):
        pv = R.Pvc(world, kind)
        out[f"{tag}_info"] = np.array([pv.p_coll_agnts, pv.p_coll_dist, pv.c_pair])
        A, B = pv.pair_halfplanes(0, 1)
        out[f"{tag}_hp_A"], out[f"{tag}_hp_B"] = A, B
        confs, sat = [], np.full((n_plans, n_probe), -1, dtype=np.int8)
        off = 0
        for i in range(n_plans):
            trajs = []
            for r in range(k):
                trajs.append(states[off:off + lens[i, r]]); off += lens[i, r]
            c = pv.validate_plan(trajs)
            confs.append(np.array([[i, a, b, s] for a, b, s in c], dtype=np.int32).reshape(-1, 4))
            if c:
                a = c[0][0]
                cons = pv.create_constraint(trajs, a)
                steps = [x[2] for x in c]
                assert cons[0] == a and np.array_equal(cons[2], np.array(steps) * 0.2)
                for j in range(n_probe):
                    sat[i, j] = pv.satisfies_constraints(probes[i, j, :probe_len[i, j]], a, [(cons[1], steps, cons[3])])
        out[f"{tag}_conflicts"] = np.concatenate(confs, axis=0)       # rows (plan, a, b, step)
        out[f"{tag}_satisfies"] = sat                                  # -1: the plan had no conflict
    return out


def main():
    assert R.build(force=True), "needs /root/reference"
    out_dir = os.path.join(ROOT, "tests", "golden")
    for name, mp, sc, k in CASES:
        w = R.World(f"{R.REFERENCE}/maps/{mp}.map", f"{R.REFERENCE}/scens/{sc}.scen", k, 0.9)
        s = R.Svc(w)
        b, u, steps = beliefs(3000, w.x_max, w.y_max, 20240501 + k)
        pv = {}
        prop = s.propagate(b, u)
        valid = s.is_valid(b)
        fin, nval = s.rollout(b, u, steps)
        pv = pvc_vectors(w, k, 777 + k) if k > 1 else {}
        if k == 4 and w.n_obstacles > 0:      # adaptive-risk checkers (their erf_inv at data-dependent arguments is the stand-in's): structure pin
            sa = R.Svc(w, kind=1)
            pv["adaptive_is_valid"] = sa.is_valid(b[:500])
            lens, states = pv["plan_lens"], pv["plan_states"]
            for kind, tag in ((2, "amink"), (3, "abbox")):
                ap = R.Pvc(w, kind)
                confs, off = [], 0
                for i in range(25):
                    trajs = []
                    for r in range(k):
                        trajs.append(states[off:off + lens[i, r]]); off += lens[i, r]
                    confs.append(np.array([[i, a_, b_, s_] for a_, b_, s_ in ap.validate_plan(trajs)], dtype=np.int32).reshape(-1, 4))
                pv[f"{tag}_conflicts"] = np.concatenate(confs, axis=0)
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), map=mp, scen=sc, k=k, p_safe=0.9, **pv,
                            n_obstacles=w.n_obstacles, x_max=w.x_max, y_max=w.y_max, centres=w.centres, starts=w.starts, goals=w.goals,
                            bounding_rad=w.bounding_rad, p_safe_obs=w.p_safe_obs, p_safe_agents=w.p_safe_agents,
                            A=s.A, B=s.B, erf_inv_result=s.erf_inv_result, p_collision=s.p_collision,
                            beliefs=b, controls=u, steps=steps, propagate=prop, is_valid=valid, rollout_final=fin, rollout_valid=nval)
        print(name, "M", w.n_obstacles, "valid %.2f" % valid.mean(), "full %.2f" % (nval == steps).mean())


if __name__ == "__main__":
    main()
