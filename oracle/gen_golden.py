#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ from the CPU oracle.

Run HERE (the container that has /root/reference): the reference's own maps/ and scens/
files are parsed by the oracle's restatement of Instance::load_map_/load_agents_, and the
resulting tables + seeded inputs + oracle outputs are stored as small .npz files.  The GPU
box has no /root/reference; tests there rebuild the oracle world from the stored obstacle
centres (World.synthetic builds the same unit-square obstacles the map parser does) and
compare both the oracle and the CUDA path with the stored outputs.

The reference ships no golden vectors of its own: these fixtures pin the
ORACLE (regression) and the product; they do not pin the oracle to the reference binary.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[0] = ROOT   # replace the script directory (it would shadow the oracle package)
from oracle import oracle as O  # noqa: E402

REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")

CONFIGS = {
    # name: (map, scen, k agents, p_safe, dim, n beliefs)
    "config1_empty8_linear_k2": ("empty-8-8.map", "empty-8-8-Rectangles-2DUncertainLinear.scen", 2, 0.9, 2, 1024),
    "config2_random32_10_linear_k4": ("random-32-32-10.map", "random-32-32-10-Rectangles-2DUncertainLinear.scen", 4, 0.9, 2, 2048),
    "config2_random32_10_linear_k8": ("random-32-32-10.map", "random-32-32-10-Rectangles-2DUncertainLinear.scen", 8, 0.9, 2, 1024),
    "config3_random32_5_unicycle_k4": ("random-32-32-5.map", "random-32-32-5-Rectangles-Uncertain-Unicycle.scen", 4, 0.9, 4, 1024),
    "config4_narrow_linear_k2": ("narrowPassage.map", "narrowPassage-2DUncertainLinear.scen", 2, 0.9, 2, 1024),
}


def tree_like_beliefs(rng, world, dim, n):
    """Beliefs as the low-level planner sees them: means spread over the map (a share right at
    the robots' starts), covariances = the reference start covariance advanced 0..40 steps."""
    W = O.width(dim)
    b = O.default_beliefs(n, dim)
    depth = rng.integers(0, 41, size=n)
    # covariance depends on depth only (control independent): roll one belief forward
    chain = [O.default_beliefs(1, dim)[0]]
    if dim == 4:
        chain[0][:4] = [0, 0, 0, 0.1]
    zero_u = np.zeros(dim) if dim == 2 else np.array([0.0, 0.0, 0.0, 0.1])
    for _ in range(40):
        chain.append(O.propagate(chain[-1][None, :], zero_u[None, :], dim)[0])
    chain = np.array(chain)
    b[:, dim:] = chain[depth][:, dim:]
    b[:, 0] = rng.uniform(-0.5, world.x_max - 0.5, n)
    b[:, 1] = rng.uniform(-0.5, world.y_max - 0.5, n)
    k = world.n_robots
    b[:k, :2] = world.starts
    if dim == 4:
        b[:, 2] = rng.uniform(-np.pi, np.pi, n)
        b[:, 3] = rng.uniform(0.05, 1.5, n)
        b[:k, 2], b[:k, 3] = 0.0, 0.1
        lo, hi = world.bounds(4)
        u = np.stack([rng.uniform(lo[i], hi[i], n) for i in range(4)], axis=1)
    else:
        u = rng.uniform(-1.0, 1.0, (n, 2))
    steps = rng.integers(1, 11, size=n).astype(np.int32)      # si->setMinMaxControlDuration(1, 10)
    assert b.shape[1] == W
    return b, u, steps


def plan_trajectories(rng, world, dim, T=60):
    """k time-aligned trajectories that converge on the map centre (conflict heavy), of
    different lengths (robots that finish early wait at their last state)."""
    k = world.n_robots
    cx, cy = world.x_max / 2, world.y_max / 2
    trajs = []
    for r in range(k):
        ang = 2 * np.pi * r / k
        rad = 0.35 * min(world.x_max, world.y_max)
        b0 = O.default_beliefs(1, dim)[0]
        b0[0], b0[1] = cx + rad * np.cos(ang), cy + rad * np.sin(ang)
        if dim == 2:
            u = np.array([-np.cos(ang), -np.sin(ang)]) * rng.uniform(0.85, 1.0)
        else:
            b0[2], b0[3] = ang + np.pi if ang + np.pi <= np.pi else ang - np.pi, 0.3
            u = np.array([cx, cy, b0[2], 0.8])
        n = max(12, int(rad / 0.2 * 1.15) + 10 - 5 * r)
        trajs.append(np.concatenate([b0[None, :], O.propagate_n(b0, u, n, dim)]))
    return trajs


def main():
    os.makedirs(OUT, exist_ok=True)
    rng = np.random.default_rng(20240501)
    for name, (mp, sc, k, p_safe, dim, n) in CONFIGS.items():
        world = O.World.from_files(f"{REF}/maps/{mp}", f"{REF}/scens/{sc}", k, p_safe)
        out = {"centres": world.centres, "x_max": world.x_max, "y_max": world.y_max, "k": k, "p_safe": p_safe, "dim": dim,
               "starts": world.starts, "goals": world.goals, "p_safe_obs": world.p_safe_obs, "p_safe_agents": world.p_safe_agents}
        b, u, steps = tree_like_beliefs(rng, world, dim, n)
        out.update(beliefs=b, controls=u, steps=steps)
        for kind, tag in ((O.SVC_BLACKMORE, "bm"), (O.SVC_ADAPTIVE, "ad"), (O.SVC_CHISQUARED, "cs")):
            svc = O.Svc(world, kind, dim)
            fin, valid, cnt = svc.rollout(b, u, steps)
            ok, _ = svc.is_valid(b)
            out[f"{tag}_final"], out[f"{tag}_valid"], out[f"{tag}_isvalid"] = fin, valid, ok
            if kind == O.SVC_BLACKMORE:
                out["bm_A"], out["bm_B"], out["bm_c"] = svc.A, svc.B, svc.erf_inv_result
                out["bm_counters"] = cnt
        out["prop1"] = O.propagate(b, u, dim)
        trajs = plan_trajectories(rng, world, dim)
        out["traj_lens"] = np.array([len(t) for t in trajs], dtype=np.int32)
        out["traj_states"] = np.concatenate(trajs, axis=0)
        for kind, tag in ((O.PVC_BLACKMORE, "pbm"), (O.PVC_BOUNDINGBOX, "pbb"), (O.PVC_CHISQUARED, "pcs"), (O.PVC_ADAPTIVE_BLACKMORE, "pad")):
            conf = O.Pvc(world, kind).validate_plan(trajs, dim)
            out[f"{tag}_conflicts"] = np.array(conf, dtype=np.int32).reshape(-1, 3)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
        print(name, "M =", world.n_obstacles, "n =", n, "valid(full) =", float((out["bm_valid"] == steps).mean()),
              "conflicts:", {t: len(out[f"{t}_conflicts"]) for t in ("pbm", "pbb", "pcs", "pad")})
    # known answers for the record
    kat = {"erf_inv_0p8": O.erf_inv(0.8), "chi2_0p9": O.chi2_quantile_2dof(0.9), "chi2_0p95": O.chi2_quantile_2dof(0.95)}
    np.savez(os.path.join(OUT, "kat_constants.npz"), **kat)


if __name__ == "__main__":
    main()
