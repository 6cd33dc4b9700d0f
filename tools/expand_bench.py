#!/usr/bin/env python
"""Row a12 (batched low-level expansion) at BASELINE config 2's shapes: K in {256, 4096, 65536} candidates per
launch on random-32-32-10 (M = 102 obstacles, k = 4 agents), durations uniform in [1, 10] steps, one active
BeliefConstraint (17 steps) and the chance-constrained goal test, through the host-pointer C ABI call
(cck_expand_batch: H2D + ONE launch + D2H) and device-resident (cck_expand_batch_dev, CUDA events).
The CPU column is the oracle's propagateWhileValid alone on the same candidates (1 thread, as the reference
expands one candidate at a time) — a lower bound on the reference's per-candidate cost.

  python tools/expand_bench.py [--out gpurun_out/expand.json]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--model", default="linear", choices=["linear", "unicycle"])
    a = ap.parse_args()
    import torch
    import cck_b200 as K
    from cck_b200 import workloads as W
    from oracle import oracle as O

    fx = "config2_random32_10_linear_k4.npz" if a.model == "linear" else "config3_random32_5_unicycle_k4.npz"
    g = np.load(os.path.join(ROOT, "tests", "golden", fx))
    dim, k = int(g["dim"]), int(g["k"])
    world = O.World.synthetic(g["centres"], float(g["x_max"]), float(g["y_max"]), k, float(g["p_safe"]))
    svc = O.Svc(world, O.SVC_BLACKMORE, dim)
    ctx = K.Context(0)
    ctx.set_model(K.MODEL_LINEAR2D if dim == 2 else K.MODEL_UNICYCLE)
    ctx.set_obstacles(K.SVC_BLACKMORE, svc.low, svc.high, svc.A, svc.B, svc.centres, svc.rects, erf_inv_result=svc.erf_inv_result)
    rad = K.rect_robot_bounding_radius()
    K.install_pvc(ctx, K.PVC_BLACKMORE, dim, [rad] * k, world.p_safe_agents)
    other = O.default_beliefs(1, dim)[0]; other[:2] = [16.0, 16.0]
    if dim == 4:
        other[3] = 0.1
    csteps = list(range(8, 25))
    ctx.set_constraints(*K.constraint_entries(k, dim, 0, [(1, csteps, np.repeat(other[None, :], len(csteps), axis=0))]))
    dev = torch.device("cuda", 0)
    res = {"model": a.model, "M": int(len(g["centres"])), "agents": k, "rows": []}
    rng = np.random.default_rng(5)
    # the tree's covariances by depth (what a planner's nodes carry): the device's own single-step propagator from 0.01 I
    cur = K.aos_to_soa(O.default_beliefs(1, dim))
    if dim == 4:
        cur[3] = 0.1
    u0 = np.zeros((dim, 1))
    if dim == 4:
        u0[:, 0] = [5.0, 5.0, 0.0, 0.5]
    covs = [cur[dim:, 0].copy()]
    for _ in range(45):
        cur = ctx.propagate(cur, u0)
        covs.append(cur[dim:, 0].copy())
    covs = np.array(covs)
    for Kc in (256, 4096, 65536):
        bel, ctrl = W.synthetic_beliefs(Kc, a.model, seed=100 + Kc)
        bel[0] = rng.uniform(0, float(g["x_max"]) - 1, Kc); bel[1] = rng.uniform(0, float(g["y_max"]) - 1, Kc)
        bel[dim:] *= 0.05
        if dim == 4:
            ctrl[0] = rng.uniform(-1, float(g["x_max"]), Kc); ctrl[1] = rng.uniform(-1, float(g["y_max"]), Kc)
        steps = rng.integers(1, 11, size=Kc).astype(np.int32)
        pstep = rng.integers(0, 30, size=Kc).astype(np.int32)
        pcost = rng.uniform(0, 5, Kc)
        bel[dim:] = covs[pstep].T                      # parents are tree nodes: covariance = the tree's covariance at their depth
        goal = (float(g["goals"][0, 0]), float(g["goals"][0, 1]))
        # host-pointer call (what the planner issues)
        for _ in range(3):
            out, valid, cost, gd, fl = ctx.expand_batch(bel, ctrl, steps, pstep, pcost, goal)
        reps = 50 if Kc <= 4096 else 20
        t0 = time.perf_counter()
        for _ in range(reps):
            ctx.expand_batch(bel, ctrl, steps, pstep, pcost, goal)
        host_us = (time.perf_counter() - t0) / reps * 1e6
        # device-resident, without and with the tree's covariance table (cck_set_cov_table)
        Wd = K.width(dim)
        d = {n_: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for n_, v in dict(bel=bel, ctrl=ctrl, steps=steps, pstep=pstep, pcost=pcost).items()}
        d_out = torch.empty((Wd, Kc), dtype=torch.float64, device=dev); d_valid = torch.empty(Kc, dtype=torch.int32, device=dev)
        d_cost = torch.empty(Kc, dtype=torch.float64, device=dev); d_gd = torch.empty(Kc, dtype=torch.float64, device=dev)
        d_fl = torch.empty(Kc, dtype=torch.uint8, device=dev)
        st = torch.cuda.Stream(device=dev)
        torch.cuda.set_stream(st)

        def launch():
            ctx.expand_batch_dev(st.cuda_stream, Kc, d["bel"], Kc, d["ctrl"], Kc, d["steps"], d["pstep"], d["pcost"], goal, d_out, Kc, d_valid, d_cost, d_gd, d_fl)
        for _ in range(5):
            launch()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(100):
            launch()
        e1.record()
        torch.cuda.synchronize()
        dev_us = e0.elapsed_time(e1) / 100 * 1e3
        assert np.array_equal(d_valid.cpu().numpy(), valid)
        ref_out = d_out.cpu().numpy().copy()
        ctx.set_cov_table(covs[0], 45)
        for _ in range(5):
            launch()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(100):
            launch()
        e1.record()
        torch.cuda.synchronize()
        tab_us = e0.elapsed_time(e1) / 100 * 1e3
        tab_frac = float(((d_fl.cpu().numpy() & K.FLAG_COV_TABLE) != 0).mean())
        assert np.array_equal(d_out.cpu().numpy().view(np.uint64), ref_out.view(np.uint64)), "covariance table changed a result"
        ctx.set_cov_table(None, 0)
        # CPU: the oracle's propagateWhileValid on the same candidates, one thread
        aos_b, aos_c = K.soa_to_aos(bel), K.soa_to_aos(ctrl)
        n_cpu = min(Kc, 4096)
        t0 = time.perf_counter()
        _, rv, cnt = svc.rollout(aos_b[:n_cpu], aos_c[:n_cpu], steps[:n_cpu], nthreads=1)
        cpu_us_per_cand = (time.perf_counter() - t0) / n_cpu * 1e6
        assert np.array_equal(rv, valid[:n_cpu]) or dim == 4
        units = int(np.minimum(valid + 1, steps).sum())
        row = {"K": Kc, "belief_steps": units, "full_fraction": float(((fl & 1) != 0).mean()),
               "host_call_us": host_us, "host_candidates_per_s": Kc / host_us * 1e6,
               "device_launch_us": dev_us, "device_launch_us_cov_table": tab_us, "cov_table_fraction": tab_frac, "device_candidates_per_s": Kc / dev_us * 1e6, "device_belief_steps_per_s": units / dev_us * 1e6,
               "cpu_oracle_us_per_candidate_1thread": cpu_us_per_cand, "host_call_vs_cpu": cpu_us_per_cand * Kc / host_us}
        res["rows"].append(row)
        print(json.dumps(row), flush=True)
    if a.out:
        json.dump(res, open(a.out, "w"), indent=1)
    ctx.close()


if __name__ == "__main__":
    main()
