"""The CPU arm of the solve-time metric (oracle/cck_oracle_plan.cpp: the reference's ConstraintRespectingBSST / KCBS loops on
the oracle, one candidate per iteration).  Its plans are verified by the oracle's own checkers, and the two reference-faithful
behaviours that change results (no pruning, INVALID_START) are asserted.  CPU only."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

sys.path.insert(0, os.path.join(ROOT, "tools"))
import solve_bench as sb  # noqa: E402

PLAN = os.path.join(ROOT, "oracle", "_build", "cck_oracle_plan")


def _read_plan(path):
    trajs, cur = [], None
    for ln in open(path):
        if ln.startswith("robot"):
            cur = []
            trajs.append(cur)
        elif ln.strip():
            cur.append([float(x) for x in ln.split()])
    return [np.array(t) for t in trajs]


def _run(O, tmp_path, fixture, k, p_safe, seed, extra=()):
    O.build()
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s", "_build/cck_oracle_plan"])
    g = dict(np.load(os.path.join(GOLDEN, fixture + ".npz")))
    mp, sc = sb.write_instance(str(tmp_path), g, fixture)
    out_path = str(tmp_path / "plan.txt")
    out = subprocess.run([PLAN, "-m", mp, "-s", sc, "-k", str(k), "-p", str(p_safe), "-t", "60", "--lltime", "20", "--seed", str(seed),
                          "-o", out_path, *extra], capture_output=True, text=True, timeout=180)
    return g, json.loads(out.stdout.strip().splitlines()[-1]), out_path, (mp, sc)


@pytest.mark.parametrize("fixture,k,p_safe", [("config1_empty8_linear_k2", 2, 0.9), ("config4_narrow_linear_k2", 2, 0.8),
                                              ("config3_random32_5_unicycle_k4", 4, 0.9)])
def test_cpu_planner_plans_are_valid(O, tmp_path, fixture, k, p_safe):
    g, res, path, (mp, sc) = _run(O, tmp_path, fixture, k, p_safe, seed=2)
    assert res["solved"] and res["K"] == 1 and res["candidates"] > 0
    dim = int(g["dim"])
    trajs = _read_plan(path)
    assert len(trajs) == k
    world = O.World.from_files(mp, sc, k, p_safe)
    for r, t in enumerate(trajs):
        svc = O.Svc(world, O.SVC_BLACKMORE, dim, robot=r)
        assert np.array_equal(t[0, :2], g["starts"][r])
        assert svc.is_valid(t)[0].all(), f"robot {r}: invalid state in the CPU planner's plan"
        assert O.goal(t[-1:], dim, float(g["goals"][r, 0]), float(g["goals"][r, 1]))[0][0]
        if dim == 2:
            assert np.abs(np.diff(t[:, :2], axis=0)).max() <= 0.2 + 1e-12      # one dt of a control in [-1, 1]^2
    assert O.Pvc(world, O.PVC_BLACKMORE).validate_plan(trajs, dim) == []


def test_cpu_planner_reports_invalid_start_like_ompl(O, tmp_path):
    """narrowPassage at p_safe = 0.9: Robot 0's start state fails PCCBlackmoreSVC by 1.4e-4 -> INVALID_START, at once."""
    _, res, _, _ = _run(O, tmp_path, "config4_narrow_linear_k2", 2, 0.9, seed=1)
    assert not res["solved"] and res["seconds"] < 1.0 and res["cbs_nodes"] == 0 and res["replans"] == 0


def test_reference_never_prunes(O, tmp_path):
    """As written, the reference never deactivates an old representative (ConstraintRespectingBSST.cpp:388-404: `inactive_` is
    only set inside a loop guarded by it).  Default = as written; --prune = stock SST.  Both must solve."""
    _, a, _, _ = _run(O, tmp_path, "config2_random32_10_linear_k4", 4, 0.9, seed=3)
    _, b, _, _ = _run(O, tmp_path, "config2_random32_10_linear_k4", 4, 0.9, seed=3, extra=("--prune",))
    assert a["solved"] and b["solved"] and a["prune"] is False and b["prune"] is True
