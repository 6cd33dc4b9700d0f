#!/usr/bin/env python
"""CCK-CBS solve-time distribution (the second half of BASELINE.json's metric) on the map-based
configs, through the C++ host (host/cck_plan: K-CBS over the batched GPU BSST).  Instances are
rebuilt from the committed golden fixtures (obstacle centres, starts, goals of the reference's
maps/ and scens/), so this runs on the GPU box without /root/reference.

  python tools/solve_bench.py [--seeds 20] [--time 60] [--K 1024] [--out gpurun_out/solve.json]
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "chance-constrained-k-cbs_b200", "host")
GOLDEN = os.path.join(ROOT, "tests", "golden")
# (label, fixture, agents, seed-count scale): the conflict-heavy k >= 6 case is sampled with fewer seeds
CONFIGS = [("config1 empty-8-8 linear k=2", "config1_empty8_linear_k2", 2, 1.0),
           ("config2 random-32-32-10 linear k=4", "config2_random32_10_linear_k4", 4, 1.0),
           ("config2 random-32-32-10 linear k=6 (first 6 rows of the k=8 scenario)", "config2_random32_10_linear_k8", 6, 0.25),
           ("config3 random-32-32-5 unicycle k=4", "config3_random32_5_unicycle_k4", 4, 1.0),
           ("config4 narrowPassage linear k=2", "config4_narrow_linear_k2", 2, 1.0)]


def write_instance(d, g, name):
    W, H = int(g["x_max"]), int(g["y_max"])
    grid = [["."] * W for _ in range(H)]
    for x, y in g["centres"]:
        grid[int(y)][int(x)] = "@"
    mp = os.path.join(d, f"{name}.map")
    open(mp, "w").write(f"type octile\nheight {H}\nwidth {W}\nmap\n" + "\n".join("".join(r) for r in grid) + "\n")
    model = "2D-Uncertain-Linear-Model" if int(g["dim"]) == 2 else "Uncertain-Unicycle-Model"
    lines = ["version 1"]
    for s, t in zip(g["starts"], g["goals"]):
        lines.append(f"1\t{name}.map\t{W}\t{H}\t{int(s[0])}\t{int(s[1])}\t{int(t[0])}\t{int(t[1])}\tRectangle\t{model}\t1.0")
    sc = os.path.join(d, f"{name}.scen")
    open(sc, "w").write("\n".join(lines) + "\n")
    return mp, sc


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=20)
    ap.add_argument("--time", type=float, default=60.0)
    ap.add_argument("--K", type=int, default=1024)
    ap.add_argument("--only", default="")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    subprocess.check_call(["make", "-C", HOST, "-s"])
    res = {}
    with tempfile.TemporaryDirectory() as d:
        for label, fx, k, scale in CONFIGS:
            if a.only and a.only not in label:
                continue
            g = np.load(os.path.join(GOLDEN, fx + ".npz"))
            mp, sc = write_instance(d, g, fx)
            runs = []
            nseeds = max(1, int(round(a.seeds * scale)))
            for seed in range(1, nseeds + 1):
                out = None
                try:
                    out = subprocess.run([os.path.join(HOST, "cck_plan"), "-m", mp, "-s", sc, "-k", str(k), "-t", str(a.time), "--K", str(a.K),
                                          "--seed", str(seed), "--lltime", str(min(a.time, 20.0))], capture_output=True, text=True, timeout=a.time + 60)
                    runs.append(json.loads(out.stdout.strip().splitlines()[-1]))
                except Exception as e:  # noqa: BLE001
                    err = (out.stderr if "out" in dir() and out is not None else "")[-300:]
                    runs.append({"solved": False, "seconds": a.time, "error": str(e)[:200], "stderr": err})
                    print(f"  seed {seed}: failed: {str(e)[:80]} | {err.strip()[-200:]}", flush=True)
            ok = [r for r in runs if r.get("solved")]
            t = np.array([r["seconds"] for r in ok]) if ok else np.array([np.nan])
            res[label] = {"seeds": nseeds, "solved": len(ok), "time_limit_s": a.time, "K": a.K,
                          "seconds_median": float(np.median(t)), "seconds_mean": float(np.mean(t)), "seconds_p90": float(np.percentile(t, 90)),
                          "seconds_min": float(np.min(t)), "seconds_max": float(np.max(t)),
                          "launches_median": float(np.median([r["launches"] for r in ok])) if ok else None,
                          "candidates_median": float(np.median([r["candidates"] for r in ok])) if ok else None,
                          "cbs_nodes_median": float(np.median([r["cbs_nodes"] for r in ok])) if ok else None}
            if ok and "phase_seconds" in ok[0]:      # where the wall clock goes (medians; planner phases are summed over robots)
                res[label]["root_seconds_median"] = float(np.median([r["root_seconds"] for r in ok]))
                res[label]["validate_seconds_median"] = float(np.median([r["validate_seconds"] for r in ok]))
                res[label]["phase_seconds_median"] = {k2: float(np.median([r["phase_seconds"][k2] for r in ok])) for k2 in ok[0]["phase_seconds"]}
            print(label, json.dumps(res[label]), flush=True)
    if a.out:
        json.dump(res, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
