#!/usr/bin/env python
"""K-CBS with the per-robot low-level planners - and therefore the two concurrent child replans of every constraint-tree
expansion (KCBS.cpp:417-467) - spread over 1, 2, 4, 8 GPUs of one box (`cck_plan --devices N`: robot r plans on GPU r mod N,
each with its own context, stream, device trees and tables; nothing but the finished trajectories crosses between them, on the
host).  BASELINE configs 2 (k = 4, 6) and 4.  Run on a multi-GPU box:

    python tools/multi_gpu_plan.py --out gpurun_out/multi_gpu_plan.json [--seeds 10]
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import solve_bench as sb  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=10)
    ap.add_argument("--time", type=float, default=60.0)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    import torch
    ngpu = torch.cuda.device_count()
    subprocess.check_call(["make", "-C", sb.HOST, "-s"])
    res = {"gpus_visible": ngpu, "rows": []}
    cases = [c for c in sb.CONFIGS if ("k=4" in c[0] and "config2" in c[0]) or "k=6" in c[0] or "config4" in c[0]]
    with tempfile.TemporaryDirectory() as d:
        for label, fx, k, _scale, p_safe in cases:
            g = np.load(os.path.join(sb.GOLDEN, fx + ".npz"))
            mp, sc = sb.write_instance(d, g, fx)
            for dev in (1, 2, 4, 8):
                if dev > ngpu:
                    continue
                runs = []
                for seed in range(1, a.seeds + 1):
                    out = subprocess.run([os.path.join(sb.HOST, "cck_plan"), "-m", mp, "-s", sc, "-k", str(k), "-p", str(p_safe), "-t", str(a.time), "--K", "1024",
                                          "--seed", str(seed), "--lltime", "20", "--devices", str(dev)], capture_output=True, text=True, timeout=a.time + 60)
                    try:
                        runs.append(json.loads(out.stdout.strip().splitlines()[-1]))
                    except Exception:  # noqa: BLE001
                        runs.append({"solved": False, "seconds": a.time, "stderr": out.stderr[-200:]})
                ok = [r for r in runs if r.get("solved")]
                t = np.array([r["seconds"] for r in ok]) if ok else np.array([np.nan])
                row = {"config": label, "devices": dev, "seeds": a.seeds, "solved": len(ok), "seconds_median": float(np.median(t)),
                       "seconds_q25": float(np.percentile(t, 25)), "seconds_q75": float(np.percentile(t, 75)),
                       "replans_median": float(np.median([r["replans"] for r in ok])) if ok else None,
                       "cbs_nodes_median": float(np.median([r["cbs_nodes"] for r in ok])) if ok else None}
                res["rows"].append(row)
                print(json.dumps(row), flush=True)
    if a.out:
        json.dump(res, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
