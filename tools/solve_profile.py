#!/usr/bin/env python
"""GPU box: ncu launch list of ONE K-CBS solve (config 2, k = 4, seed 1) through host/cck_plan, summarised per kernel:
which kernels a solve launches, how often, and their share of the GPU time.
  python tools/solve_profile.py [--out gpurun_out/solve_launches.json]"""
import argparse, csv, json, os, subprocess, sys, tempfile
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import solve_bench as sb

ap = argparse.ArgumentParser()
ap.add_argument("--out", default="")
ap.add_argument("--fixture", default="config2_random32_10_linear_k4")
ap.add_argument("-k", type=int, default=4)
a = ap.parse_args()
subprocess.check_call(["make", "-C", sb.HOST, "-s"])
g = np.load(os.path.join(sb.GOLDEN, a.fixture + ".npz"))
with tempfile.TemporaryDirectory() as d:
    mp, sc = sb.write_instance(d, g, "inst")
    log = os.path.join(d, "launches.csv")
    cmd = ["ncu", "--metrics", "gpu__time_duration.sum", "--clock-control", "none", "--csv", "--log-file", log,
           os.path.join(sb.HOST, "cck_plan"), "-m", mp, "-s", sc, "-k", str(a.k), "-t", "120", "--K", "1024", "--seed", "1", "--lltime", "60"]
    out = subprocess.run(cmd, capture_output=True, text=True)
    rows = [r for r in csv.reader(open(log)) if len(r) > 10]
    h = rows[0]
    ik, iv = h.index("Kernel Name"), h.index("Metric Value")
    agg = {}
    for r in rows[1:]:
        name = r[ik].split("(")[0].replace("void ", "").replace("cck::", "")
        e = agg.setdefault(name, [0, 0.0])
        e[0] += 1; e[1] += float(r[iv].replace(",", "")) / 1e3     # ns -> us
    tot = sum(v[1] for v in agg.values())
    res = {"solve": json.loads(out.stdout.strip().splitlines()[-1]) if out.stdout.strip() else None, "gpu_time_us": tot,
           "kernels": {k: {"launches": v[0], "us": round(v[1], 1), "share": round(v[1] / tot, 4)} for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])}}
    for k, v in res["kernels"].items():
        print(f"{v['share']*100:5.1f}%  {v['us']:10.1f} us  {v['launches']:5d} x  {k}")
    print("total GPU time %.1f us (under ncu: serialised, cold caches)" % tot)
    if a.out:
        json.dump(res, open(a.out, "w"), indent=1)
