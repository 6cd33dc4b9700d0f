#!/bin/bash
# GPU box with >= 2 GPUs: K-CBS with the per-agent planners (and the concurrent child replans) spread over 2 devices
set -e
python - <<'PY'
import os, sys, json, subprocess, tempfile
sys.path.insert(0, 'tools')
import numpy as np
import solve_bench as sb
subprocess.check_call(["make", "-C", sb.HOST, "-s"])
g = np.load(os.path.join(sb.GOLDEN, "config2_random32_10_linear_k4.npz"))
with tempfile.TemporaryDirectory() as d:
    mp, sc = sb.write_instance(d, g, "k4")
    for dev in (1, 2):
        ts = []
        for seed in range(1, 6):
            out = subprocess.run([os.path.join(sb.HOST, "cck_plan"), "-m", mp, "-s", sc, "-k", "4", "-t", "60", "--K", "1024", "--seed", str(seed), "--lltime", "20",
                                  "--devices", str(dev)], capture_output=True, text=True)
            r = json.loads(out.stdout.strip().splitlines()[-1])
            assert r["solved"], out.stdout + out.stderr
            ts.append(r["seconds"])
        print(f"devices={dev} solved 5/5 median_seconds={np.median(ts):.3f}")
PY
