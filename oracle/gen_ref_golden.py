#!/usr/bin/env python
"""Generate tests/golden/ref_*.npz from oracle/_ref (the REFERENCE'S OWN SOURCES for the headline path compiled
against oracle/shim/, see oracle/ref_capi.cpp): parser outputs, Minkowski half-plane tables, the erf_inv
constant, propagate / isValid / propagateWhileValid results on seeded beliefs, for the reference's maps and
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


def main():
    assert R.build(force=True), "needs /root/reference"
    out_dir = os.path.join(ROOT, "tests", "golden")
    for name, mp, sc, k in CASES:
        w = R.World(f"{R.REFERENCE}/maps/{mp}.map", f"{R.REFERENCE}/scens/{sc}.scen", k, 0.9)
        s = R.Svc(w)
        b, u, steps = beliefs(3000, w.x_max, w.y_max, 20240501 + k)
        prop = s.propagate(b, u)
        valid = s.is_valid(b)
        fin, nval = s.rollout(b, u, steps)
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), map=mp, scen=sc, k=k, p_safe=0.9,
                            n_obstacles=w.n_obstacles, x_max=w.x_max, y_max=w.y_max, centres=w.centres, starts=w.starts, goals=w.goals,
                            bounding_rad=w.bounding_rad, p_safe_obs=w.p_safe_obs, p_safe_agents=w.p_safe_agents,
                            A=s.A, B=s.B, erf_inv_result=s.erf_inv_result, p_collision=s.p_collision,
                            beliefs=b, controls=u, steps=steps, propagate=prop, is_valid=valid, rollout_final=fin, rollout_valid=nval)
        print(name, "M", w.n_obstacles, "valid %.2f" % valid.mean(), "full %.2f" % (nval == steps).mean())


if __name__ == "__main__":
    main()
