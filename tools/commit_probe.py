#!/usr/bin/env python
"""GPU box: time the planner's commit-side calls (nearest witness, distance columns dense / sparse, tree append) at
planner-like sizes, to see where a launch's commit() spends its ~1 ms."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import importlib
K = importlib.import_module("chance-constrained-k-cbs_b200")
ctx = K.Context(0)
rng = np.random.default_rng(1)
for dim in (2, 4):
    W = dim + 2 * dim * dim
    for n, nf in ((700, 80), (700, 300), (700, 700)):
        b = np.zeros((n, W))
        b[:, :dim] = rng.uniform(0, 32, (n, dim))
        for blk in (0, 1):
            L = np.tril(rng.uniform(0.05, 0.2, (n, dim, dim)))
            b[:, dim + blk * dim * dim: dim + (blk + 1) * dim * dim] = np.einsum("nij,nkj->nik", L, L).reshape(n, dim * dim)
        soa = K.aos_to_soa(b)
        cols = np.sort(rng.choice(n, size=nf, replace=False)).astype(np.int32)
        lim = np.full(n, 0.5)
        def t(f, reps=20):
            f(); f()
            t0 = time.perf_counter()
            for _ in range(reps): f()
            return (time.perf_counter() - t0) / reps * 1e6
        d_us = t(lambda: K.belief_distance_cols(ctx, soa, dim, cols))
        s_us = t(lambda: K.capi.belief_distance_cols_near(ctx, soa, dim, cols, lim))
        l_us = t(lambda: K.belief_distance(ctx, soa, soa, dim))
        print(f"dim {dim} n {n} nf {nf}: cols dense {d_us:.0f} us, cols near {s_us:.0f} us, plain distance(n) {l_us:.0f} us", flush=True)
