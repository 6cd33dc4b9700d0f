"""CPU tests of the product library's host side: it loads, exports every symbol the header
declares, its constant builders match the oracle bit-for-bit, and compute entry points fail
loudly without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(K):
    L = K.capi.lib()
    hdr = open(os.path.join(ROOT, "include", "cck_b200.h")).read()
    declared = set(re.findall(r"\b(cck_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(K.SYMBOLS), declared ^ set(K.SYMBOLS)
    for s in declared:
        assert hasattr(L, s), s


def test_struct_sizes_match_header(K):
    assert ctypes.sizeof(K.capi.ModelParams) == 8 + 8 + 3 * 32 + 4 * 128 + 64 + 32
    assert ctypes.sizeof(K.capi.SvcParams) == 8 + 64 + 32
    assert ctypes.sizeof(K.capi.ConflictRun) == 16 and ctypes.sizeof(K.capi.GoalParams) == 32


def test_host_builders_match_oracle_bitexact(K, O, W):
    zs = np.concatenate([np.linspace(-0.99, 0.99, 50), 1 - np.logspace(-14, -1, 40)])
    assert all(K.erf_inv(z) == O.erf_inv(z) for z in zs)
    assert K.chi2_quantile_2dof(0.9) == O.chi2_quantile_2dof(0.9)
    centres = W.synthetic_obstacles(256)
    world = O.World.synthetic(centres, 64.0, 64.0, 4, 0.9)
    svc = O.Svc(world)
    rad = K.rect_robot_bounding_radius()
    assert rad == world.bounding_rad[0]
    A, B, R = K.build_obstacle_tables(centres, rad)
    assert np.array_equal(A, svc.A) and np.array_equal(B, svc.B) and np.array_equal(R, svc.rects)
    assert K.split_risk(0.9, 256, 4) == (world.p_safe_obs, world.p_safe_agents)
    assert K.erf_inv(1 - 2 * ((1 - world.p_safe_obs) / 256)) == svc.erf_inv_result
    p = O.Pvc(world)
    for bb, kind in ((False, O.PVC_BLACKMORE), (True, O.PVC_BOUNDINGBOX)):
        A2, B2 = K.build_pair_halfplanes(rad, rad, bb)
        Ao, Bo = O.Pvc(world, kind).pair_halfplanes(0, 1)
        assert np.array_equal(A2, Ao) and np.array_equal(B2, Bo)
    assert K.erf_inv(1 - 2 * p.p_coll_dist) == p.c_pair


def test_model_params_match_oracle(K, O):
    lp, op = K.model_params(K.MODEL_LINEAR2D, 0.2), O.linear_params(0.2)
    assert list(lp.lin_Q) == list(op["Q"]) and list(lp.lin_R) == list(op["R"]) and list(lp.lin_Acl) == list(op["Acl"])
    up, ou = K.model_params(K.MODEL_UNICYCLE, 0.2), O.unicycle_params(0.2)
    for name, key in (("uni_F", "F"), ("uni_Acl", "Acl"), ("uni_Q", "Q"), ("uni_R", "R"), ("uni_gains", "gains")):
        assert list(getattr(up, name)) == list(ou[key]), name
    assert list(up.uni_acc) == list(ou["acc"]) and list(up.uni_turn) == list(ou["turn"])


def test_distinct_normal_classes_are_few(K, W):
    """The kernel hoists sqrt(a^T P a) per class of bit-identical normals: the synthetic field
    and the reference maps yield only a handful of classes."""
    centres = W.synthetic_obstacles(256)
    A, _, _ = K.build_obstacle_tables(centres, K.rect_robot_bounding_radius())
    classes = {a.tobytes() for a in A}
    assert 1 <= len(classes) <= 64


def test_no_cpu_fallback_without_gpu(K):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(K.CckError) as e:
        K.Context(0)
    assert e.value.code == K.capi.ERR_NODEVICE


def test_kernel_pair_tests_on_the_host_match_the_oracle(K, O):
    """The pair tests (*PVC::isSafe_) the validatePlan / satisfiesConstraints kernels call are plain functions of
    cck_kernels.cuh; cck_host_pair_safe runs that very source compiled for the host.  Against the oracle, pair by pair, for
    the five kinds without a data-dependent risk allocation - bit-exact verdicts (same libm on both sides), including the
    Blackmore2 and CDF-grid checkers and non-symmetric / degenerate covariances."""
    rng = np.random.default_rng(33)
    n = 4000
    world = O.World.synthetic(np.zeros((0, 2)), 20.0, 20.0, 4, 0.9)
    rad = K.rect_robot_bounding_radius()
    mu_a, mu_b = rng.uniform(0, 3, (n, 2)), rng.uniform(0, 3, (n, 2))
    mu_b[::5] = mu_a[::5] + rng.uniform(-0.6, 0.6, (len(mu_a[::5]), 2))        # close pairs
    L = rng.uniform(0.01, 0.3, (2, n, 3))
    cov = [np.stack([l[:, 0] ** 2, l[:, 0] * l[:, 1], l[:, 0] * l[:, 1], l[:, 1] ** 2 + l[:, 2] ** 2], axis=1) for l in L]
    cov[0][::7, 1] += 1e-3                                                      # non-symmetric (what the unicycle recursion produces)
    cov[1][::11] = [0.02, 0.0, 0.0, 0.02]                                       # multiples of I: the linear model's covariances
    cov[0][::11] = [0.02, 0.0, 0.0, 0.02]
    cov[1][3::13] = [0.03, 0.0, 0.0, 0.01]                                      # diagonal, unequal
    beliefs = [np.concatenate([m, c, np.zeros((n, 4))], axis=1) for m, c in ((mu_a, cov[0]), (mu_b, cov[1]))]   # Sigma = cov, Lambda = 0
    cases = [(K.PVC_BLACKMORE, O.PVC_BLACKMORE, 2), (K.PVC_BOUNDINGBOX, O.PVC_BOUNDINGBOX, 2), (K.PVC_CHISQUARED, O.PVC_CHISQUARED, 2),
             (K.PVC_BLACKMORE2, O.PVC_BLACKMORE2, 2), (K.PVC_CDFGRID, O.PVC_CDFGRID, 2), (K.PVC_CDFGRID, O.PVC_CDFGRID, 3), (K.PVC_CDFGRID, O.PVC_CDFGRID, 6)]
    for kk, ok_, gn in cases:
        pv = O.Pvc(world, ok_, grid_n=gn)
        A, B = K.build_pair_halfplanes(rad, rad, kk == K.PVC_BOUNDINGBOX)
        oA, oB = pv.pair_halfplanes(0, 1)
        if kk not in (K.PVC_CHISQUARED, K.PVC_CDFGRID):
            assert np.array_equal(A, oA) and np.array_equal(B, oB)
        unsafe = 0
        for i in range(n):
            got = K.host_pair_safe(kk, mu_a[i], cov[0][i], mu_b[i], cov[1][i], A, B, rad, rad, c_pair=pv.c_pair, p_coll_agnts=pv.p_coll_agnts, sc=pv.sc, grid_n=gn)
            # the oracle's verdict for the same pair: a two-robot, one-step plan conflicts iff the pair is unsafe
            conf = pv.validate_plan([beliefs[0][i:i + 1], beliefs[1][i:i + 1]] + [np.array([[100.0 + 10 * r, 100.0, 1e-4, 0, 0, 1e-4, 0, 0, 0, 0]]) for r in range(2)], 2)
            want = 0 if (conf and conf[0][:2] == (0, 1) and conf[0][2] == 0) else 1
            assert got == want, (kk, gn, i)
            unsafe += 1 - want
        assert 0.02 * n < unsafe < 0.9 * n, (kk, gn, unsafe)


def test_kernel_decagon_pair_test_on_the_host_matches_the_oracle(K, O):
    """k_chisq_pairs' source compiled for the host against the oracle, incl. vertex-vertex contacts a few ulps either way."""
    rng = np.random.default_rng(21)
    n = 6000
    mu_a = rng.uniform(0, 8, (n, 2))
    L = rng.uniform(0.02, 0.3, (2, n, 3))
    cov = [np.ascontiguousarray(np.stack([l[:, 0] ** 2, l[:, 0] * l[:, 1], l[:, 0] * l[:, 1], l[:, 1] ** 2 + l[:, 2] ** 2], axis=1)) for l in L]
    rad = rng.uniform(0.1, 0.3, (2, n))
    sc = K.chi2_quantile_2dof(0.9)
    lam = [0.5 * (c[:, 0] + c[:, 3]) + np.sqrt((0.5 * (c[:, 0] + c[:, 3])) ** 2 - (c[:, 0] * c[:, 3] - c[:, 1] * c[:, 2])) for c in cov]
    ra, rb = lam[0] * sc + rad[0], lam[1] * sc + rad[1]
    ang = rng.uniform(0, 2 * np.pi, n)
    d = (ra + rb) * rng.uniform(0.85, 1.1, n)
    k = np.arange(0, n, 3)
    ang[k] = rng.integers(0, 10, len(k)) * (2 * np.pi / 10)
    d[k] = (ra + rb)[k] * (1 + rng.uniform(-3e-15, 3e-15, len(k)))
    mu_b = np.ascontiguousarray(mu_a + np.stack([d * np.cos(ang), d * np.sin(ang)], axis=1))
    # degenerate covariances (NaN / inf / hugely negative entries): non-finite decagons; the oracle's separating-axis test never
    # takes a NaN projection, the kernel's must not either
    bad = np.arange(5, n, 7)
    for j, i in enumerate(bad):
        cov[j % 2][i, j % 4] = [np.nan, np.inf, -np.inf, -1e300, 1e300][j % 5]
    with np.errstate(all="ignore"):
        ref = O.chisq_pairs_disjoint(mu_a, cov[0], rad[0], mu_b, cov[1], rad[1], sc)
    L_ = K.lib()
    got = np.array([L_.cck_host_chisq_pair_disjoint(mu_a[i].ctypes.data, cov[0][i].ctypes.data, float(rad[0][i]), mu_b[i].ctypes.data, cov[1][i].ctypes.data,
                                                     float(rad[1][i]), float(sc)) for i in range(n)], dtype=bool)
    assert np.array_equal(got, ref)
    assert 0.2 < ref.mean() < 0.8


def test_build_entry_points_run_without_a_gpu(K):
    """__graft_entry__.build() is the driver's "does it build" check: every source it names must exist and the in-tree
    library must be up to date with them (a stale source list once made it raise FileNotFoundError)."""
    so = K.capi.build()
    assert os.path.exists(so)
    srcs = [f for f in os.listdir(os.path.join(os.path.dirname(so), "csrc")) if f.endswith((".cu", ".cuh"))]
    assert "cck_b200.cu" in srcs and "cck_rollout_grid.cuh" in srcs
    assert all(os.path.getmtime(so) >= os.path.getmtime(os.path.join(os.path.dirname(so), "csrc", f)) for f in srcs)
