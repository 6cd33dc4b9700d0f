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
