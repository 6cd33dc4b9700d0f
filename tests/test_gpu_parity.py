"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same
seeded inputs.  Bars: linear model + Blackmore/ChiSquared checkers are BIT-EXACT (values,
verdicts, first-violation step); unicycle values within 1e-9 relative (fp64; device sincos
vs glibc), verdicts exact outside the +-1e-6 band around the chance threshold; adaptive
checkers (device erfinv vs Boost-style erf_inv) exact outside the band.
"""
import numpy as np
import pytest

import helpers
from helpers import assert_rollout_mismatches_in_band, band_mask_halfplane, install_from_oracle_svc, oracle_synth

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-9, 1e-13


def _cov2(aos, dim):
    S = aos[:, dim:dim + dim * dim]
    L = aos[:, dim + dim * dim:]
    return np.stack([S[:, 0] + L[:, 0], S[:, 1] + L[:, 1], S[:, dim] + L[:, dim], S[:, dim + 1] + L[:, dim + 1]], axis=1)


# --------------------------------------------------------------------------- propagate
def test_propagate_linear_bitexact(K, O, W, ctx):
    ctx.set_model(K.MODEL_LINEAR2D)
    bel, ctrl = W.synthetic_beliefs(10000, "linear")
    out = ctx.propagate(bel, ctrl)
    ref = O.propagate(K.soa_to_aos(bel), K.soa_to_aos(ctrl), 2)
    assert np.array_equal(K.soa_to_aos(out), ref)
    # chained: 3 steps from the allocState default reproduce the closed-form known answers
    b = K.aos_to_soa(O.default_beliefs(1, 2))
    u = np.array([[0.5], [-0.25]])
    for _ in range(3):
        b = ctx.propagate(b, u)
    assert b[2, 0] == 0.0061904761904761916 and b[6, 0] == 0.019541513809523813


def test_propagate_unicycle(K, O, W, ctx):
    ctx.set_model(K.MODEL_UNICYCLE)
    bel, ctrl = W.synthetic_beliefs(6000, "unicycle")
    out = K.soa_to_aos(ctx.propagate(bel, ctrl))
    ref = O.propagate(K.soa_to_aos(bel), K.soa_to_aos(ctrl), 4)
    assert np.allclose(out, ref, rtol=RTOL, atol=ATOL)
    # the covariance recursion has no transcendental: bit-exact
    assert np.array_equal(out[:, 4:], ref[:, 4:])
    # `duration` argument is honoured by the unicycle (quirk 1)
    out2 = K.soa_to_aos(ctx.propagate(bel, ctrl, duration=0.35))
    ref2 = O.propagate(K.soa_to_aos(bel), K.soa_to_aos(ctrl), 4, duration=0.35)
    assert np.allclose(out2, ref2, rtol=RTOL, atol=ATOL)


def test_propagate_empty_and_ragged(K, O, W, ctx):
    ctx.set_model(K.MODEL_LINEAR2D)
    out = ctx.propagate(np.zeros((10, 0)), np.zeros((2, 0)))
    assert out.shape == (10, 0)
    for n in (1, 31, 33, 257):
        bel, ctrl = W.synthetic_beliefs(n, "linear", seed=7 + n)
        assert np.array_equal(K.soa_to_aos(ctx.propagate(bel, ctrl)), O.propagate(K.soa_to_aos(bel), K.soa_to_aos(ctrl), 2))


# --------------------------------------------------------------------------- isValid
@pytest.mark.parametrize("model", ["linear", "unicycle"])
@pytest.mark.parametrize("kind", [0, 1, 2])
def test_is_valid(K, O, W, ctx, model, kind):
    dim = 2 if model == "linear" else 4
    world, svc = oracle_synth(O, W, 256, "realistic", kind, dim)
    install_from_oracle_svc(K, ctx, svc, K.MODEL_LINEAR2D if dim == 2 else K.MODEL_UNICYCLE)
    bel, _ = W.synthetic_beliefs(8000, model, seed=11)
    # shrink covariances for half of the batch so that both verdicts are well represented
    bel[dim:, ::2] *= 0.05
    got = ctx.is_valid(bel)
    aos = K.soa_to_aos(bel)
    ref, _ = svc.is_valid(aos)
    mism = np.nonzero(got != ref)[0]
    assert 0.05 < ref.mean() < 0.95
    if kind == 1:   # device erfinv: exact outside the band
        near = band_mask_halfplane(svc, aos[mism], _cov2(aos[mism], dim), 3.0, band=1e-6) if len(mism) else np.array([])
        assert len(mism) <= 2 and near.all()
    else:
        assert len(mism) == 0


def test_is_valid_no_obstacles_is_bounds_only(K, O, W, ctx):
    world = O.World.synthetic(np.zeros((0, 2)), 8.0, 8.0, 2, 0.9)
    svc = O.Svc(world, 0, 2)
    install_from_oracle_svc(K, ctx, svc, K.MODEL_LINEAR2D)
    bel, _ = W.synthetic_beliefs(4000, "linear", seed=3)
    bel[0] = np.linspace(-2, 10, 4000)
    bel[1] = np.linspace(0.0, 7.0, 4000)[::-1]
    bel[1, 2000:2010] = np.nan        # NaN passes satisfiesBounds (OMPL semantics)
    bel[0, 0] = -1.0 - 1e-16          # inside by the DBL_EPSILON slack
    got = ctx.is_valid(bel)
    ref, _ = svc.is_valid(K.soa_to_aos(bel))
    assert np.array_equal(got, ref) and ref[2000:2010].all() and 0.5 < ref.mean() < 0.9


# --------------------------------------------------------------------------- propagateWhileValid
@pytest.mark.parametrize("variant", ["realistic", "all_survive"])
def test_rollout_linear_blackmore_bitexact(K, O, W, ctx, variant):
    world, svc = oracle_synth(O, W, 256, variant, 0, 2)
    install_from_oracle_svc(K, ctx, svc, K.MODEL_LINEAR2D)
    n, T = 20000, 50
    bel, ctrl = W.synthetic_beliefs(n, "linear", variant=variant)
    steps = np.full(n, T, dtype=np.int32)
    out, valid = ctx.propagate_while_valid(bel, ctrl, steps)
    ref_out, ref_valid, cnt = svc.rollout(K.soa_to_aos(bel), K.soa_to_aos(ctrl), steps, nthreads=8)
    assert np.array_equal(valid, ref_valid)
    assert np.array_equal(K.soa_to_aos(out), ref_out)
    if variant == "all_survive":
        assert (valid == T).all() and int(cnt[0]) == n * T
    else:
        assert 0 < (valid == T).mean() < 1 and (valid == 0).any()


def test_rollout_linear_ragged_steps_and_traj(K, O, W, ctx):
    world, svc = oracle_synth(O, W, 102, "realistic", 0, 2, seed=5)
    install_from_oracle_svc(K, ctx, svc, K.MODEL_LINEAR2D)
    n = 5003
    bel, ctrl = W.synthetic_beliefs(n, "linear", seed=9)
    bel[2:, :] *= 0.02
    rng = np.random.default_rng(1)
    steps = rng.integers(0, 11, size=n).astype(np.int32)     # includes steps == 0 (copy, return 0)
    out, valid, traj = ctx.propagate_while_valid(bel, ctrl, steps, traj_T=10)
    ref_out, ref_valid, cnt, ref_traj = svc.rollout(K.soa_to_aos(bel), K.soa_to_aos(ctrl), steps, traj_stride=10)
    assert np.array_equal(valid, ref_valid)
    assert np.array_equal(K.soa_to_aos(out), ref_out)
    assert (valid[steps == 0] == 0).all()
    for t in range(10):
        m = valid > t
        assert np.array_equal(traj[t][:, m].T, ref_traj[m, t, :])


@pytest.mark.parametrize("kind", [1, 2])
def test_rollout_linear_other_checkers(K, O, W, ctx, kind):
    world, svc = oracle_synth(O, W, 64, "realistic", kind, 2, seed=21)
    install_from_oracle_svc(K, ctx, svc, K.MODEL_LINEAR2D)
    n = 6000
    bel, ctrl = W.synthetic_beliefs(n, "linear", seed=22)
    bel[2:, :] *= 0.05
    steps = np.full(n, 12, dtype=np.int32)
    out, valid = ctx.propagate_while_valid(bel, ctrl, steps)
    ref_out, ref_valid, _ = svc.rollout(K.soa_to_aos(bel), K.soa_to_aos(ctrl), steps, nthreads=8)
    if kind == 2:
        assert np.array_equal(valid, ref_valid) and np.array_equal(K.soa_to_aos(out), ref_out)
    else:
        # adaptive: device erfinv differs from the oracle's erf_inv in the last ulps - a verdict may differ only where the
        # deciding state is inside the +-1e-6 band, which is asserted state by state
        ok = assert_rollout_mismatches_in_band(O, svc, K.soa_to_aos(bel), K.soa_to_aos(ctrl), valid, ref_valid)
        assert np.array_equal(K.soa_to_aos(out)[ok], ref_out[ok])


@pytest.mark.parametrize("kind", [0, 1, 2])
def test_rollout_unicycle(K, O, W, ctx, kind):
    world, svc = oracle_synth(O, W, 40, "realistic", kind, 4, seed=31)
    install_from_oracle_svc(K, ctx, svc, K.MODEL_UNICYCLE)
    n = 4000
    bel, ctrl = W.synthetic_beliefs(n, "unicycle", seed=32)
    bel[4:, :] *= 0.02
    rng = np.random.default_rng(2)
    steps = rng.integers(1, 11, size=n).astype(np.int32)
    out, valid = ctx.propagate_while_valid(bel, ctrl, steps)
    ref_out, ref_valid, _ = svc.rollout(K.soa_to_aos(bel), K.soa_to_aos(ctrl), steps, nthreads=8)
    # device sincos differs from libm in the last ulps: verdicts exact outside the +-1e-6 band, asserted state by state
    ok = assert_rollout_mismatches_in_band(O, svc, K.soa_to_aos(bel), K.soa_to_aos(ctrl), valid, ref_valid)
    assert np.allclose(K.soa_to_aos(out)[ok], ref_out[ok], rtol=RTOL, atol=1e-12)
    assert 0.02 < (ref_valid == steps).mean() < 0.98


def test_rollout_full_size_properties(K, W, ctx):
    """BASELINE config-5 size (2^20 x 50 x 256) through size-independent properties:
    all-survive => every belief reaches T and the mean moved exactly T*dt*u (linearity);
    rollout(T) == rollout(T/2) o rollout(T/2) (composition); realistic final == traj[valid-1]."""
    n, T = 1 << 20, 50
    W.install_synthetic(ctx, "linear", 256, variant="all_survive")
    bel, ctrl = W.synthetic_beliefs(n, "linear", variant="all_survive")
    steps = np.full(n, T, dtype=np.int32)
    out, valid = ctx.propagate_while_valid(bel, ctrl, steps)
    assert (valid == T).all()
    assert np.allclose(out[0] - bel[0], T * 0.2 * ctrl[0], rtol=0, atol=1e-9)
    half = np.full(n, T // 2, dtype=np.int32)
    mid, v1 = ctx.propagate_while_valid(bel, ctrl, half)
    out2, v2 = ctx.propagate_while_valid(mid, ctrl, half)
    assert np.array_equal(out2, out) and (v1 + v2 == T).all()
    # covariance recursion is control independent: permuting controls leaves Sigma/Lambda unchanged
    out3, _ = ctx.propagate_while_valid(bel, ctrl[:, ::-1].copy(), steps)
    assert np.array_equal(out3[2:], out[2:])


# --------------------------------------------------------------------------- launch accounting / errors
def test_launch_count_and_errors(K, W, ctx):
    with pytest.raises(K.CckError):
        ctx.propagate(np.zeros((10, 4)), np.zeros((2, 4)))       # no model installed
    ctx.set_model(K.MODEL_LINEAR2D)
    with pytest.raises(K.CckError):
        ctx.is_valid(np.zeros((10, 4)))                           # no checker installed
    before = ctx.launch_count
    bel, ctrl = W.synthetic_beliefs(100, "linear")
    ctx.propagate(bel, ctrl)
    assert ctx.launch_count == before + 1
    assert ctx.num_sms >= 100 or helpers.HOSTSIM


def test_summary_and_winner_records(K, W, ctx):
    """The per-rank result records that are all-gathered over NCCL (k_summary, k_winner)."""
    import torch
    W.install_synthetic(ctx, "linear", 256, variant="realistic")
    n, T = 50000, 20
    bel, ctrl = W.synthetic_beliefs(n, "linear", seed=4)
    steps = np.random.default_rng(0).integers(0, T + 1, n).astype(np.int32)
    out, valid = ctx.propagate_while_valid(bel, ctrl, steps)
    dev = helpers.torch_device()
    d_steps, d_valid = torch.from_numpy(steps).to(dev), torch.from_numpy(valid).to(dev)
    d_sum = torch.zeros(3, dtype=torch.int64, device=dev)
    ctx.rollout_summary_dev(None, n, d_steps, d_valid, d_sum)
    ctx.synchronize()
    helpers.torch_sync()
    ref = [int((valid == steps).sum()), int(np.where(steps > 0, np.minimum(valid + 1, steps), 0).sum()), int(valid.sum())]
    assert d_sum.tolist() == ref
    rng = np.random.default_rng(1)
    cost = rng.uniform(0, 100, n)
    flags = rng.integers(0, 8, n).astype(np.uint8)
    adm = np.nonzero((flags & 3) == 3)[0]
    cost[adm[[10, 500]]] = -1.0                                  # tie: lowest index wins
    d_cost, d_flags = torch.from_numpy(cost).to(dev), torch.from_numpy(flags).to(dev)
    rec = torch.zeros(2, dtype=torch.float64, device=dev)
    ctx.select_winner_dev(None, n, d_cost, d_flags, 3, 1000, rec)
    ctx.synchronize()
    helpers.torch_sync()
    assert rec.tolist() == [-1.0, float(1000 + adm[10])]
    ctx.select_winner_dev(None, n, d_cost, torch.zeros_like(d_flags), 3, 0, rec)
    ctx.synchronize()
    helpers.torch_sync()
    assert rec.tolist() == [float("inf"), -1.0]


@pytest.mark.parametrize("model", ["linear", "unicycle"])
def test_leading_dimensions_are_honoured(K, O, W, ctx, model):
    """Every batch argument of the C ABI carries its own leading dimension (plane f of item i at base[f*ld + i]); the planner
    passes ld = capacity > n.  Strided calls must give the contiguous results bit for bit and must not touch the padding."""
    P = K.capi._ptr
    L = K.capi.lib()
    dim = 2 if model == "linear" else 4
    Wd = K.width(dim)
    world, svc = oracle_synth(O, W, 102, "realistic", 0, dim, seed=5)
    install_from_oracle_svc(K, ctx, svc, K.MODEL_LINEAR2D if dim == 2 else K.MODEL_UNICYCLE)
    n, ld, ldc, ldo, ldt, T = 777, 1000, 901, 1203, 811, 12
    bel, ctrl = W.synthetic_beliefs(n, model, seed=3)
    bel[dim:] *= 0.05
    steps = np.random.default_rng(1).integers(0, T + 1, n).astype(np.int32)
    ref_p = ctx.propagate(bel, ctrl)
    ref_v = ctx.is_valid(bel)
    ref_out, ref_valid, ref_traj = ctx.propagate_while_valid(bel, ctrl, steps, traj_T=T)
    sentinel = -7.25
    b2 = np.full((Wd, ld), sentinel); b2[:, :n] = bel
    c2 = np.full((dim, ldc), sentinel); c2[:, :n] = ctrl
    o2 = np.full((Wd, ldo), sentinel)
    assert L.cck_propagate(ctx.h, n, P(b2), ld, P(c2), ldc, 0.2, P(o2), ldo) == 0
    assert np.array_equal(o2[:, :n], ref_p, equal_nan=True) and (o2[:, n:] == sentinel).all()
    v2 = np.full(n + 5, 9, dtype=np.uint8)
    assert L.cck_is_valid(ctx.h, n, P(b2), ld, P(v2)) == 0
    assert np.array_equal(v2[:n].astype(bool), ref_v) and (v2[n:] == 9).all()
    o3 = np.full((Wd, ldo), sentinel)
    val = np.full(n + 3, -5, dtype=np.int32)
    tr = np.full((T, Wd, ldt), sentinel)
    assert L.cck_propagate_while_valid(ctx.h, n, P(b2), ld, P(c2), ldc, P(steps), P(o3), ldo, P(val), P(tr), ldt, T) == 0
    assert np.array_equal(val[:n], ref_valid) and (val[n:] == -5).all()
    assert np.array_equal(o3[:, :n], ref_out, equal_nan=True) and (o3[:, n:] == sentinel).all()
    mask = (np.arange(T)[:, None] < ref_valid[None, :])                   # [T, n]: the valid steps of every belief
    assert np.array_equal(tr[:, :, :n].transpose(0, 2, 1)[mask], ref_traj.transpose(0, 2, 1)[mask], equal_nan=True)
    assert (tr[:, :, n:] == sentinel).all()
    # batched expansion with strides
    rad = K.rect_robot_bounding_radius()
    K.install_pvc(ctx, K.PVC_BLACKMORE, dim, [rad, rad], world.p_safe_agents)
    other = O.default_beliefs(1, dim)[0]; other[:2] = [30.0, 30.0]
    ctx.set_constraints(*K.constraint_entries(2, dim, 0, [(1, list(range(3, 20)), np.repeat(other[None, :], 17, axis=0))]))
    pstep = np.random.default_rng(2).integers(0, 10, n).astype(np.int32)
    pcost = np.random.default_rng(3).uniform(0, 3, n)
    st1 = np.maximum(steps, 1).astype(np.int32)
    ref = ctx.expand_batch(bel, ctrl, st1, pstep, pcost, (30.0, 30.0))
    g = K.capi.GoalParams(30.0, 30.0, 2.0, K.chi2_quantile_2dof(0.95))
    o4 = np.full((Wd, ldo), sentinel)
    val4, cost4, gd4, fl4 = np.zeros(n, dtype=np.int32), np.zeros(n), np.zeros(n), np.zeros(n, dtype=np.uint8)
    import ctypes
    assert L.cck_expand_batch(ctx.h, n, P(b2), ld, P(c2), ldc, P(st1), P(pstep), P(pcost), ctypes.byref(g), P(o4), ldo, P(val4), P(cost4), P(gd4), P(fl4)) == 0
    assert np.array_equal(o4[:, :n], ref[0], equal_nan=True) and (o4[:, n:] == sentinel).all()
    assert np.array_equal(val4, ref[1]) and np.array_equal(cost4, ref[2]) and np.array_equal(gd4, ref[3]) and np.array_equal(fl4, ref[4])
    ctx.set_constraints([], [], np.zeros((0, 2)), np.zeros((0, 4)))
