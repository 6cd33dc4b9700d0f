"""Adversarial parity tests for the fp32 broad phase of the linear + PCCBlackmore rollout kernel
(csrc/cck_rollout_grid.cuh).  The broad phase may only PROVE safety; everything it cannot prove is
decided by the reference's fp64 expression, so verdicts must stay bit-identical to the oracle

  * within a few ulps of the decision threshold of individual half-planes,
  * for non-finite / degenerate beliefs (NaN, inf, negative or zero variances, huge magnitudes),
  * for half-plane tables the map parser never produces (rotated or duplicated planes, non-finite
    offsets, far-away coordinates, many obstacles per grid cell, unusable erf_inv constants).
"""
import numpy as np
import pytest

import helpers
from helpers import install_from_oracle_svc, oracle_synth

pytestmark = pytest.mark.gpu


def _rollout_both(K, ctx, svc, bel, ctrl, steps):
    # isValid on the start states goes through the same broad phase (k_is_valid_grid)
    with np.errstate(all="ignore"):
        assert np.array_equal(ctx.is_valid(bel), svc.is_valid(K.soa_to_aos(bel))[0])
    out, valid = ctx.propagate_while_valid(bel, ctrl, steps)
    ref_out, ref_valid, _ = svc.rollout(K.soa_to_aos(bel), K.soa_to_aos(ctrl), steps, nthreads=8)
    return K.soa_to_aos(out), valid, ref_out, ref_valid


def _bb(a0, a1, P, c):
    """b_bar exactly as HyperplaneCCValidityChecker_ evaluates it (numpy float64 scalars: IEEE, no FMA)."""
    t0 = a0 * P[0] + a1 * P[2]
    t1 = a0 * P[1] + a1 * P[3]
    tmp = t0 * a0 + t1 * a1
    return np.sqrt(np.float64(2.0)) * np.sqrt(tmp) * c


def test_blackmore_threshold_ulp_sweep(K, O, W, ctx):
    """Means placed within +-8 ulps of x*w = B + b_bar (and the y analogue) for 64 obstacles x 4 planes."""
    world, svc = oracle_synth(O, W, 256, "realistic", 0, 2)
    install_from_oracle_svc(K, ctx, svc, K.MODEL_LINEAR2D)
    rng = np.random.default_rng(5)
    base, _ = W.synthetic_beliefs(64, "linear", seed=77)
    base[2:] *= 0.3
    # post-step covariance of each base belief (the state the checker sees after the single step)
    post = O.propagate(K.soa_to_aos(base), np.zeros((64, 2)), 2)
    P = post[:, 2:6] + post[:, 6:10]
    c = np.float64(svc.erf_inv_result)
    cols = []
    obs = rng.choice(256, size=64, replace=False)
    for bi, o in enumerate(obs):
        for i in range(4):
            a0, a1 = np.float64(svc.A[o, i, 0]), np.float64(svc.A[o, i, 1])
            rhs = np.float64(svc.B[o, i]) + _bb(a0, a1, P[bi], c)
            cx, cy = svc.centres[o]
            if a1 == 0:      # x plane: fl(x*a0) >= rhs  around x = rhs / a0, y at the obstacle centre
                v0 = rhs / a0
                for k in range(-8, 9):
                    v = v0
                    for _ in range(abs(k)):
                        v = np.nextafter(v, np.inf if k > 0 else -np.inf)
                    b = base[:, bi].copy(); b[0] = v; b[1] = cy
                    cols.append(b)
            else:
                v0 = rhs / a1
                for k in range(-8, 9):
                    v = v0
                    for _ in range(abs(k)):
                        v = np.nextafter(v, np.inf if k > 0 else -np.inf)
                    b = base[:, bi].copy(); b[0] = cx; b[1] = v
                    cols.append(b)
    bel = np.ascontiguousarray(np.stack(cols, axis=1))
    n = bel.shape[1]
    ctrl = np.zeros((2, n))
    steps = np.ones(n, dtype=np.int32)
    out, valid, ref_out, ref_valid = _rollout_both(K, ctx, svc, bel, ctrl, steps)
    assert np.array_equal(valid, ref_valid)
    assert np.array_equal(out, ref_out)
    # the sweep really straddles thresholds: both verdicts occur inside many 17-point groups
    g = ref_valid.reshape(-1, 17)
    assert ((g.min(axis=1) == 0) & (g.max(axis=1) == 1)).sum() >= 100


@pytest.mark.parametrize("variant", ["realistic", "all_survive"])
def test_blackmore_nonfinite_and_degenerate_beliefs(K, O, W, ctx, variant):
    """all_survive: the obstacle field is far away, so every query window misses the grid and the kernel's one-branch
    fast exit decides - it must not call a belief safe whose covariance poisons the reference's expression (0 * inf)."""
    world, svc = oracle_synth(O, W, 256, variant, 0, 2)
    install_from_oracle_svc(K, ctx, svc, K.MODEL_LINEAR2D)
    n = 6000
    bel, ctrl = W.synthetic_beliefs(n, "linear", seed=123, variant=variant)
    rng = np.random.default_rng(9)
    specials = [np.nan, np.inf, -np.inf, 0.0, -0.0, 1e300, -1e300, 1e30, 1e39, 3.5e38, 1e-300, 5e-324, -1e-3, 1e-45, 1e-38]
    for j in range(n):
        kind = j % 6
        if kind == 0:
            continue                                   # untouched
        f = rng.integers(0, 10) if kind < 4 else rng.integers(2, 10)
        bel[f, j] = specials[rng.integers(len(specials))]
        if kind == 5:                                   # a second special entry
            bel[rng.integers(0, 10), j] = specials[rng.integers(len(specials))]
    bel[2:, 100:200] = 0.0                              # zero covariance
    bel[2:, 200:300] *= -1.0                            # negative "covariance"
    ctrl[:, 300:320] = np.nan
    ctrl[:, 320:340] = 1e300
    steps = rng.integers(0, 4, size=n).astype(np.int32)
    with np.errstate(all="ignore"):
        out, valid, ref_out, ref_valid = _rollout_both(K, ctx, svc, bel, ctrl, steps)
    assert np.array_equal(valid, ref_valid)
    assert np.array_equal(out, ref_out, equal_nan=True)
    assert 0.05 < (ref_valid == steps).mean() < 0.95
    if variant == "all_survive":   # the untouched beliefs all survive; a good part of the poisoned ones must not
        assert (ref_valid == steps)[342::6].all() and (ref_valid < steps).mean() > 0.2   # 342 = first untouched index past the special rows


def _random_tables(rng, M, mode, size=1.0):
    """A[M,4,2], B[M,4] of convex quadrilateral obstacles in a 64x64 world."""
    A, B = np.zeros((M, 4, 2)), np.zeros((M, 4))
    for o in range(M):
        cx, cy = rng.uniform(0, 63, 2)
        hw, hh = rng.uniform(0.3, 1.2, 2) * size
        th = 0.0 if mode == "axis" else rng.uniform(0, np.pi)
        if mode == "mixed" and o % 2 == 0:
            th = 0.0
        w = rng.uniform(0.5, 3.0)
        ct, st = (1.0, 0.0) if th == 0.0 else (np.cos(th), np.sin(th))
        normals = [(st, -ct), (ct, st), (-st, ct), (-ct, -st)]     # bottom, right, top, left (outward)
        ext = [hh, hw, hh, hw]
        for i, ((nx, ny), e) in enumerate(zip(normals, ext)):
            A[o, i] = (w * nx, w * ny)
            B[o, i] = w * (nx * cx + ny * cy + e)
    return A, B


@pytest.mark.parametrize("mode", ["axis", "rotated", "mixed"])
def test_blackmore_arbitrary_tables(K, O, W, ctx, mode):
    rng = np.random.default_rng({"axis": 1, "rotated": 2, "mixed": 3}[mode])
    world, svc = oracle_synth(O, W, 8, "realistic", 0, 2)
    M = 200
    A, B = _random_tables(rng, M, mode)
    if mode == "axis":
        # (the generator leaves both +0.0 and -0.0 components in A)  duplicated directions (two +x planes), swapped plane order, non-finite offsets
        A[0] = [[2.0, 0.0], [1.5, 0.0], [0.0, 1.0], [0.0, -1.0]]; B[0] = [2.0 * 30.0, 1.5 * 31.0, 20.0, -18.0]
        A[1] = A[1][[2, 0, 3, 1]]; B[1] = B[1][[2, 0, 3, 1]]
        B[2, 1] = np.inf; B[3, 0] = -np.inf; B[4, 2] = np.nan; B[5, 1:] = np.inf
        A[6, 3] = [0.0, 0.0]                                           # degenerate plane: 0 >= B + 0
        A[7] *= 1e150; B[7] *= 1e150                                   # normals outside the usable range
        B[8:12] = B[12:16]; A[8:12] = A[12:16]                          # several obstacles in one grid cell
    svc.override_tables(A, B, 2.2)
    install_from_oracle_svc(K, ctx, svc, K.MODEL_LINEAR2D)
    n = 20000
    bel, ctrl = W.synthetic_beliefs(n, "linear", seed=31)
    bel[2:, ::3] *= 0.05
    steps = np.full(n, 12, dtype=np.int32)
    with np.errstate(all="ignore"):
        out, valid, ref_out, ref_valid = _rollout_both(K, ctx, svc, bel, ctrl, steps)
    assert np.array_equal(valid, ref_valid)
    assert np.array_equal(out, ref_out, equal_nan=True)
    assert 0.02 < (ref_valid == steps).mean() < 0.98


@pytest.mark.parametrize("c", [0.0, -1.5, np.nan, np.inf, 1e-12, 7.9])
def test_blackmore_unusual_erf_inv_constants(K, O, W, ctx, c):
    world, svc = oracle_synth(O, W, 102, "realistic", 0, 2, seed=5)
    svc.override_tables(svc.A, svc.B, c)
    install_from_oracle_svc(K, ctx, svc, K.MODEL_LINEAR2D)
    n = 8000
    bel, ctrl = W.synthetic_beliefs(n, "linear", seed=41)
    steps = np.full(n, 8, dtype=np.int32)
    with np.errstate(all="ignore"):
        out, valid, ref_out, ref_valid = _rollout_both(K, ctx, svc, bel, ctrl, steps)
    assert np.array_equal(valid, ref_valid)
    assert np.array_equal(out, ref_out, equal_nan=True)


def test_blackmore_far_away_and_large_tables(K, O, W, ctx):
    """Obstacle fields at large coordinates (fp32 ulp ~ 0.06 at 1e6) and a 4000-obstacle table."""
    rng = np.random.default_rng(17)
    world, svc = oracle_synth(O, W, 8, "realistic", 0, 2)
    for off, M in ((1.0e6, 300), (0.0, 4000), (-3.0e4, 64), (0.0, 20000)):     # M = 4000 / 20000: box arrays stay in global memory
        A, B = _random_tables(rng, M, "axis", size=(0.15 if M <= 4000 else 0.05) if M > 1000 else 1.0)
        A[A == 0] = 0.0
        B[:, 1] += A[:, 1, 0] * off; B[:, 3] += A[:, 3, 0] * off        # shift the field by `off` in x
        B[:, 2] += A[:, 2, 1] * off; B[:, 0] += A[:, 0, 1] * off        # and in y
        svc2 = O.Svc(world, 0, 2, low=np.array([-1.0 + off, -1.0 + off]), high=np.array([64.0 + off, 64.0 + off]))
        svc2.override_tables(A, B, 2.0)
        install_from_oracle_svc(K, ctx, svc2, K.MODEL_LINEAR2D)
        n = 12000
        bel, ctrl = W.synthetic_beliefs(n, "linear", seed=51)
        bel[0] += off; bel[1] += off
        bel[2:, ::2] *= 0.05
        steps = np.full(n, 10, dtype=np.int32)
        out, valid, ref_out, ref_valid = _rollout_both(K, ctx, svc2, bel, ctrl, steps)
        assert np.array_equal(valid, ref_valid), (off, M)
        assert np.array_equal(out, ref_out), (off, M)
        if M <= 4000:        # the 20000-obstacle field is too dense for survivors; it exercises the global-memory box path
            assert 0.02 < (ref_valid == steps).mean() < 0.98, (off, M)
        else:
            assert (ref_valid > 0).any() and (ref_valid == 0).any()


@pytest.mark.parametrize("kind", [0, 2])
def test_unicycle_sparse_step_degenerate_inputs(K, O, W, ctx, kind):
    """The second-generation unicycle kernel skips the exact-zero terms of F, A_cl_d, Q, R; beliefs with
    non-finite covariance entries (where 0*inf = NaN matters) must take the dense step.  Zero, negative and
    huge-but-finite covariances stay on the sparse path.  Covariances are bit-exact, means within 1e-9."""
    world, svc = oracle_synth(O, W, 40, "realistic", kind, 4, seed=31)
    install_from_oracle_svc(K, ctx, svc, K.MODEL_UNICYCLE)
    n = 3000
    bel, ctrl = W.synthetic_beliefs(n, "unicycle", seed=61)
    bel[4:, :] *= 0.02
    rng = np.random.default_rng(13)
    # non-finite covariance entries reach the chi-squared checker too: a decagon with NaN vertices is "separated" from every
    # rectangle in the oracle's separating-axis test (running extrema that only move on a TRUE comparison) and the kernel's
    # literal path must say the same (found by tests/hostsim/fuzz_rollout.py --generic)
    specials = [np.nan, np.inf, -np.inf, 0.0, -0.0, 1e200, -1e-3, 1e-300, 1e300, -1e300]
    for j in range(0, n, 3):
        bel[rng.integers(4, 36), j] = specials[rng.integers(len(specials))]
    bel[4:, 1:300:3] = 0.0
    steps = rng.integers(1, 6, size=n).astype(np.int32)
    aos, cta = K.soa_to_aos(bel), K.soa_to_aos(ctrl)
    with np.errstate(all="ignore"):
        out, valid = ctx.propagate_while_valid(bel, ctrl, steps)
        ref_out, ref_valid, _ = svc.rollout(aos, cta, steps, nthreads=8)
        # device sincos differs from libm in the last ulps: a first-violation step may differ only where the deciding state
        # sits inside the +-1e-6 band of its checker
        ok = helpers.assert_rollout_mismatches_in_band(O, svc, aos, cta, valid, ref_valid, max_count=3)
    out = K.soa_to_aos(out)
    assert np.array_equal(out[ok][:, 4:], ref_out[ok][:, 4:], equal_nan=True)          # Sigma, Lambda: no transcendental
    assert np.allclose(out[ok][:, :4], ref_out[ok][:, :4], rtol=1e-9, atol=1e-12, equal_nan=True)
    assert 0.05 < (ref_valid == steps).mean() < 0.95


def ctx_bounds_ok(svc, aos):
    """OMPL satisfiesBounds on the state values (NaN passes: both comparisons are false)."""
    d = svc.dim
    return ~((aos[:, :d] - 1e-15 > svc.high[None, :d]) | (aos[:, :d] + 1e-15 < svc.low[None, :d])).any(axis=1)


def test_chisq_nonfinite_covariance_follows_the_oracle(K, O, W, ctx):
    """ChiSquaredBoundarySVC on beliefs whose (Sigma + Lambda) block is NaN / inf / hugely negative: lambda_max, the radius and
    every decagon vertex are non-finite.  The oracle's separating-axis test never takes a NaN projection, so such a decagon is
    disjoint from every obstacle and the state is valid as long as its mean is inside the bounds; both models."""
    for model, dim in (("linear", 2), ("unicycle", 4)):
        world, svc = oracle_synth(O, W, 102, "realistic", 2, dim, seed=5)
        install_from_oracle_svc(K, ctx, svc, K.MODEL_LINEAR2D if dim == 2 else K.MODEL_UNICYCLE)
        n = 2000
        bel, ctrl = W.synthetic_beliefs(n, model, seed=9)
        bel[dim:] *= 0.01
        rng = np.random.default_rng(3)
        cov_rows = [dim + f for f in (0, 1, dim, dim + 1)]            # the block the checker reads
        for j in range(n):
            if j % 2 == 0:
                bel[cov_rows[int(rng.integers(4))], j] = [np.nan, np.inf, -np.inf, -1e300, 1e300][int(rng.integers(5))]
        aos = K.soa_to_aos(bel)
        with np.errstate(all="ignore"):
            got, ref = ctx.is_valid(bel), svc.is_valid(aos)[0]
        assert np.array_equal(got, ref), (model, np.flatnonzero(got != ref)[:5])
        touched = np.arange(n) % 2 == 0
        assert 0.3 < ref[touched].mean() < 0.999                       # both verdicts occur among the degenerate beliefs
        nan_cov = np.isnan(aos[:, [dim + 0, dim + 1, dim + dim, dim + dim + 1]]).any(axis=1) & ctx_bounds_ok(svc, aos)
        assert nan_cov.sum() > 50 and ref[nan_cov].all()                # NaN block + mean inside the bounds: valid
        steps = np.full(n, 3, dtype=np.int32)
        with np.errstate(all="ignore"):
            out, valid = ctx.propagate_while_valid(bel, ctrl, steps)
            ref_out, ref_valid, _ = svc.rollout(aos, K.soa_to_aos(ctrl), steps, nthreads=8)
            ok = helpers.assert_rollout_mismatches_in_band(O, svc, aos, K.soa_to_aos(ctrl), valid, ref_valid, max_count=0 if dim == 2 else 3)
        assert ok.sum() >= n - 3


def test_unicycle_far_field_nonfinite(K, O, W, ctx):
    """Unicycle + PCCBlackmore with the obstacle field far away: every window misses the grid, so the one-branch fast
    exit of grid_check decides.  NaN means pass the reference's bounds test and then poison every plane (0 * NaN);
    non-finite covariance entries do the same - none of them may be called safe."""
    world, svc = oracle_synth(O, W, 64, "all_survive", 0, 4, seed=31)
    install_from_oracle_svc(K, ctx, svc, K.MODEL_UNICYCLE)
    n = 3000
    bel, ctrl = W.synthetic_beliefs(n, "unicycle", seed=67, variant="all_survive")
    bel[4:, :] *= 0.02
    rng = np.random.default_rng(17)
    specials = [np.nan, np.inf, -np.inf, 0.0, -0.0, 1e200, -1e200, 1e30, -1e-3, 1e-300]
    for j in range(0, n, 2):
        bel[rng.integers(0, 36), j] = specials[rng.integers(len(specials))]
    with np.errstate(all="ignore"):
        got_v = ctx.is_valid(bel)
        ref_v = svc.is_valid(K.soa_to_aos(bel))[0]
    assert np.array_equal(got_v, ref_v)
    assert ref_v[1::2].all() and 0.1 < (~ref_v[0::2]).mean() < 0.9
    steps = rng.integers(1, 4, size=n).astype(np.int32)
    with np.errstate(all="ignore"):
        out, valid = ctx.propagate_while_valid(bel, ctrl, steps)
        ref_out, ref_valid, _ = svc.rollout(K.soa_to_aos(bel), K.soa_to_aos(ctrl), steps, nthreads=8)
        # device sincos: a step may differ only where the deciding state is inside the +-1e-6 band of a bound / half-plane
        ok = helpers.assert_rollout_mismatches_in_band(O, svc, K.soa_to_aos(bel), K.soa_to_aos(ctrl), valid, ref_valid, max_count=2)
    out = K.soa_to_aos(out)
    assert np.array_equal(out[ok][:, 4:], ref_out[ok][:, 4:], equal_nan=True)
    assert np.allclose(out[ok][:, :4], ref_out[ok][:, :4], rtol=1e-9, atol=1e-12, equal_nan=True)


@pytest.mark.parametrize("M", [0, 256])
def test_linear_diagonal_step_equals_dense_step(K, W, ctx, M, monkeypatch):
    """The rollout kernel skips the exact-zero terms of the reference's diagonal Q, R, A_cl (lin_step_diag) for beliefs
    whose covariance is finite; CCK_LIN_DENSE=1 forces the reference's dense accumulation.  Both must give the same
    values (at most the sign of an exact zero may differ), also for beliefs that carry inf / NaN / huge entries - with
    M = 0 those survive the (bounds-only) check and keep being propagated."""
    import os
    n = 20000
    bel, ctrl = W.synthetic_beliefs(n, "linear", seed=321, variant="all_survive")
    rng = np.random.default_rng(23)
    specials = [np.nan, np.inf, -np.inf, 0.0, -0.0, 1e300, -1e300, 1e200, 1e30, 1e-300, 5e-324, -1e-3]
    for j in range(0, n, 4):
        bel[rng.integers(2, 10), j] = specials[rng.integers(len(specials))]
    bel[2:, 1:2000:4] = 0.0                               # all-zero covariances
    bel[[3, 4, 7, 8], 2:4000:4] = 0.0                     # diagonal Sigma / Lambda (exact zeros stay exact zeros)
    steps = rng.integers(0, 12, size=n).astype(np.int32)
    W.install_synthetic(ctx, "linear", M, variant="all_survive")
    with np.errstate(all="ignore"):
        out_a, valid_a = ctx.propagate_while_valid(bel, ctrl, steps)
    dense = K.Context(0)
    try:
        monkeypatch.setenv("CCK_LIN_DENSE", "1")
        W.install_synthetic(dense, "linear", M, variant="all_survive")
        monkeypatch.delenv("CCK_LIN_DENSE")
        with np.errstate(all="ignore"):
            out_b, valid_b = dense.propagate_while_valid(bel, ctrl, steps)
    finally:
        dense.close()
    assert np.array_equal(valid_a, valid_b)
    assert np.array_equal(out_a, out_b, equal_nan=True)
    # bit patterns: identical except NaN payloads and the sign of exact zeros
    diff = out_a.view(np.uint64) != out_b.view(np.uint64)
    assert (np.isnan(out_a[diff]) | (out_a[diff] == 0.0)).all()
    assert (valid_a == steps).mean() > 0.5


def test_traj_T_is_enforced(K, W, ctx):
    """a trajectory buffer of traj_T steps must never be written past its end - the host entry point refuses
    max(steps) > traj_T, the device entry point refuses traj_T < 1 and the kernels drop stores beyond traj_T."""
    W.install_synthetic(ctx, "linear", 0, variant="all_survive")
    n = 256
    bel, ctrl = W.synthetic_beliefs(n, "linear", variant="all_survive")
    steps = np.full(n, 8, dtype=np.int32)
    with pytest.raises(K.CckError):
        ctx.propagate_while_valid(bel, ctrl, steps, traj_T=4)
    out, valid, traj = ctx.propagate_while_valid(bel, ctrl, steps, traj_T=8)
    assert (valid == 8).all() and np.array_equal(traj[7], out)
    import torch
    dev = helpers.torch_device()
    d_b, d_c, d_s = (torch.from_numpy(x).to(dev) for x in (bel, ctrl, steps))
    d_o = torch.empty_like(d_b); d_v = torch.empty(n, dtype=torch.int32, device=dev)
    T = 4
    d_t = torch.full(((T + 2) * 10, n), -7.0, dtype=torch.float64, device=dev)       # two guard steps behind the buffer
    with pytest.raises(K.CckError):
        ctx.propagate_while_valid_dev(None, n, d_b, n, d_c, n, d_s, d_o, n, d_v, traj=d_t, traj_ld=n, traj_T=0)
    ctx.propagate_while_valid_dev(None, n, d_b, n, d_c, n, d_s, d_o, n, d_v, traj=d_t, traj_ld=n, traj_T=T)
    ctx.synchronize()
    t = d_t.cpu().numpy().reshape(T + 2, 10, n)
    assert (t[T:] == -7.0).all(), "stores beyond traj_T"
    assert np.array_equal(t[:T], traj[:T]) and (d_v.cpu().numpy() == 8).all()
