"""CPU-only tests of the fp32 box + cell-grid broad phase of the PCCBlackmore rollout kernels
(csrc/cck_rollout_grid.cuh): the table is built by host code (cck_build_grid_table needs no GPU), and the device's
fp32 query is re-enacted here in numpy float32 arithmetic (IEEE round-to-nearest, like the kernel's FADD/FMUL).

What must hold for verdicts to stay bit-identical to PCCBlackmoreSVC::isValid (PCCBlackmoreSVC.cpp:54-75):
  1. every obstacle with a finite box is binned into exactly the cell the monotone cell function assigns to it,
  2. outer boxes contain, inner boxes are contained in, the exact thresholds b/a with at least one fp32 ulp to spare,
  3. "interval misses the outer box" (proven safe) implies the reference's fp64 predicate is TRUE,
     "interval inside the inner box" (proven unsafe) implies it is FALSE,
  4. every obstacle that is not proven safe lies in the queried cell window (or in the always-list).
"""
import numpy as np
import pytest

f32 = np.float32


def _tables(K, W, which, rng):
    rad = K.rect_robot_bounding_radius()
    if which == "synthetic":
        centres = W.synthetic_obstacles(256, W.SEED, "realistic")
        A, B, _ = K.build_obstacle_tables(centres, rad)
    elif which == "far":
        centres = W.synthetic_obstacles(256, W.SEED, "all_survive")
        A, B, _ = K.build_obstacle_tables(centres, rad)
    elif which == "dense32":
        cells = rng.choice(32 * 32, size=102, replace=False)
        centres = np.stack([cells % 32, cells // 32], axis=1).astype(np.float64)
        A, B, _ = K.build_obstacle_tables(centres, rad)
    else:   # arbitrary: axis-aligned boxes of random size/scale, some rotated, some with non-finite or duplicated planes
        M = 300
        A, B = np.zeros((M, 4, 2)), np.zeros((M, 4))
        for o in range(M):
            cx, cy = rng.uniform(0, 63, 2)
            hw, hh = rng.uniform(0.2, 1.5, 2)
            w = rng.uniform(0.5, 3.0)
            th = rng.uniform(0, np.pi) if o % 5 == 0 else 0.0
            ct, st = (1.0, 0.0) if th == 0.0 else (np.cos(th), np.sin(th))
            for i, ((nx, ny), e) in enumerate(zip([(st, -ct), (ct, st), (-st, ct), (-ct, -st)], [hh, hw, hh, hw])):
                A[o, i] = (w * nx, w * ny)
                B[o, i] = w * (nx * cx + ny * cy + e)
        B[3, 1] = np.inf; B[7, 0] = -np.inf; B[11, 2] = np.nan
        A[13] = [[2.0, 0.0], [1.5, 0.0], [0.0, 1.0], [0.0, -1.0]]; B[13] = [60.0, 46.5, 20.0, -18.0]
    return A, B


def _exact_safe(A, B, x, y, P, c):
    """HyperplaneCCValidityChecker_ per (belief, obstacle): [n, M] bool, float64 in the reference's operation order."""
    a0, a1 = A[None, :, :, 0], A[None, :, :, 1]                       # [1, M, 4]
    P0, P1, P2, P3 = (P[:, k][:, None, None] for k in range(4))
    with np.errstate(all="ignore"):
        t0 = a0 * P0 + a1 * P2
        t1 = a0 * P1 + a1 * P3
        tmp = t0 * a0 + t1 * a1
        bb = np.sqrt(np.float64(2.0)) * np.sqrt(tmp) * c
        ok = (x[:, None, None] * a0 + y[:, None, None] * a1) >= (B[None, :, :] + bb)
    return ok.any(axis=2)


def _device_query(t, x, y, P, sqrt_err=0.0):
    """The kernel's fp32 prologue (grid_check), in numpy float32.  sqrt_err perturbs the approximate sqrt."""
    xf, yf = x.astype(f32), y.astype(f32)
    p0, p3 = P[:, 0].astype(f32), P[:, 3].astype(f32)
    offd = (np.abs(P[:, 1]) + np.abs(P[:, 2])).astype(f32)
    with np.errstate(all="ignore"):
        fin = ((np.abs(xf) + np.abs(yf)) + (p0 + p3) + offd < f32(1e30)) & (p0 >= 0) & (p3 >= 0)
        s0 = (np.sqrt(p0) * f32(1.0 + sqrt_err)).astype(f32)
        s3 = (np.sqrt(p3) * f32(1.0 + sqrt_err)).astype(f32)
        ex, ey = f32(t["k"]) * s0, f32(t["k"]) * s3
        # __fmaf_rn: the product of two fp32 numbers is exact in fp64; one rounding to fp32 after the add
        mx = ((np.abs(xf) + ex).astype(np.float64) * np.float64(f32(3.814697265625e-06)) + np.float64(f32(1e-17))).astype(f32)
        my = ((np.abs(yf) + ey).astype(np.float64) * np.float64(f32(3.814697265625e-06)) + np.float64(f32(1e-17))).astype(f32)
        wx, wy = ex + mx, ey + my
        xm, xp, ym, yp = xf - wx, xf + wx, yf - wy, yf + wy
        mx2, my2 = mx + mx, my + my
    q = {"fin": fin, "xm": xm, "xp": xp, "ym": ym, "yp": yp,
         "xmi": xm + mx2, "xpi": xp - mx2, "ymi": ym + my2, "ypi": yp - my2}
    return q


def _cell(v, g0, ic):
    with np.errstate(all="ignore"):
        return np.floor((v.astype(f32) - f32(g0)) * f32(ic))


def _sub_rd(a, b):
    """fp32 subtraction rounded towards -inf (__fsub_rd)."""
    with np.errstate(all="ignore"):
        exact = a.astype(np.float64) - np.float64(b)
        r = exact.astype(f32)
        r = np.where(r.astype(np.float64) > exact, np.nextafter(r, f32(-np.inf)), r)
    return r.astype(f32)


@pytest.mark.parametrize("which", ["synthetic", "far", "dense32", "arbitrary"])
def test_grid_table_invariants(K, W, which):
    rng = np.random.default_rng(3)
    A, B = _tables(K, W, which, rng)
    M = len(A)
    c = 2.2
    t = K.capi.build_grid_table(A, B, c)
    assert t["M"] == M and sorted(t["order"].tolist()) == list(range(M))
    ng = t["n_grid"]
    outer, inner, order = t["outer"], t["inner"], t["order"]
    assert np.isfinite(outer[:ng]).all() and not np.isfinite(outer[ng:]).all(axis=1).any()
    assert float(t["k"]) >= np.sqrt(2.0) * c
    # 1. binning: cell function, dense row-major CSR range
    cx = _cell(outer[:ng, 0], t["gx0"], t["icx"]).astype(int)
    cy = _cell(outer[:ng, 1], t["gy0"], t["icy"]).astype(int)
    assert (cx >= 0).all() and (cx <= 63).all() and (cy >= 0).all() and (cy < t["R"]).all()
    ds = t["dstart"].astype(int)
    assert len(ds) == t["R"] * 64 + 1 and (np.diff(ds) >= 0).all() and ds[0] == 0 and ds[-1] == ng
    for j in range(ng):
        idx = int(cy[j]) * 64 + int(cx[j])
        assert ds[idx] <= j < ds[idx + 1]
    if ng:
        assert float(t["SX"]) >= (outer[:ng, 2].astype(np.float64) - outer[:ng, 0]).max()
        assert float(t["SY"]) >= (outer[:ng, 3].astype(np.float64) - outer[:ng, 1]).max()
    # 2. box sides vs the exact thresholds b/a of the axis-aligned planes, with >= 1 ulp to spare.  Several planes of
    #    the same direction are OR-ed by the reference, so the weakest threshold (min for +, max for -) defines the side.
    for j in range(M):
        o = order[j]
        thr = {"+x": [], "-x": [], "+y": [], "-y": []}
        for i in range(4):
            a0, a1, b = A[o, i, 0], A[o, i, 1], B[o, i]
            if np.isnan(b) or not (np.isfinite(a0) and np.isfinite(a1)):
                continue
            if a1 == 0 and 1e-100 < abs(a0) < 1e100:
                thr["+x" if a0 > 0 else "-x"].append(b / a0)
            elif a0 == 0 and 1e-100 < abs(a1) < 1e100:
                thr["+y" if a1 > 0 else "-y"].append(b / a1)
        with np.errstate(all="ignore"):
            for key, col, up in (("+x", 2, True), ("+y", 3, True), ("-x", 0, False), ("-y", 1, False)):
                side = outer[j, col]
                if not thr[key]:
                    assert side == (np.inf if up else -np.inf)
                    continue
                e = min(thr[key]) if up else max(thr[key])
                if up:
                    assert float(side) >= e and (np.isinf(side) or float(np.nextafter(side, f32(-np.inf))) >= e)
                else:
                    assert float(side) <= e and (np.isinf(side) or float(np.nextafter(side, f32(np.inf))) <= e)
        if np.isfinite(inner[j]).all():      # inner box strictly inside the outer one
            assert inner[j, 0] > outer[j, 0] and inner[j, 1] > outer[j, 1] and inner[j, 2] < outer[j, 2] and inner[j, 3] < outer[j, 3]


@pytest.mark.parametrize("which", ["synthetic", "dense32", "arbitrary"])
@pytest.mark.parametrize("sqrt_err", [0.0, 2.0 ** -22, -(2.0 ** -22)])
def test_fp32_broad_phase_never_contradicts_the_fp64_predicate(K, W, which, sqrt_err):
    rng = np.random.default_rng(11)
    A, B = _tables(K, W, which, rng)
    M = len(A)
    c = 2.2
    t = K.capi.build_grid_table(A, B, c)
    outer, inner, order, ng = t["outer"], t["inner"], t["order"], t["n_grid"]
    Ao, Bo = A[order], B[order]                                      # rows in table order
    n = 40000
    span = 64.0 if which != "dense32" else 32.0
    x, y = rng.uniform(-1.5, span + 0.5, n), rng.uniform(-1.5, span + 0.5, n)
    Lc = rng.uniform(0.01, 0.35, (n, 3))
    P = np.stack([Lc[:, 0] ** 2, Lc[:, 0] * Lc[:, 1], Lc[:, 0] * Lc[:, 1], Lc[:, 1] ** 2 + Lc[:, 2] ** 2], axis=1)
    P[::11] *= 1e-6                                                  # tiny covariances
    # a third of the beliefs sit within a few fp32 ulps of a box side widened by their own (ex, ey)
    kf = float(t["k"])
    sel = rng.integers(0, max(ng, 1), n)
    for side in range(4):
        idx = np.arange(side, n, 12)
        if ng == 0:
            break
        jj = sel[idx]
        e = kf * np.sqrt(P[idx, 0] if side in (0, 2) else P[idx, 3])
        jitter = rng.uniform(-3e-5, 3e-5, len(idx))
        if side == 0:
            x[idx] = outer[jj, 2] + e + jitter; y[idx] = (outer[jj, 1].astype(np.float64) + outer[jj, 3]) / 2
        elif side == 2:
            x[idx] = outer[jj, 0] - e + jitter; y[idx] = (outer[jj, 1].astype(np.float64) + outer[jj, 3]) / 2
        elif side == 1:
            y[idx] = outer[jj, 3] + e + jitter; x[idx] = (outer[jj, 0].astype(np.float64) + outer[jj, 2]) / 2
        else:
            y[idx] = outer[jj, 1] - e + jitter; x[idx] = (outer[jj, 0].astype(np.float64) + outer[jj, 2]) / 2
    q = _device_query(t, x, y, P, sqrt_err)
    fin = q["fin"]
    assert fin.all()
    exact = _exact_safe(Ao, Bo, x, y, P, c)                          # [n, M]
    with np.errstate(all="ignore"):
        proven_safe = ((q["xm"][:, None] >= outer[None, :, 2]) | (q["xp"][:, None] <= outer[None, :, 0]) |
                       (q["ym"][:, None] >= outer[None, :, 3]) | (q["yp"][:, None] <= outer[None, :, 1]))
        proven_unsafe = ((q["xmi"][:, None] < inner[None, :, 2]) & (q["xpi"][:, None] > inner[None, :, 0]) &
                         (q["ymi"][:, None] < inner[None, :, 3]) & (q["ypi"][:, None] > inner[None, :, 1]))
    # 3. the broad phase may only prove what the fp64 predicate says
    assert not (proven_safe & ~exact).any()
    assert not (proven_unsafe & exact).any()
    assert not (proven_safe & proven_unsafe).any()
    undecided = ~proven_safe & ~proven_unsafe
    if which == "arbitrary":      # a fifth of those obstacles is rotated: no box, always decided by the exact expression
        assert proven_safe.mean() > 0.7 and proven_unsafe.any() and undecided[:, np.isfinite(outer).all(axis=1)].mean() < 0.01
    else:
        assert proven_safe.mean() > 0.9 and proven_unsafe.any() and undecided.mean() < 0.01
    # 4. every obstacle that is not proven safe is found through the cell window (or sits in the always-list)
    clo = np.maximum(_cell(_sub_rd(q["xm"], t["SX"]), t["gx0"], t["icx"]), 0)
    chi = np.minimum(_cell(q["xp"], t["gx0"], t["icx"]), 63)
    rlo = np.maximum(_cell(_sub_rd(q["ym"], t["SY"]), t["gy0"], t["icy"]), 0)
    rhi = np.minimum(_cell(q["yp"], t["gy0"], t["icy"]), t["Rm1"])
    if ng:
        cx = _cell(outer[:ng, 0], t["gx0"], t["icx"])
        cy = _cell(outer[:ng, 1], t["gy0"], t["icy"])
        in_window = ((cx[None, :] >= clo[:, None]) & (cx[None, :] <= chi[:, None]) &
                     (cy[None, :] >= rlo[:, None]) & (cy[None, :] <= rhi[:, None]))
        assert not (~proven_safe[:, :ng] & ~in_window).any()
        # the kernel's enumeration (one contiguous CSR range per grid row of the window) visits exactly the window
        ds = t["dstart"].astype(int)
        for i in rng.choice(n, 400, replace=False):
            got = []
            if clo[i] <= chi[i] and rlo[i] <= rhi[i]:
                for r in range(int(rlo[i]), int(rhi[i]) + 1):
                    got.extend(range(ds[r * 64 + int(clo[i])], ds[r * 64 + int(chi[i]) + 1]))
            assert sorted(got) == np.nonzero(in_window[i])[0].tolist()
        # the window is not vacuous: it skips most of the table
        assert in_window.mean() < 0.05


@pytest.mark.parametrize("which", ["synthetic", "far", "dense32"])
def test_clearance_field_never_contradicts_the_fp64_predicate(K, W, which):
    """The FIRST test of a belief-step: one byte of the clearance field proves "safe w.r.t. every obstacle" when
    (16 sqrt2 c sqrt(max(P00, P11)))^2 < q^2.  Re-enacted in numpy float32: whatever it proves must be true for the reference's
    fp64 predicate, and it must not be vacuous."""
    rng = np.random.default_rng(17)
    A, B = _tables(K, W, which, rng)
    c = 2.378449440526247
    span = 64.0 if which != "dense32" else 32.0
    low, high = np.array([-1.0, -1.0]), np.array([span, span])
    df = K.capi.build_clearance_field(A, B, c, low, high)
    assert df is not None
    q, icx, icy, k2 = df["q"], f32(df["icx"]), f32(df["icy"]), f32(df["k2"])
    n = 60000
    x, y = rng.uniform(-1.0, span, n), rng.uniform(-1.0, span, n)
    Lc = rng.uniform(0.01, 0.3, (n, 3))
    P = np.stack([Lc[:, 0] ** 2, Lc[:, 0] * Lc[:, 1], Lc[:, 0] * Lc[:, 1], Lc[:, 1] ** 2 + Lc[:, 2] ** 2], axis=1)
    P[::3] *= 0.05
    # adversarial: means a few fp32 ulps from cell borders, covariances a hair from the cell's threshold
    blx = np.nextafter(f32(low[0]), f32(np.inf)); bly = np.nextafter(f32(low[1]), f32(np.inf))
    k = np.arange(0, n, 4)
    x[k] = np.float64(blx) + rng.integers(1, K.capi.DF_N, len(k)) / np.float64(icx) + rng.uniform(-2e-5, 2e-5, len(k))
    xf, yf = x.astype(f32), y.astype(f32)
    bhx, bhy = np.nextafter(f32(high[0]), f32(-np.inf)), np.nextafter(f32(high[1]), f32(-np.inf))
    pre = (xf > blx) & (xf < bhx) & (yf > bly) & (yf < bhy)
    with np.errstate(all="ignore"):
        cx = ((xf - blx) * icx).astype(f32).astype(np.int64)
        cy = ((yf - bly) * icy).astype(f32).astype(np.int64)
    assert (cx[pre] >= 0).all() and (cx[pre] < K.capi.DF_N).all() and (cy[pre] >= 0).all() and (cy[pre] < K.capi.DF_N).all()
    qq = np.where(pre, q[np.clip(cy, 0, 127), np.clip(cx, 0, 127)], 0).astype(f32)
    j = np.arange(1, n, 4)                                   # variance right at the threshold of its cell (+- 1e-5 relative)
    thr = (qq[j].astype(np.float64) ** 2) / np.float64(k2)
    s = thr / np.maximum(np.maximum(P[j, 0], P[j, 3]), 1e-300) * rng.uniform(1 - 1e-5, 1 + 1e-5, len(j))
    P[j] *= np.where(thr > 0, s, 1.0)[:, None]
    p0, p3 = P[:, 0].astype(f32), P[:, 3].astype(f32)
    proven = pre & ((np.maximum(p0, p3) * k2).astype(f32) < qq * qq)
    exact = _exact_safe(A, B, x, y, P, c).all(axis=1)
    assert not (proven & ~exact).any()
    if which == "far":
        rest = pre.copy(); rest[j] = False         # (the threshold-hugging quarter fails half of the time by construction)
        assert proven[rest].mean() > 0.99          # obstacles beyond the saturation distance: everything is proven at once
    else:
        assert proven.mean() > 0.1 and (exact & ~proven).any()
    # no field: rotated obstacle, no obstacles
    A2, B2 = A.copy(), B.copy()
    A2[0] = [[0.3, -0.4], [0.4, 0.3], [-0.3, 0.4], [-0.4, -0.3]]
    assert K.capi.build_clearance_field(A2, B2, c, low, high) is None
    assert K.capi.build_clearance_field(A[:0], B[:0], c, low, high) is None
