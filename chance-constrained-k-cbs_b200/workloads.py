"""Seeded synthetic inputs for the hot path (BASELINE.json config 5 and the map-based configs).

Pure numpy input generation shared by tests/ and bench.py; it computes nothing of the path
itself.  Generator: numpy Philox (64-bit counter-based), key = seed (default 20240501).
"""
import numpy as np

from . import capi

SEED = 20240501
WORLD = 64           # synthetic world is 64 x 64 cells, bounds [-1, 64]^2 (OmplSetUp.cpp:195-200 convention)
ROBOT_LEN = 0.25     # RectangularRobot(…, 0.25, 0.25)  (Instance.cpp:166)


def _rng(seed, stream=0):
    return np.random.Generator(np.random.Philox(key=int(seed) + (int(stream) << 32)))


VARIANTS = ("all_survive", "dense_survive", "realistic")
# dense_survive: clearance (Chebyshev distance mean -> obstacle centre) that keeps every belief valid for 50 steps:
# inflated half-side 0.5 + 0.177, plus sqrt2 * c * sqrt(P) <= 1.01 (c = 2.38 at M = 256, P <= 0.09), plus 50 * 0.2 * 0.05
# of travel, plus slack
DENSE_CLEARANCE = 2.25
DENSE_CTRL = 0.05


def synthetic_obstacles(M=256, seed=SEED, variant="realistic"):
    """M distinct integer cells of the 64x64 world (unit squares).  variant 'all_survive'
    translates the field to x >= 1000 so no belief ever meets an obstacle; 'dense_survive' and 'realistic' keep
    it inside the world."""
    rng = _rng(seed, 1)
    cells = rng.choice(WORLD * WORLD, size=M, replace=False)
    centres = np.stack([cells % WORLD, cells // WORLD], axis=1).astype(np.float64)
    if variant == "all_survive":
        centres[:, 0] += 1000.0
    return centres


def _spd(rng, n, d):
    """Sigma = L L^T with L lower triangular, entries ~ U(0.05, 0.15) (SPD, non-isotropic)."""
    L = np.zeros((n, d, d))
    idx = np.tril_indices(d)
    L[:, idx[0], idx[1]] = rng.uniform(0.05, 0.15, size=(n, len(idx[0])))
    return np.einsum("nij,nkj->nik", L, L).reshape(n, d * d)


def synthetic_beliefs(n, model="linear", seed=SEED, variant="realistic", offset=0):
    """SoA belief batch [W, n] and controls [dim, n].  `offset` selects a disjoint RNG stream
    (rank sharding: rank r uses offset=r)."""
    if variant == "dense_survive":
        return _dense_survive_beliefs(n, model, seed, offset)
    rng = _rng(seed, 2 + 16 * offset)
    dim = 2 if model == "linear" else 4
    W = capi.width(dim)
    bel = np.empty((W, n))
    # all_survive keeps the means >= 11 cells inside the bounds: 50 steps * 0.2 * |u|<=1 moves <= 10
    lo, hi = (0.0, WORLD - 1.0) if variant == "realistic" else (11.0, WORLD - 12.0)
    bel[0] = rng.uniform(lo, hi, n)
    bel[1] = rng.uniform(lo, hi, n)
    if dim == 4:
        bel[2] = rng.uniform(-np.pi, np.pi, n)
        bel[3] = rng.uniform(0.1, 1.0, n)
    bel[dim:dim + dim * dim] = _spd(rng, n, dim).T
    bel[dim + dim * dim:] = _spd(rng, n, dim).T
    if dim == 2:
        ctrl = rng.uniform(-1.0, 1.0, size=(2, n))
    else:
        ctrl = np.stack([rng.uniform(-1.0, WORLD, n), rng.uniform(-1.0, WORLD, n), rng.uniform(-np.pi, np.pi, n),
                         rng.uniform(0.01, 10.0, n)])
    return np.ascontiguousarray(bel), np.ascontiguousarray(ctrl)


def _dense_survive_beliefs(n, model, seed, offset, M=256):
    """SURVEY 8(d) variant (ii) with the obstacle field INSIDE the world: every belief survives all 50 steps (so a launch is
    exactly B*T belief-steps) but lives among the obstacles, so the broad phase finds candidates and the box tests run on
    surviving beliefs.  Means are uniform over the part of the world (rastered at 1/8 cell) whose Chebyshev distance to every obstacle centre is
    >= DENSE_CLEARANCE (about 30 % of the world); the control is small (|u| <= DENSE_CTRL: <= 0.5 cells of travel)."""
    if model != "linear":
        raise ValueError("dense_survive is defined for the linear model only (a unicycle cannot hold its position)")
    rng = _rng(seed, 3 + 16 * offset)
    dim = 2
    W = capi.width(dim)
    centres = synthetic_obstacles(M, seed, "dense_survive")
    # free-space raster at 1/8 cell: a sub-cell is allowed when even its farthest point keeps the clearance
    R = 8
    lo, hi = 3.0, WORLD - 4.0
    ns = int((hi - lo) * R)
    sub = lo + (np.arange(ns) + 0.5) / R                        # sub-cell centres
    allowed = np.ones((ns, ns), dtype=bool)
    lim = DENSE_CLEARANCE + 0.5 / R
    for ox, oy in centres:
        ix = np.nonzero(np.abs(sub - ox) < lim)[0]
        iy = np.nonzero(np.abs(sub - oy) < lim)[0]
        if len(ix) and len(iy):
            allowed[ix[0]:ix[-1] + 1, iy[0]:iy[-1] + 1] = False
    free = np.flatnonzero(allowed)
    pick = free[rng.integers(0, len(free), n)]
    bel = np.empty((W, n))
    bel[0] = lo + (pick // ns + rng.uniform(0.0, 1.0, n)) / R
    bel[1] = lo + (pick % ns + rng.uniform(0.0, 1.0, n)) / R
    bel[dim:dim + dim * dim] = _spd(rng, n, dim).T
    bel[dim + dim * dim:] = _spd(rng, n, dim).T
    ctrl = rng.uniform(-DENSE_CTRL, DENSE_CTRL, size=(2, n))
    return np.ascontiguousarray(bel), np.ascontiguousarray(ctrl)


def bounds(dim, x_max=float(WORLD), y_max=float(WORLD)):
    if dim == 2:
        return np.array([-1.0, -1.0]), np.array([x_max, y_max])
    return np.array([-1.0, -1.0, -np.pi, 0.01]), np.array([x_max, y_max, np.pi, 10.0])


def install_synthetic(ctx, model="linear", M=256, seed=SEED, variant="realistic", svc_kind=capi.SVC_BLACKMORE, k_agents=4,
                      p_safe=0.9, dt=0.2):
    """Configure a Context for the synthetic sweep using the library's own host-side builders
    (risk split, Minkowski half-planes, erf_inv).  Returns a dict describing the tables."""
    dim = 2 if model == "linear" else 4
    ctx.set_model(capi.MODEL_LINEAR2D if dim == 2 else capi.MODEL_UNICYCLE, dt)
    centres = synthetic_obstacles(M, seed, variant)
    rad = capi.rect_robot_bounding_radius(0.0, 0.0, ROBOT_LEN, ROBOT_LEN)
    A, B, rects = capi.build_obstacle_tables(centres, rad)
    p_safe_obs, p_safe_agents = capi.split_risk(p_safe, M, k_agents)
    low, high = bounds(dim)
    p_coll = 1 - p_safe_obs
    c = capi.erf_inv(1 - 2 * (p_coll / M)) if M > 0 else 0.0
    sc = capi.chi2_quantile_2dof(0.9)
    ctx.set_obstacles(svc_kind, low, high, A, B, centres, rects, erf_inv_result=c, p_collision=p_coll, sc=sc, max_rad=rad)
    return {"centres": centres, "A": A, "B": B, "rects": rects, "p_safe_obs": p_safe_obs, "p_safe_agents": p_safe_agents,
            "erf_inv_result": c, "robot_rad": rad, "low": low, "high": high, "dim": dim}
