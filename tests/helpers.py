"""Shared helpers for the parity tests: build matching oracle / product configurations."""
import numpy as np


def oracle_synth(O, W, M, variant, kind, dim, k_agents=4, p_safe=0.9, seed=None):
    centres = W.synthetic_obstacles(M, W.SEED if seed is None else seed, variant)
    world = O.World.synthetic(centres, float(W.WORLD), float(W.WORLD), k_agents, p_safe)
    return world, O.Svc(world, kind, dim)


def install_from_oracle_svc(K, ctx, svc, model):
    """Configure the product context with the ORACLE's tables (isolates kernel parity from
    table-builder parity)."""
    ctx.set_model(model)
    ctx.set_obstacles(svc.kind, svc.low, svc.high, svc.A, svc.B, svc.centres, svc.rects,
                      erf_inv_result=svc.erf_inv_result, p_collision=svc.p_collision, sc=svc.sc, max_rad=svc.max_rad)


def band_mask_halfplane(svc, aos, cov2, c, band=1e-6):
    """True where some decision inequality of the belief is within `band` of equality
    (the stated +-1e-6 band around the chance threshold).  aos: [n, W] means in cols 0,1;
    cov2: [n, 4] = (Sigma+Lambda) 2x2 block."""
    x, y = aos[:, 0], aos[:, 1]
    near = np.zeros(len(aos), dtype=bool)
    for o in range(len(svc.A)):
        for i in range(4):
            a0, a1 = svc.A[o, i]
            q = a0 * (a0 * cov2[:, 0] + a1 * cov2[:, 2]) + a1 * (a0 * cov2[:, 1] + a1 * cov2[:, 3])
            rhs = svc.B[o, i] + np.sqrt(2) * np.sqrt(np.maximum(q, 0)) * c
            near |= np.abs(x * a0 + y * a1 - rhs) < band
    return near
