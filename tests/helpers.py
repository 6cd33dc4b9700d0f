"""Shared helpers for the parity tests: build matching oracle / product configurations."""
import os

import numpy as np

HOSTSIM = os.environ.get("CCK_HOSTSIM", "") == "1"      # tests/hostsim: the kernels run on the CPU, "device" memory is host memory


def torch_device():
    """Where the device-pointer entry points expect their tensors: cuda:0 (tests/hostsim: plain host memory)."""
    import torch
    return torch.device("cpu") if HOSTSIM else torch.device("cuda", 0)


def torch_sync():
    import torch
    if not HOSTSIM:
        torch.cuda.synchronize()


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


def cov2_of(aos, dim):
    """(Sigma + Lambda).block<2,2>(0,0) of AoS records - for 4-D states the (x, xdot) block (PCCBlackmoreSVC.cpp:41)."""
    S, L = aos[:, dim:dim + dim * dim], aos[:, dim + dim * dim:]
    idx = [0, 1, dim, dim + 1]
    return np.stack([S[:, i] + L[:, i] for i in idx], axis=1)


def adaptive_c(O, svc, aos):
    """AdaptiveRiskBlackmoreSVC's per-state erf_inv argument (AdaptiveRiskBlackmoreSVC.cpp:69-92, eta_list[0] quirk)."""
    out = np.zeros(len(aos))
    for i, b in enumerate(aos):
        d = np.sqrt(((b[:2] - svc.centres) ** 2).sum(axis=1))
        alpha = 1 / (1 / d).sum()
        out[i] = O.erf_inv(1 - 2 * ((alpha / d[0]) * svc.p_collision))
    return out


def svc_band(O, svc, aos, band=1e-6):
    """True where a state sits inside the stated +-1e-6 band around SOME decision threshold of its state checker: a state
    bound, a half-plane inequality (Blackmore / Adaptive) or the chi-squared decagon's clearance (verdict flips when the
    robot radius moves by +-band).  Only such states may get a different verdict from the CUDA path."""
    aos = np.asarray(aos, dtype=np.float64)
    dim = svc.dim
    near = np.zeros(len(aos), dtype=bool)
    for d in range(dim):
        near |= (np.abs(aos[:, d] - svc.low[d]) < band) | (np.abs(aos[:, d] - svc.high[d]) < band)
    if len(svc.A) == 0:
        return near
    P = cov2_of(aos, dim)
    if svc.kind == O.SVC_BLACKMORE:
        near |= band_mask_halfplane(svc, aos, P, svc.erf_inv_result, band)
    elif svc.kind == O.SVC_ADAPTIVE:
        c = adaptive_c(O, svc, aos)
        for i in range(len(aos)):
            near[i] |= band_mask_halfplane(svc, aos[i:i + 1], P[i:i + 1], c[i], band)[0] or not np.isfinite(c[i])
    else:
        v0, _ = svc.is_valid(aos)
        r0 = svc.max_rad
        for dr in (-band, band):
            svc.set_max_rad(r0 + dr)
            v, _ = svc.is_valid(aos)
            near |= v != v0
        svc.set_max_rad(r0)
    return near


def assert_rollout_mismatches_in_band(O, svc, bel_aos, ctrl_aos, valid_got, valid_ref, max_count=None):
    """Every belief whose first-violation step differs from the oracle's must have its DECIDING state (the first state on
    which the two verdicts differ) inside the +-1e-6 band; returns the mask of agreeing beliefs."""
    mism = np.nonzero(np.asarray(valid_got) != np.asarray(valid_ref))[0]
    for i in mism:
        t = int(min(valid_got[i], valid_ref[i]))
        state = O.propagate_n(bel_aos[i], ctrl_aos[i], t + 1, svc.dim)[-1:]
        assert svc_band(O, svc, state)[0], f"belief {i}: verdicts differ at step {t} outside the +-1e-6 band"
    if max_count is not None:
        assert len(mism) <= max_count, f"{len(mism)} verdict mismatches"
    ok = np.ones(len(valid_ref), dtype=bool)
    ok[mism] = False
    return ok
