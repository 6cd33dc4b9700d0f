"""belief-space distance (RealVectorBeliefSpace.cpp:16-62), selectNode and
findClosestWitness (ConstraintRespectingBSST.cpp:119-178).

CPU part: the oracle's restatement against closed forms (commuting covariances: the trace term
is the squared Frobenius distance of the square roots) and scipy's sqrtm.
GPU part: the batched device queries against the oracle — same operation sequence, so indices
and distances must be bit-identical.
"""
import numpy as np
import pytest


def _spd(rng, n, d, lo=0.05, hi=0.3):
    L = np.zeros((n, d, d))
    idx = np.tril_indices(d)
    L[:, idx[0], idx[1]] = rng.uniform(lo, hi, size=(n, len(idx[0])))
    return np.einsum("nij,nkj->nik", L, L).reshape(n, d * d)


def _beliefs(rng, n, dim, spread=10.0):
    W = dim + 2 * dim * dim
    b = np.zeros((n, W))
    b[:, :dim] = rng.uniform(0, spread, size=(n, dim))
    b[:, dim:dim + dim * dim] = _spd(rng, n, dim)
    b[:, dim + dim * dim:] = _spd(rng, n, dim)
    return b


@pytest.mark.parametrize("dim", [2, 4])
def test_oracle_distance_known_answers(O, dim):
    rng = np.random.default_rng(3)
    W = dim + 2 * dim * dim
    a, b = np.zeros((5, W)), np.zeros((5, W))
    a[:, :dim] = rng.uniform(0, 5, (5, dim)); b[:, :dim] = rng.uniform(0, 5, (5, dim))
    da, db = rng.uniform(0.01, 0.5, (5, dim)), rng.uniform(0.01, 0.5, (5, dim))
    for i in range(dim):     # diagonal covariances split between Sigma and Lambda
        a[:, dim + i * dim + i] = 0.25 * da[:, i]; a[:, dim + dim * dim + i * dim + i] = 0.75 * da[:, i]
        b[:, dim + i * dim + i] = 0.5 * db[:, i]; b[:, dim + dim * dim + i * dim + i] = 0.5 * db[:, i]
    ref = ((a[:, :dim] - b[:, :dim]) ** 2).sum(1) + ((np.sqrt(da) - np.sqrt(db)) ** 2).sum(1)
    assert np.allclose(O.belief_distance(a, b, dim), ref, rtol=1e-12, atol=1e-15)
    # coinciding means => 0 regardless of the covariances (RealVectorBeliefSpace.cpp:28-29)
    b2 = b.copy(); b2[:, :dim] = a[:, :dim]
    assert (O.belief_distance(a, b2, dim) == 0).all()
    # identical covariances => pure squared mean distance
    b3 = a.copy(); b3[:, :dim] += 1.0
    assert np.allclose(O.belief_distance(a, b3, dim), dim * 1.0, rtol=1e-12)


@pytest.mark.parametrize("dim", [2, 4])
def test_oracle_distance_vs_scipy_sqrtm(O, dim):
    from scipy.linalg import sqrtm
    rng = np.random.default_rng(4)
    a, b = _beliefs(rng, 40, dim), _beliefs(rng, 40, dim)
    got = O.belief_distance(a, b, dim)
    for i in range(40):
        C1 = (a[i, dim:dim + dim * dim] + a[i, dim + dim * dim:]).reshape(dim, dim)
        C2 = (b[i, dim:dim + dim * dim] + b[i, dim + dim * dim:]).reshape(dim, dim)
        s2 = np.real(sqrtm(C2))
        ref = abs(((a[i, :dim] - b[i, :dim]) ** 2).sum() + np.trace(C1 + C2 - 2 * np.real(sqrtm(s2 @ C1 @ s2))))
        assert abs(got[i] - ref) <= 1e-9 * max(1.0, ref)


def test_oracle_select_and_witness_rules(O):
    dim = 2
    nodes = np.zeros((4, 10)); nodes[:, 2] = nodes[:, 5] = 0.01; nodes[:, 6] = nodes[:, 9] = 0.01
    nodes[:, 0] = [0.0, 0.1, 0.2, 5.0]
    cost = np.array([3.0, 1.0, 1.0, 0.5])
    active = np.array([1, 1, 1, 1], dtype=np.uint8)
    s = nodes[:1].copy(); s[0, 0] = 0.05
    idx, dist = O.select_nodes(nodes, cost, active, s, dim, radius=0.2)
    assert idx[0] == 1                      # nodes 0,1,2 are within the radius; lowest cost, first index on ties
    active[1] = 0
    assert O.select_nodes(nodes, cost, active, s, dim, radius=0.2)[0][0] == 2
    s[0, 0] = 3.0                           # nothing within the radius -> nearest active
    assert O.select_nodes(nodes, cost, active, s, dim, radius=0.2)[0][0] == 3
    idx, dist = O.nearest_witness(nodes, s, dim)
    assert idx[0] == 3 and np.isclose(dist[0], 4.0)


# --------------------------------------------------------------------------- GPU parity
@pytest.mark.gpu
@pytest.mark.parametrize("dim", [2, 4])
def test_gpu_distance_bitexact(K, O, ctx, dim):
    rng = np.random.default_rng(10 + dim)
    a, b = _beliefs(rng, 5000, dim), _beliefs(rng, 5000, dim)
    b[:50, :dim] = a[:50, :dim]                          # coinciding means
    a[50:60, dim:] = 0.0                                  # zero covariance
    got = K.belief_distance(ctx, K.aos_to_soa(a), K.aos_to_soa(b), dim)
    assert np.array_equal(got, O.belief_distance(a, b, dim))


@pytest.mark.gpu
@pytest.mark.parametrize("dim", [2, 4])
def test_gpu_distance_lower_matrix(K, O, ctx, dim):
    rng = np.random.default_rng(30 + dim)
    b = _beliefs(rng, 300, dim, spread=3.0)
    b[7, :dim] = b[3, :dim]
    got = K.belief_distance_lower(ctx, K.aos_to_soa(b), dim)
    for q in (1, 7, 150, 299):
        ref = O.belief_distance(b[:q], np.repeat(b[q:q + 1], q, axis=0), dim)
        assert np.array_equal(got[q, :q], ref)
    assert got[7, 3] == 0.0
    cols = np.array([0, 3, 5, 120, 298], dtype=np.int32)
    sub = K.belief_distance_cols(ctx, K.aos_to_soa(b), dim, cols)
    assert np.array_equal(sub, got[:, cols])
    assert K.belief_distance_cols(ctx, K.aos_to_soa(b), dim, np.zeros(0, dtype=np.int32)).shape == (300, 0)
    # the sparse form the planner uses: per row the pairs below a per-row limit, in column order, count beyond the cap
    cols = np.sort(rng.choice(300, size=140, replace=False)).astype(np.int32)
    full = got[:, cols]
    lim = np.where(rng.random(300) < 0.3, np.inf, rng.uniform(0.5, 6.0, 300))
    lim[5] = 0.0
    for cap in (4, 64):
        cnt, idx, dist = K.capi.belief_distance_cols_near(ctx, K.aos_to_soa(b), dim, cols, lim, cap=cap)
        for q in range(300):
            want = [a for a in range(len(cols)) if cols[a] < q and full[q, a] < lim[q]]
            assert cnt[q] == len(want)
            m = min(len(want), cap)
            assert idx[q, :m].tolist() == want[:m]
            assert np.array_equal(dist[q, :m], full[q, want[:m]])
        assert (cnt > 4).any()
    cnt, _, _ = K.capi.belief_distance_cols_near(ctx, K.aos_to_soa(b), dim, np.zeros(0, dtype=np.int32), lim)
    assert (cnt == 0).all()


@pytest.mark.gpu
@pytest.mark.parametrize("dim", [2, 4])
def test_gpu_select_and_nearest_bitexact(K, O, ctx, dim):
    rng = np.random.default_rng(20 + dim)
    tree = K.Tree(ctx, dim)
    N, Kq = 2500, 160
    nodes = _beliefs(rng, N, dim, spread=6.0)
    cost = np.round(rng.uniform(0, 20, N), 1)            # coarse costs: plenty of ties
    first = tree.append(K.aos_to_soa(nodes[:10]), cost[:10])
    assert first == 0
    assert tree.append(K.aos_to_soa(nodes[10:]), cost[10:]) == 10 and len(tree) == N    # growth re-layout keeps the data
    active = (rng.uniform(size=N) > 0.3).astype(np.uint8)
    off = np.nonzero(active == 0)[0].astype(np.int32)
    tree.update(off, active=np.zeros(len(off), dtype=np.uint8))
    new_cost = cost.copy(); new_cost[:100] += 1.0
    tree.update(np.arange(100, dtype=np.int32), cost=new_cost[:100])
    samples = _beliefs(rng, Kq, dim, spread=6.0)
    samples[:20, :dim] = nodes[:20, :dim]                 # exact mean hits (distance 0)
    for radius in (0.2, 1.5):
        idx, dist = tree.select(K.aos_to_soa(samples), radius)
        ridx, rdist = O.select_nodes(nodes, new_cost, active, samples, dim, radius)
        assert np.array_equal(idx, ridx) and np.array_equal(dist, rdist)
    idx, dist = tree.nearest(K.aos_to_soa(samples))
    ridx, rdist = O.nearest_witness(nodes, samples, dim)
    assert np.array_equal(idx, ridx) and np.array_equal(dist, rdist)
    # empty / all-inactive trees
    t2 = K.Tree(ctx, dim)
    assert (t2.select(K.aos_to_soa(samples[:5]))[0] == -1).all() and (t2.nearest(K.aos_to_soa(samples[:5]))[0] == -1).all()
    tree.update(np.arange(N, dtype=np.int32), active=np.zeros(N, dtype=np.uint8))
    assert (tree.select(K.aos_to_soa(samples[:5]))[0] == -1).all()
    tree.clear()
    assert len(tree) == 0
    tree.close(); t2.close()


@pytest.mark.gpu
@pytest.mark.parametrize("dim", [2, 4])
@pytest.mark.parametrize("with_eligibility", [False, True])
def test_gpu_witness_batch_replays_the_sequential_rules(K, O, ctx, dim, with_eligibility):
    """cck_tree_witness_batch against a plain sequential replay of ConstraintRespectingBSST.cpp:150-178 / :250-326 built
    from the oracle's distances: survivor by survivor, every witness created so far is scanned in creation order."""
    rng = np.random.default_rng(50 + dim)
    wit = _beliefs(rng, 60, dim, spread=25.0)
    n = 700
    b = _beliefs(rng, n, dim, spread=25.0)
    b[100:140] = b[99]                                     # identical survivors (distance 0)
    b[200:260, :dim] = b[199, :dim] + rng.normal(0, 0.05, (60, dim))   # a tight cluster
    pr = 2.0
    elig = (rng.random(n) < 0.7) if with_eligibility else None
    tree = K.Tree(ctx, dim)
    tree.append(K.aos_to_soa(wit), np.zeros(len(wit)))
    widx, wdist, near, ndist, fresh = tree.witness_batch(K.aos_to_soa(b), pr, elig)
    ridx, rdist = O.nearest_witness(wit, b, dim)
    assert np.array_equal(widx, ridx) and np.array_equal(wdist, rdist)
    created = []                                           # survivors that became (persisting) witnesses, in order
    n_new = n_adopt = 0
    for s in range(n):
        cd, best = rdist[s], -1
        if created:
            d = O.belief_distance(b[created], np.repeat(b[s:s + 1], len(created), axis=0), dim)   # distance(witness, node)
            for f, df in zip(created, d):
                if df < cd:
                    cd, best = df, f
        is_new = cd > pr
        keeps = is_new and (elig is None or elig[s])
        assert bool(fresh[s]) == keeps, s
        if best >= 0 and cd <= pr:                         # the scan found a same-launch witness that matters
            assert near[s] == best and ndist[s] == cd, s
            n_adopt += 1
        else:
            assert near[s] == -1, s
        if keeps:
            created.append(s)
            n_new += 1
    assert n_new > 50 and n_adopt > 50, (n_new, n_adopt)
    tree.close()
