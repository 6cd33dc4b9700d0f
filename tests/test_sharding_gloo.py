"""CPU tests (gloo, world_size 2) of the multi-GPU host logic: block partition of the batch,
all-gather of the per-rank result records (rollout summary, expansion winner)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import cck_b200  # noqa: F401
    from cck_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(123)                       # same stream on both ranks: the global batch
    cost = rng.uniform(0, 10, n)
    cost[[5, n - 3]] = cost.min() - 1.0                    # a tie across shards: lowest global index must win
    ok = rng.uniform(size=n) < 0.4
    ok[[5, n - 3]] = True
    steps = rng.integers(1, 11, n)
    valid = np.minimum(steps, rng.integers(0, 12, n))
    lo, hi = sharding.shard_bounds(n, rank, world)
    # what the device kernels hand each rank (k_winner / k_summary), restated with numpy for this host-logic test
    lc = np.where(ok[lo:hi], cost[lo:hi], np.inf)
    li = int(np.argmin(lc)) if np.isfinite(lc).any() else -1
    local_w = torch.tensor([lc[li] if li >= 0 else np.inf, float(lo + li) if li >= 0 else -1.0], dtype=torch.float64)
    local_s = torch.tensor([int((valid[lo:hi] == steps[lo:hi]).sum()), int(np.minimum(valid[lo:hi] + 1, steps[lo:hi]).sum()), int(valid[lo:hi].sum())],
                           dtype=torch.int64)
    w = sharding.gather_winner(local_w)
    s = sharding.gather_summary(local_s)
    # sharded conflict search: rank 0 found (step 40, pair rank 3) in its range of steps, rank 1 (step 12, pair rank 7) and,
    # in a second plan, nothing; a third plan has the same step on both ranks - the lower pair rank wins
    key = lambda t, r: (t << 32) | r       # noqa: E731
    c1 = sharding.gather_conflict(torch.tensor([key(40, 3), 2, 5, 40, 6] if rank == 0 else [key(12, 7), 1, 9, 12, 2], dtype=torch.int64))
    c2 = sharding.gather_conflict(torch.tensor([key(40, 3), 2, 5, 40, 6] if rank == 0 else [-1, -1, -1, -1, 0], dtype=torch.int64))
    c3 = sharding.gather_conflict(torch.tensor([key(7, 9), 4, 6, 7, 1] if rank == 0 else [key(7, 2), 0, 3, 7, 5], dtype=torch.int64))
    c4 = sharding.gather_conflict(torch.tensor([-1, -1, -1, -1, 0], dtype=torch.int64))
    assert c1 == (1, 9, 12, 2) and c2 == (2, 5, 40, 6) and c3 == (0, 3, 7, 5) and c4 is None, (c1, c2, c3, c4)
    gc = np.where(ok, cost, np.inf)
    q.put((rank, lo, hi, w, s.tolist(), (float(gc.min()), int(np.argmin(gc))),
           [int((valid == steps).sum()), int(np.minimum(valid + 1, steps).sum()), int(valid.sum())]))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds_cover_exactly():
    import cck_b200  # noqa: F401
    from cck_b200 import sharding
    for n in (0, 1, 7, 8, 1000, 1 << 20):
        for world in (1, 2, 3, 4, 8):
            b = [sharding.shard_bounds(n, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
    assert [sharding.child_to_rank(c, 2) for c in range(4)] == [0, 1, 0, 1]


def test_gather_records_world2():
    world, n, port = 2, 1001, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, lo, hi, w, s, ref_w, ref_s in res:
        assert w[1] == 5 and w == ref_w, (w, ref_w)      # tie broken towards the lowest global index
        assert s == ref_s


def test_single_process_gather_is_identity():
    import cck_b200  # noqa: F401
    from cck_b200 import sharding
    assert sharding.gather_winner(torch.tensor([3.5, 17.0], dtype=torch.float64)) == (3.5, 17)
    assert sharding.gather_winner(torch.tensor([float("inf"), -1.0], dtype=torch.float64)) == (float("inf"), -1)
    assert sharding.gather_summary(torch.tensor([1, 2, 3])).tolist() == [1, 2, 3]
    assert sharding.gather_conflict(torch.tensor([(3 << 32) | 1, 0, 2, 3, 4], dtype=torch.int64)) == (0, 2, 3, 4)
    assert sharding.gather_conflict(torch.tensor([-1, -1, -1, -1, 0], dtype=torch.int64)) is None
