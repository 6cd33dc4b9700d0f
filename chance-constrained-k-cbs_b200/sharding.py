"""Multi-GPU host logic: independent units (beliefs / candidates / CBS child replans) are
block-partitioned over the ranks of one node; tables are replicated; no data-path collective.
Only per-rank RESULT RECORDS cross NVLink (torch.distributed, NCCL on GPUs; gloo in CPU tests):
the rollout summary (3 x int64), the expansion winner (cost, global index) and the earliest conflict of a plan
(key, a, b, first_step, run_len)."""
import torch
import torch.distributed as dist


def shard_bounds(n, rank, world):
    """Contiguous block partition of n units: (lo, hi) of `rank`; sizes differ by at most one."""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def child_to_rank(child_index, world):
    """K-CBS: the two independent child replans of an expansion (KCBS.cpp:417-467) go to different GPUs."""
    return child_index % world


def gather_summary(local3, group=None):
    """all-gather of per-rank [n_full, executed_steps, sum_valid] int64 records -> summed totals (tensor[3])."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local3.clone()
    parts = [torch.zeros_like(local3) for _ in range(dist.get_world_size(group))]
    dist.all_gather(parts, local3, group=group)
    return torch.stack(parts).sum(dim=0)


def gather_winner(local2, group=None):
    """all-gather of per-rank (cost, global_index) float64 records -> the global winner:
    smallest cost, ties to the smallest global index; (inf, -1) when no rank has a candidate."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        parts = [torch.zeros_like(local2) for _ in range(dist.get_world_size(group))]
        dist.all_gather(parts, local2, group=group)
        recs = torch.stack(parts).cpu()
    else:
        recs = local2.reshape(1, 2).cpu()
    best_c, best_i = float("inf"), -1
    for c, i in recs.tolist():
        if i < 0:
            continue
        if c < best_c or (c == best_c and i < best_i):
            best_c, best_i = c, int(i)
    return best_c, best_i


def gather_conflict(local5, group=None):
    """Sharded validatePlan: every rank scans its own range of steps (cck_validate_plan_range) and contributes
    [key, a, b, first_step, run_len] (int64; key = first_step << 32 | pair rank in std::map order, < 0: no conflict in the
    range).  all-gather -> the record with the smallest key = what the reference's validatePlan reports:
    (a, b, first_step, run_len), or None when no rank found a conflict."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        parts = [torch.zeros_like(local5) for _ in range(dist.get_world_size(group))]
        dist.all_gather(parts, local5, group=group)
        recs = torch.stack(parts).cpu()
    else:
        recs = local5.reshape(1, 5).cpu()
    return pick_conflict(recs.tolist())


def pick_conflict(records):
    """The earliest conflict among per-range records [key, a, b, first_step, run_len]: smallest non-negative key."""
    best = None
    for rec in records:
        if rec[0] < 0:
            continue
        if best is None or rec[0] < best[0]:
            best = rec
    return None if best is None else (int(best[1]), int(best[2]), int(best[3]), int(best[4]))
