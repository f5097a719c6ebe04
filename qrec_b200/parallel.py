"""Multi-GPU plumbing for the BPR path: one process per GPU,
torch.distributed (NCCL over NVLink on the box, gloo in the CPU tests).

Sharding of throughput-mode BPR:
  * users are range-partitioned: rank r owns users [lo, hi) -- their P rows and all of their
    triples live on r, so P needs no communication at all;
  * Q (25.6 MB at the benchmark scale) is replicated.  Every rank trains on its replica; at a sync
    point the per-rank deltas are summed:  Q <- Q_base + sum_r (Q_r - Q_base).  With scatter-add
    SGD this is exactly what one GPU would have accumulated had all ranks' triples read the same
    Q_base, i.e. the data-parallel reading of the same step.
Parity mode is a serial dependency chain and does not shard (replicas only).
"""
import torch
import torch.distributed as dist


def user_range(rank, world, num_users):
    """Contiguous, balanced user range of `rank`: sizes differ by at most one."""
    base, rem = divmod(num_users, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_triples_by_user(u, i, j, rank, world, num_users):
    """Keeps the triples whose user belongs to `rank`, with user ids made local."""
    lo, hi = user_range(rank, world, num_users)
    keep = (u >= lo) & (u < hi)
    return (u[keep] - lo).contiguous(), i[keep].contiguous(), (j[keep].contiguous() if j is not None else None)


class ReplicatedTableSync(object):
    """Delta all-reduce of a replicated table (the item table Q)."""

    def __init__(self, table, group=None):
        self.table = table
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.base = table.clone() if self.world > 1 else None
        self._delta = torch.empty_like(table) if self.world > 1 else None
        self.syncs = 0

    def sync(self):
        """table <- base + sum over ranks of (table - base); base <- table."""
        if self.world == 1:
            return self.table
        torch.sub(self.table, self.base, out=self._delta)
        dist.all_reduce(self._delta, op=dist.ReduceOp.SUM, group=self.group)
        self.base.add_(self._delta)
        self.table.copy_(self.base)
        self.syncs += 1
        return self.table


def sync_points(n, pieces):
    """Boundaries that cut n triples into `pieces` nearly equal launches: [0, ..., n]."""
    pieces = max(1, int(pieces))
    return [n * s // pieces for s in range(pieces + 1)]
