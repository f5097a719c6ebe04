"""N>1 host logic on CPU: world_size-2 gloo process group (torch.multiprocessing spawn)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from qrec_b200 import parallel


def test_user_range_partitions_exactly():
    for world in (1, 2, 3, 8):
        for n in (0, 1, 7, 8, 1000, 1_000_003):
            spans = [parallel.user_range(r, world, n) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    assert parallel.sync_points(10, 4) == [0, 2, 5, 7, 10] and parallel.sync_points(5, 0) == [0, 5]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        Q0 = torch.randn(50, 8)
        Q = Q0.clone()
        sync = parallel.ReplicatedTableSync(Q)
        # sharding: every triple lands on exactly one rank, with local user ids
        g = torch.Generator().manual_seed(1)
        u = torch.randint(0, 11, (200,), generator=g, dtype=torch.int32)
        i = torch.randint(0, 50, (200,), generator=g, dtype=torch.int32)
        j = torch.randint(0, 50, (200,), generator=g, dtype=torch.int32)
        lu, li, lj = parallel.shard_triples_by_user(u, i, j, rank, world, 11)
        lo, hi = parallel.user_range(rank, world, 11)
        assert int(lu.min()) >= 0 and int(lu.max()) < hi - lo
        cnt = torch.tensor([lu.numel()])
        dist.all_reduce(cnt)
        assert int(cnt) == 200
        # two rounds of "local training" (rank-specific scatter-adds) + delta sync
        expect = Q0.clone()
        for rnd in range(2):
            for r in range(world):
                gg = torch.Generator().manual_seed(100 * rnd + r)
                rows = torch.randint(0, 50, (30,), generator=gg)
                upd = torch.randn(30, 8, generator=gg)
                expect.index_add_(0, rows, upd)
                if r == rank:
                    Q.index_add_(0, rows, upd)
            sync.sync()
            assert torch.allclose(Q, expect, atol=1e-5), 'rank %d round %d' % (rank, rnd)
            assert torch.equal(Q, sync.base)
        out[rank] = 1
    finally:
        dist.destroy_process_group()


def test_replicated_table_delta_sync_world2():
    world = 2
    port = 29500 + os.getpid() % 2000
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert dict(out) == {0: 1, 1: 1}


def test_single_process_sync_is_identity():
    Q = torch.ones(4, 4)
    s = parallel.ReplicatedTableSync(Q)
    Q += 1
    assert s.sync() is Q and bool((Q == 2).all()) and s.world == 1
