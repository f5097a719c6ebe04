"""K7 device-side bucketing (qrec_bucket_requests) and the -1 ("empty slot") convention of the row gather / scatter-add
kernels it feeds.  Index work: exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def torch():
    import torch
    assert torch.cuda.is_available()
    return torch


@pytest.fixture(scope='module')
def E():
    from qrec_b200 import engine
    return engine


@pytest.mark.parametrize('n,world,rows_per_rank', [(100003, 2, 50000), (70001, 8, 12500), (33, 3, 40), (5000, 1, 100000)])
def test_bucket_requests_is_a_partition_into_owner_buckets(torch, E, n, world, rows_per_rank):
    rng = np.random.default_rng(n)
    ids = rng.integers(0, world * rows_per_rank, n).astype(np.int32)
    cap = int(np.bincount(ids // rows_per_rank, minlength=world).max()) + 7
    dev = lambda a: torch.from_numpy(a).cuda()      # noqa: E731
    count = torch.empty(world, dtype=torch.int32, device='cuda')
    send = torch.empty(world * cap, dtype=torch.int32, device='cuda')
    pos = torch.empty(n, dtype=torch.int32, device='cuda')
    ovf = torch.zeros(1, dtype=torch.int32, device='cuda')
    E.bucket_requests(dev(ids), rows_per_rank, world, cap, count, send, pos, ovf)
    torch.cuda.synchronize()
    c, s, p = count.cpu().numpy(), send.cpu().numpy(), pos.cpu().numpy()
    owner = np.minimum(ids // rows_per_rank, world - 1)
    assert int(ovf.item()) == 0
    assert np.array_equal(c, np.bincount(owner, minlength=world))
    assert len(np.unique(p)) == n                                         # every request has its own slot
    assert np.array_equal(p // cap, owner)                                # ... inside its owner's bucket
    assert np.all(p % cap < c[owner])                                     # ... among the first count[owner] slots
    assert np.array_equal(s[p], ids - owner * rows_per_rank)              # the slot carries the owner-local row id
    used = np.zeros(world * cap, bool); used[p] = True
    assert np.all(s[~used] == -1)                                         # everything else is an empty slot
    # too small a capacity is reported, never written out of bounds
    small = max(1, cap // 2)
    send2 = torch.full((world * small + 16,), 12345, dtype=torch.int32, device='cuda')
    E.bucket_requests(dev(ids), rows_per_rank, world, small, count, send2[:world * small], pos, ovf)
    torch.cuda.synchronize()
    assert int(ovf.item()) == 1 and bool((send2[world * small:] == 12345).all())
    assert int(pos.max()) < world * small


def test_gather_and_scatter_skip_empty_slots(torch, E):
    T = torch.arange(40, dtype=torch.float32, device='cuda').view(10, 4)
    idx = torch.tensor([3, -1, 0, -1, 9], dtype=torch.int32, device='cuda')
    out = torch.full((5, 4), 7.0, device='cuda')
    E.gather_rows(T, idx, out)
    assert torch.equal(out[0], T[3]) and torch.equal(out[2], T[0]) and torch.equal(out[4], T[9])
    assert float(out[1].abs().sum()) == 0.0 and float(out[3].abs().sum()) == 0.0
    G = torch.zeros(10, 4, device='cuda')
    E.scatter_add_rows(G, idx, torch.ones(5, 4, device='cuda'))
    assert float(G.sum()) == 12.0 and float(G[3].sum()) == 4.0 and float(G[1].sum()) == 0.0
