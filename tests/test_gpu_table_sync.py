"""csrc/table_sync.cu on one GPU: the delta / merge kernels and the peer-memory exchange kernels (the
"peers" are buffers on the same device -- the kernels only see pointers), against torch arithmetic.
Bit-exact: every element is one subtraction / one addition in a fixed order."""
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


@pytest.mark.parametrize('rows', [1, 7, 100_000])
def test_delta_then_merge_keeps_late_updates(torch, E, rows):
    g = torch.Generator(device='cuda').manual_seed(rows)
    B = torch.randn(rows, 64, device='cuda', generator=g)
    local = torch.randn(rows, 64, device='cuda', generator=g) * 1e-2
    others = torch.randn(rows, 64, device='cuda', generator=g) * 1e-2
    late = torch.randn(rows, 64, device='cuda', generator=g) * 1e-2
    Q = B + local
    D, S = torch.empty(rows * 64, device='cuda'), torch.empty(rows * 64, device='cuda')
    E.table_delta(Q.view(-1), B.view(-1), D, S)
    assert torch.equal(D, (Q - B).view(-1)) and torch.equal(S, D)
    S += others.view(-1)                       # what the all-reduce would add
    Q += late                                  # the next K1 wave is already scattering into Q
    Q_before, B_before = Q.clone(), B.clone()
    E.table_merge(Q.view(-1), B.view(-1), D, S)
    torch.cuda.synchronize()
    assert torch.equal(Q.view(-1), Q_before.view(-1) + (S - D))
    assert torch.equal(B.view(-1), B_before.view(-1) + S)
    # invariant: Q - B == the late (not yet exchanged) updates, up to fp32 rounding of the sums
    assert float(((Q - B) - late).abs().max()) <= 1e-6


@pytest.mark.parametrize('world', [1, 2, 3, 8])
def test_peer_exchange_kernels_equal_allreduce(torch, E, world):
    rows = 1003                                  # 64192 floats: slices do not divide evenly
    n = rows * 64
    g = torch.Generator(device='cuda').manual_seed(world)
    Ds = [torch.randn(n, device='cuda', generator=g) for _ in range(world)]
    Ss = [torch.full((n,), float('nan'), device='cuda') for _ in range(world)]
    Bs = [torch.randn(n, device='cuda', generator=g) for _ in range(world)]
    Qs = [Bs[r] + Ds[r] for r in range(world)]
    for r in range(world):                      # reduce-scatter phase on every "rank"
        E.table_reduce_scatter_p2p([d.data_ptr() for d in Ds], r, Ss[r], n)
    torch.cuda.synchronize()
    n4 = n // 4
    slice4 = -(-n4 // world)
    for r in range(world):
        lo, hi = min(n, 4 * slice4 * r), min(n, 4 * slice4 * (r + 1))
        ref = torch.zeros(hi - lo, device='cuda')
        for q in range(world):
            ref = ref + Ds[q][lo:hi]
        assert torch.equal(Ss[r][lo:hi], ref), 'slice of rank %d' % r
    for r in range(world):                      # all-gather + merge phase
        Qb, Bb = Qs[r].clone(), Bs[r].clone()
        E.table_gather_merge_p2p([s.data_ptr() for s in Ss], Qs[r], Bs[r], Ds[r])
        torch.cuda.synchronize()
        full = torch.cat([Ss[min(world - 1, k)][min(n, 4 * slice4 * k):min(n, 4 * slice4 * (k + 1))] for k in range(world)])
        assert torch.equal(Bs[r], Bb + full)
        assert torch.equal(Qs[r], Qb + (full - Ds[r]))
    for r in range(world):                      # nothing was updated in between: Q - B is only rounding residue
        assert float((Qs[r] - Bs[r]).abs().max()) <= 1e-5


def test_bad_arguments(torch, E):
    x = torch.zeros(6, device='cuda')
    with pytest.raises(E.QRecError):
        E.table_delta(x, x, x)                  # not a multiple of 4
    with pytest.raises(E.QRecError):
        E.table_merge(torch.zeros(8), x, x, x)  # CPU tensor
