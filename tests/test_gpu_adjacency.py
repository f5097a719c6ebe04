"""f-2 kernels (csrc/adj_kernels.cu) on the device: the full normalised adjacency against the matrix recorded from
the reference (base/graphRecommender.py:10-29), and the per-epoch edge-dropout rebuild against a scipy restatement of
SGL._create_adj_mat (model/ranking/SGL.py:113-155, aug_type 1) fed with the same kept lines.  Structure (rowptr,
cols) must be EXACT; values within 2 ulp of fp32."""
import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def torch():
    import torch
    assert torch.cuda.is_available()
    return torch


def _sgl_adj(u, i, keep, nu, ni):
    """SGL.py:135-155: kept lines -> csr_matrix (duplicates summed) -> A + A^T -> D^-1/2 A D^-1/2 (float32)."""
    n = nu + ni
    uk, ik = u[keep], i[keep]
    tmp = sp.csr_matrix((np.ones(len(uk), np.float32), (uk, ik + nu)), shape=(n, n))
    adj = tmp + tmp.T
    rowsum = np.array(adj.sum(1))
    with np.errstate(divide='ignore'):
        d_inv = np.power(rowsum, -0.5).flatten()
    d_inv[np.isinf(d_inv)] = 0.
    out = sp.diags(d_inv).dot(adj).dot(sp.diags(d_inv)).tocsr()
    out.sort_indices()
    out.eliminate_zeros()
    return out


def test_full_adjacency_equals_reference_matrix(torch, golden_graph, graph_ids):
    from qrec_b200.graph_build import norm_adjacency_csr
    g = golden_graph
    u, i, nu, ni = graph_ids
    rowptr, cols, vals = norm_adjacency_csr(torch.from_numpy(u), torch.from_numpy(i), nu, ni, device='cuda')
    assert np.array_equal(rowptr.cpu().numpy(), g['adj_indptr'])
    assert np.array_equal(cols.cpu().numpy(), g['adj_indices'])
    np.testing.assert_allclose(vals.cpu().numpy(), g['adj_data'], rtol=5e-7, atol=0)


@pytest.mark.parametrize('drop', [0.0, 0.1, 0.5, 0.97])
def test_edge_dropout_rebuild_equals_sgl_restatement(torch, golden_graph, graph_ids, drop):
    from qrec_b200 import engine as E
    from qrec_b200.graph_build import JointAdjacency
    u, i, nu, ni = graph_ids
    J = JointAdjacency(torch.from_numpy(u), torch.from_numpy(i), nu, ni, device='cuda')
    keep = E.edge_keep_philox(len(u), drop, 77, 1, 3, 'cuda')
    kh = keep.cpu().numpy().astype(bool)
    assert abs(kh.mean() - (1 - drop)) < 0.02
    rp, co, va = J.edge_dropout(drop, 77, 1, 3)
    rp2, co2, va2 = J.edge_dropout(drop, 77, 1, 3, keep=keep)                 # caller-supplied mask: same result
    assert torch.equal(rp, rp2) and torch.equal(co, co2) and torch.equal(va, va2)
    ref = _sgl_adj(u.astype(np.int64), i.astype(np.int64), kh, nu, ni)
    assert np.array_equal(rp.cpu().numpy(), ref.indptr)
    assert np.array_equal(co.cpu().numpy(), ref.indices)
    np.testing.assert_allclose(va.cpu().numpy(), ref.data, rtol=5e-7, atol=0)
    # another view / epoch draws another sub-graph; the full graph is untouched
    rp3, _, _ = J.edge_dropout(drop, 77, 2, 3)
    if 0.0 < drop < 0.9:
        assert not torch.equal(rp3, rp)
    frp, fco, fva = J.full()
    assert int(frp[-1]) == int(fco.shape[0]) and int(frp[-1]) >= int(rp[-1])


def test_large_random_graph_scan_and_isolated_rows(torch):
    """> 1024 rows (multi-block scan), rows longer than a warp, empty rows, duplicate lines."""
    from qrec_b200.graph_build import JointAdjacency
    rng = np.random.default_rng(4)
    nu, ni, n = 5000, 700, 60000
    u = rng.integers(0, nu, n); i = (rng.random(n) ** 3 * ni).astype(np.int64)      # skewed items: long rows
    u[u % 17 == 0] = 1                                                               # a hot user, many empty users
    J = JointAdjacency(torch.from_numpy(u), torch.from_numpy(i), nu, ni, device='cuda')
    keep = (rng.random(n) > 0.3)
    rp, co, va = J.edge_dropout(0.3, 0, 0, 0, keep=torch.from_numpy(keep.astype(np.uint8)).cuda())
    ref = _sgl_adj(u, i, keep, nu, ni)
    assert np.array_equal(rp.cpu().numpy(), ref.indptr) and np.array_equal(co.cpu().numpy(), ref.indices)
    np.testing.assert_allclose(va.cpu().numpy(), ref.data, rtol=5e-7, atol=0)
