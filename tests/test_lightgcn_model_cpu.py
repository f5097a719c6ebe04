"""Host logic of the LightGCN drop-in (a12-a15) without a GPU: device stubbed to 'cpu', kernels replaced
by dense / oracle stand-ins.  Checks the class's own step -- layer-mean propagation through the SpMM
epilogue, the last forward layer on the batch's rows only, the sparse first backward layer (sorted, -1-padded row lists), the dense Adam on the ego table,
the padded table width -- against the oracle's restatement of model/ranking/LightGCN.py:13-39 on the
reference's FilmTrust graph and sampled batches."""
import contextlib
import io

import numpy as np

from qrec_b200.util.config import ModelConf


def _stub(monkeypatch, calls):
    import scipy.sparse as sp
    import torch
    from oracle import bpr_oracle as O
    from qrec_b200 import engine as E
    from qrec_b200.base.iterativeRecommender import IterativeRecommender

    def dense(rowptr, cols, vals, n_cols):
        # a row range of a CSR is passed as a slice of rowptr over the whole cols / vals arrays (absolute offsets)
        rp = rowptr.numpy()
        a, b = int(rp[0]), int(rp[-1])
        return sp.csr_matrix((vals.numpy()[a:b], cols.numpy()[a:b], rp - a), shape=(rowptr.numel() - 1, n_cols))

    def spmm(rowptr, cols, vals, X, Y, acc=None, acc_scale=0.0, rowsplit=False):
        calls.append('spmm')
        Y.copy_(torch.from_numpy(dense(rowptr, cols, vals, X.shape[0]) @ X.numpy()))
        if acc is not None:
            acc.add_(Y, alpha=acc_scale)
        return Y

    def scatter_rows(rowptr, cols, vals, src_rows, X, Y, acc=None, acc_scale=0.0):
        calls.append('scatter_rows')
        keep = torch.zeros(X.shape[0], dtype=torch.bool)
        listed = src_rows[src_rows >= 0].long()            # -1 entries are padding
        assert listed.numel() == torch.unique(listed).numel()
        keep[listed] = True
        assert float(X[~keep].abs().sum()) == 0.0          # the caller's claim: only those rows are non-zero
        A = dense(rowptr, cols, vals, Y.shape[0])          # symmetric adjacency: B^T X == A X
        Y.copy_(torch.from_numpy(A.T @ X.numpy()))
        if acc is not None:
            acc.add_(Y, alpha=acc_scale)
        return Y

    def list_rows(rowptr, cols, vals, rows, X, Y=None, compact=False, acc=None, acc_scale=0.0):
        calls.append('rows')
        A = dense(rowptr, cols, vals, X.shape[0])
        listed = rows[rows >= 0].long()
        assert listed.numel() == torch.unique(listed).numel()
        part = torch.from_numpy(A[listed.numpy()] @ X.numpy())
        if Y is not None:
            if compact:
                Y.zero_()
                Y[(rows >= 0).nonzero().ravel()] = part
            else:
                Y[listed] = part
        if acc is not None:
            acc[listed] += acc_scale * part                # ONLY the listed rows receive the layer's term
        return Y

    def grad_scatter(U, V, u, i, j, eps, reg, gU, gV, loss):
        l, a, b = O.bpr_loss_grad(U.numpy(), V.numpy(), u.numpy(), i.numpy(), j.numpy(), eps, reg)
        gU += torch.from_numpy(a).float(); gV += torch.from_numpy(b).float()
        loss += l

    monkeypatch.setattr(IterativeRecommender, '_device', lambda self: torch.device('cpu'))
    from conftest import adjacency_kernel_stand_ins
    adjacency_kernel_stand_ins(monkeypatch)
    monkeypatch.setattr(E, 'spmm_csr', spmm)
    monkeypatch.setattr(E, 'spmm_csr_scatter_rows', scatter_rows)
    monkeypatch.setattr(E, 'spmm_csr_rows', list_rows)
    monkeypatch.setattr(E, 'bpr_grad_scatter', grad_scatter)
    monkeypatch.setattr(E, 'axpby', lambda dst, a, b, alpha, beta: dst.copy_(alpha * a + beta * b))
    monkeypatch.setattr(E, 'adam_dense_tf1', lambda var, m, v, g, lr, t, beta1=0.9, beta2=0.999, eps=1e-8:
                        O.adam_tf1(var.numpy(), m.numpy(), v.numpy(), g.numpy(), lr, t))


def test_lightgcn_step_equals_oracle_restatement(golden_graph, graph_ids, monkeypatch, tmp_path):
    from oracle import bpr_oracle as O
    import torch
    from qrec_b200.model.ranking.LightGCN import LightGCN
    g = golden_graph
    calls = []
    _stub(monkeypatch, calls)
    monkeypatch.chdir(tmp_path)
    train = [[u, i, 1.0] for u, i in zip(g['train_users'].tolist(), g['train_items'].tolist())]
    conf = ModelConf.from_string(str(g['conf']).replace('num.factors=64', 'num.factors=50'))      # the shipped confs use 50
    m = LightGCN(conf, train, [])
    with contextlib.redirect_stdout(io.StringIO()):
        m.readConfiguration()
        m.initModel()
    U_, I_ = m.num_users, m.num_items
    assert m.ego.shape == (U_ + I_, 52) and float(m.ego[:, 50:].abs().sum()) == 0.0
    uu, ii, _, _ = graph_ids
    adj = O.norm_adjacency(U_, I_, uu, ii)
    Ur, Vr = m.ego[:U_, :50].numpy().copy(), m.ego[U_:, :50].numpy().copy()
    mU, vU, mV, vV = (np.zeros_like(x) for x in (Ur, Ur, Vr, Vr))
    su, si, sj = g['shuffled_u'], g['shuffled_i'], g['pair_all_j']
    for step in range(3):
        sl = slice(step * 2048, (step + 1) * 2048)
        u, i, j = (np.ascontiguousarray(x[sl]) for x in (su, si, sj))
        ref_loss = O.lightgcn_step(adj, Ur, Vr, mU, vU, mV, vV, u, i, j, m.n_layers, m.lRate, m.regU, step + 1)
        loss = float(m.train_step(torch.from_numpy(u), torch.from_numpy(i), torch.from_numpy(j)).item())
        assert abs(loss - ref_loss) <= 1e-5 * abs(ref_loss)
        np.testing.assert_allclose(m.ego[:U_, :50].numpy(), Ur, rtol=2e-3, atol=2e-5)
        np.testing.assert_allclose(m.ego[U_:, :50].numpy(), Vr, rtol=2e-3, atol=2e-5)
        assert float(m.ego[:, 50:].abs().sum()) == 0.0                     # padding columns never move
    # per step: (n - 1) forward SpMMs + the last layer on the batch's rows only; one sparse-source product
    # + (n - 1) SpMMs backward; every whole-graph SpMM is two launches, one per bipartite half (DeviceCSR.set_split_row)
    n = m.n_layers
    assert m.norm_adj.split_row == U_
    assert calls == (['spmm'] * 2 * (n - 1) + ['rows', 'scatter_rows'] + ['spmm'] * 2 * (n - 1)) * 3
    Ue, Ve = m.propagate()
    fu, fv, _ = O.lightgcn_forward(adj, Ur, Vr, n)
    np.testing.assert_allclose(Ue[:, :50].numpy(), fu, rtol=2e-3, atol=2e-5)
    np.testing.assert_allclose(Ve[:, :50].numpy(), fv, rtol=2e-3, atol=2e-5)
