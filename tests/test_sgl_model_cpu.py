"""Host logic and algebra of the SGL drop-in (f-4; model/ranking/SGL.py mirror) without a GPU: the device is stubbed
to 'cpu' and every kernel is replaced by a numpy / torch restatement of its documented contract, so what is checked
is the class's own composition -- per-epoch view construction through JointAdjacency, the three encoders, the merged
InfoNCE, the per-view Horner backward with per-layer matrices (aug_type 2), Adam -- against the float64 AUTOGRAD
restatement of the reference's TF graph (oracle/tf_models.sgl_loss_and_grad)."""
import contextlib
import io

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import adjacency_kernel_stand_ins
from qrec_b200.util.config import ModelConf


def _stub(monkeypatch):
    import torch
    from oracle import bpr_oracle as O
    from qrec_b200 import engine as E
    from qrec_b200.base.iterativeRecommender import IterativeRecommender
    adjacency_kernel_stand_ins(monkeypatch)

    def spmm(rowptr, cols, vals, X, Y, acc=None, acc_scale=0.0, rowsplit=False):
        rp = rowptr.numpy()                     # possibly a row range: absolute offsets into cols / vals
        a, b = int(rp[0]), int(rp[-1])
        A = sp.csr_matrix((vals.numpy()[a:b], cols.numpy()[a:b], rp - a), shape=(rowptr.numel() - 1, X.shape[0]))
        Y.copy_(torch.from_numpy(A @ X.numpy()))
        if acc is not None:
            acc.add_(Y, alpha=acc_scale)
        return Y

    def keep_philox(n_lines, drop, seed, tag, epoch, device, out=None):
        rng = np.random.default_rng([seed & 0xffffffff, tag, epoch])
        return torch.from_numpy((rng.random(n_lines) >= drop).astype(np.uint8))

    def subgraph(rowptr, cols, pair, pair_w):
        rp, co = rowptr.numpy(), cols.numpy()
        w = pair_w.numpy()[pair.numpy()]
        row = np.repeat(np.arange(len(rp) - 1), np.diff(rp))
        deg = np.zeros(len(rp) - 1, np.float32)
        np.add.at(deg, row, w)
        with np.errstate(divide='ignore'):
            dinv = np.where(deg > 0, 1.0 / np.sqrt(deg.astype(np.float64)), 0.0).astype(np.float32)
        k = w > 0
        new_rp = np.zeros(len(rp), np.int64)
        np.add.at(new_rp, row[k] + 1, 1)
        return (torch.from_numpy(np.cumsum(new_rp)), torch.from_numpy(co[k].copy()),
                torch.from_numpy(((dinv[row] * w) * dinv[co])[k].astype(np.float32)))

    def grad_scatter(U, V, u, i, j, eps, reg, gU, gV, loss):
        l, a, b = O.bpr_loss_grad(U.numpy(), V.numpy(), u.numpy(), i.numpy(), j.numpy(), eps, reg)
        gU += torch.from_numpy(a).float(); gV += torch.from_numpy(b).float()
        loss += l

    def gather_normalize(T, idx, Z, norms):
        rows = T[idx.long()]
        nrm = rows.norm(dim=1).clamp(min=1e-6)
        Z.copy_(rows / nrm[:, None]); norms.copy_(nrm)

    def sgemm(A, B, C, trans_a=False, trans_b=False, alpha=1.0, beta=0.0):
        prod = alpha * ((A.t() if trans_a else A) @ (B.t() if trans_b else B))
        C.copy_(prod if beta == 0.0 else prod + beta * C)             # beta == 0: C may be uninitialised

    def infonce_rows(S, tau, loss):                       # loss += sum_r (logsumexp(S_r/tau) - S_rr/tau); S <- dLoss/dS
        L = S.double() / tau
        loss += float((torch.logsumexp(L, 1) - L.diag()).sum())
        S.copy_(((torch.softmax(L, 1) - torch.eye(S.shape[0], dtype=torch.float64)) / tau).float())

    def normalize_bwd_scatter(dZ, Z, norms, idx, scale, G):
        g = (dZ - (dZ * Z).sum(1, keepdim=True) * Z) / norms[:, None]
        G.index_add_(0, idx.long(), scale * g)

    monkeypatch.setattr(IterativeRecommender, '_device', lambda self: torch.device('cpu'))
    for name, fn in (('spmm_csr', spmm), ('edge_keep_philox', keep_philox), ('adj_subgraph', subgraph), ('bpr_grad_scatter', grad_scatter),
                     ('gather_normalize', gather_normalize), ('sgemm', sgemm), ('infonce_rows', infonce_rows),
                     ('normalize_bwd_scatter', normalize_bwd_scatter),
                     ('axpby', lambda dst, a, b, alpha, beta: dst.copy_(alpha * a + beta * b)),
                     ('adam_dense_tf1', lambda var, m, v, g, lr, t, beta1=0.9, beta2=0.999, eps=1e-8:
                      O.adam_tf1(var.numpy(), m.numpy(), v.numpy(), g.numpy(), lr, t))):
        monkeypatch.setattr(E, name, fn)


@pytest.mark.parametrize('aug', [1, 2, 0])
def test_sgl_step_equals_autograd_restatement(golden_graph, monkeypatch, tmp_path, aug):
    import torch
    from oracle import tf_models
    from qrec_b200.model.ranking.SGL import SGL
    g = golden_graph
    _stub(monkeypatch)
    from conftest import row_list_kernel_stand_ins
    calls = row_list_kernel_stand_ins(monkeypatch)
    monkeypatch.chdir(tmp_path)
    n_tr = 6000
    train = [[u, i, 1.0] for u, i in zip(g['train_users'][:n_tr].tolist(), g['train_items'][:n_tr].tolist())]
    conf = ModelConf.from_string(str(g['conf']).replace('model.name=LightGCN', 'model.name=SGL').replace('num.factors=64', 'num.factors=10')
                                 + 'SGL=-n_layer 3 -lambda 0.1 -droprate 0.3 -augtype %d -temp 0.2\n' % aug)
    m = SGL(conf, train, [])
    with contextlib.redirect_stdout(io.StringIO()):
        m.readConfiguration()
        m.initModel()
    N, d = m.num_users + m.num_items, m.emb_size
    assert m.ego.shape == (N, 12) and float(m.ego[:, 10:].abs().sum()) == 0.0          # 10 -> padded to 12
    m.ego.mul_(20.0)
    views = m.build_views(4)
    to_sp = lambda a: sp.csr_matrix((a.vals.numpy().astype(np.float64), a.cols.numpy(), a.rowptr.numpy()), shape=(N, N))  # noqa: E731
    sp_views = [[to_sp(a) for a in v] for v in views]
    full = to_sp(m.norm_adj)
    for v in sp_views:
        for a in v:
            assert a.nnz < full.nnz and abs(a - a.T).max() < 1e-7                       # sub-graphs stay symmetric
            deg = np.asarray((a != 0).sum(1)).ravel()
            assert np.all(np.abs(np.asarray(a.multiply(a).sum(1)).ravel()[deg > 0]) > 0)
    assert (views[0][0] is views[0][1]) == (aug != 2)
    ego0 = m.ego[:, :d].numpy().astype(np.float64).copy()
    u_all, i_all, _ = m.data.training_ids()
    rng = np.random.default_rng(1)
    pick = rng.choice(len(u_all), 512, replace=False)
    u, i = u_all[pick].astype(np.int32), i_all[pick].astype(np.int32)
    j = rng.integers(0, m.num_items, 512).astype(np.int32)
    m.train_step(torch.from_numpy(u), torch.from_numpy(i), torch.from_numpy(j))
    rec, ssl = m.losses()
    rrec, rssl, rgrad = tf_models.sgl_loss_and_grad(full, sp_views, ego0, m.num_users, u, i, j, 3, 0.1, 0.2, m.regU)
    assert abs(rec - rrec) <= 1e-5 * abs(rrec) and abs(ssl - rssl) <= 1e-5 * abs(rssl)
    got = m._total[:, :d].numpy()
    assert np.abs(got - rgrad).max() <= 1e-4 * np.abs(rgrad).max()
    # the last layer of each of the three encoders ran on the batch's rows, the innermost backward product of each as a scatter
    assert calls == ['rows'] * 3 + ['scatter_rows'] * 3, calls
    assert float(m._total[:, d:].abs().sum()) == 0.0 and float(m.ego[:, d:].abs().sum()) == 0.0   # padding columns never move
