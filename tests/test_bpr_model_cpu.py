"""Host logic of the BPR drop-in (model/ranking/BPR.py mirror) without a GPU: the device is stubbed to
'cpu' and the kernels are replaced by the pinned oracle, so what is checked is everything AROUND them --
the sampler hand-over with the interpreter's MT19937 state, loss assembly, the learning-rate rule, the
per-epoch shuffle, evaluation (numpy path, a1-a6, a20), and the composition of the minibatch Adam variant
(a19, BPR.trainModel_tf)."""
import contextlib
import io
import random

import numpy as np
import pytest

from qrec_b200.util.config import ModelConf


def _records(g):
    train = [[u, i, r] for u, i, r in zip(g['train_users'].tolist(), g['train_items'].tolist(), g['train_rating'].tolist())]
    test = [[u, i, r] for u, i, r in zip(g['test_users'].tolist(), g['test_items'].tolist(), g['test_rating'].tolist())]
    return train, test


def _stub_engine(monkeypatch, calls):
    import torch
    from oracle import bpr_oracle as O
    from qrec_b200 import engine as E
    from qrec_b200.base.iterativeRecommender import IterativeRecommender

    def ordered(P, Q, u, i, j, wu, wi, wj, lr, reg_u, reg_i, loss, n_warps=0):
        calls.append(('ordered', len(u), n_warps))
        eu, ei, ej = E.bpr_order_prepare(u.numpy(), i.numpy(), j.numpy(), P.shape[0], Q.shape[0])
        assert np.array_equal(eu, wu.numpy()) and np.array_equal(ei, wi.numpy()) and np.array_equal(ej, wj.numpy())
        t = np.stack([u.numpy(), i.numpy(), j.numpy()], 1)
        loss += O.bpr_sgd_sequential(P.numpy(), Q.numpy(), t, lr, reg_u, reg_i)

    def sumsq(x, out):
        out += float((x.double() * x.double()).sum())

    def axpby(dst, a, b, alpha, beta):
        dst.copy_(alpha * a + beta * b)

    def grad_scatter(U, V, u, i, j, eps, reg, gU, gV, loss):
        calls.append(('grad', len(u)))
        l, a, b = O.bpr_loss_grad(U.numpy(), V.numpy(), u.numpy(), i.numpy(), j.numpy(), eps, reg)
        gU += torch.from_numpy(a).float(); gV += torch.from_numpy(b).float()
        loss += l

    def adam(var, m, v, g, lr, t, beta1=0.9, beta2=0.999, eps=1e-8):
        O.adam_tf1(var.numpy(), m.numpy(), v.numpy(), g.numpy(), lr, t)

    monkeypatch.setattr(IterativeRecommender, '_device', lambda self: torch.device('cpu'))
    for name, fn in (('bpr_sgd_ordered', ordered), ('sumsq', sumsq), ('axpby', axpby), ('bpr_grad_scatter', grad_scatter),
                     ('adam_dense_tf1', adam)):
        monkeypatch.setattr(E, name, fn)


def test_numpy_path_life_cycle_reproduces_reference_run(golden_bpr, monkeypatch, tmp_path):
    """random.seed / np.random.seed -> execute(): same triples, P, Q, epoch losses, learning rates,
    generator states and ranking metrics as the golden run of the reference."""
    from qrec_b200.model.ranking.BPR import BPR
    g = golden_bpr
    calls = []
    _stub_engine(monkeypatch, calls)
    monkeypatch.chdir(tmp_path)
    train, test = _records(g)
    random.setstate((3, tuple(int(x) for x in g['mt_state_after_split']), None))
    np.random.seed(0)
    model = BPR(ModelConf.from_string(str(g['conf'])), train, test)
    losses, lrs, states, first = [], [], [], {}
    orig = BPR.isConverged

    def spy(self, epoch):
        losses.append(self.loss)
        before = self.lRate
        out = orig(self, epoch)
        lrs.append((before, self.lRate))
        states.append(np.array(random.getstate()[1], dtype=np.uint32))
        return out
    monkeypatch.setattr(BPR, 'isConverged', spy)
    with contextlib.redirect_stdout(io.StringIO()):
        measure = model.execute()
    assert losses == g['loss'].tolist()
    assert np.array_equal(np.array(lrs), g['lrate'])
    assert all(np.array_equal(a, b) for a, b in zip(states, g['mt_state_after_epoch']))
    assert np.array_equal(model.P.astype(np.float32), g['P_epoch3']) and np.array_equal(model.Q.astype(np.float32), g['Q_epoch3'])
    assert [m.strip() for m in measure] == g['measure'].tolist()
    assert [c[0] for c in calls] == ['ordered'] * 3 and all(64 <= c[2] <= 2368 for c in calls)


def test_trainModel_tf_composition_equals_autograd_restatement(golden_bpr, monkeypatch, tmp_path):
    """a19 (BPR.py:77-96): loss = -sum ln(sigmoid(y) + 1e-6) + regU*(l2_loss(U) + l2_loss(V)) over the FULL
    tables, TF1 Adam on both tables every minibatch.  The engine composes it from K3 (batch term), an
    axpby that seeds the gradient buffers with regU*table, and K4; here those pieces are the oracle's and
    the result is compared with float64 autograd of the stated loss."""
    import torch
    from oracle import bpr_oracle as O
    from qrec_b200.model.ranking.BPR import BPR
    g = golden_bpr
    calls = []
    _stub_engine(monkeypatch, calls)
    monkeypatch.chdir(tmp_path)
    n_tr = 3000
    train = [[u, i, r] for u, i, r in zip(g['train_users'][:n_tr].tolist(), g['train_items'][:n_tr].tolist(), [1.0] * n_tr)]
    conf = ModelConf.from_string(str(g['conf']).replace('num.max.epoch=3', 'num.max.epoch=2').replace('batch_size=2048', 'batch_size=700'))
    model = BPR(conf, train, [])
    with contextlib.redirect_stdout(io.StringIO()):
        model.readConfiguration()
        model.initModel()
    random.seed(21)
    torch.manual_seed(5)
    with contextlib.redirect_stdout(io.StringIO()):
        model.trainModel_tf()
    assert [c[1] for c in calls if c[0] == 'grad'] == [700, 700, 700, 700, 200] * 2
    # restatement: same initial tables, same (u,i,j) stream, autograd of the stated loss, TF1 Adam
    torch.manual_seed(5)
    d, nu, ni = model.emb_size, model.num_users, model.num_items
    U = torch.nn.init.trunc_normal_(torch.empty(nu, d), std=0.005, a=-0.01, b=0.01).numpy().copy()
    V = torch.nn.init.trunc_normal_(torch.empty(ni, d), std=0.005, a=-0.01, b=0.01).numpy().copy()
    mU, vU, mV, vV = (np.zeros_like(x) for x in (U, U, V, V))
    random.seed(21)
    ref = BPR(conf, train, [])
    with contextlib.redirect_stdout(io.StringIO()):
        ref.readConfiguration()
    ref.batch_size = 700
    t = 0
    for epoch in range(2):
        for u, i, j in ref.next_batch():
            t += 1
            Ut = torch.tensor(U, dtype=torch.float64, requires_grad=True)
            Vt = torch.tensor(V, dtype=torch.float64, requires_grad=True)
            ul, il, jl = (torch.from_numpy(x.astype(np.int64)) for x in (u, i, j))
            y = (Ut[ul] * Vt[il]).sum(1) - (Ut[ul] * Vt[jl]).sum(1)
            loss = -torch.log(torch.sigmoid(y) + 1e-6).sum() + ref.regU * (0.5 * (Ut ** 2).sum() + 0.5 * (Vt ** 2).sum())
            loss.backward()
            O.adam_tf1(U, mU, vU, Ut.grad.numpy().astype(np.float32), ref.lRate, t)
            O.adam_tf1(V, mV, vV, Vt.grad.numpy().astype(np.float32), ref.lRate, t)
    # Adam divides by sqrt(v): a gradient component near zero turns fp32 rounding into a visible step
    # difference, hence the absolute floor (tables are O(0.05) after ten steps of 1e-2... 1e-3)
    np.testing.assert_allclose(model.P, U, rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(model.Q, V, rtol=1e-3, atol=1e-5)
    assert float(np.abs(model.P - U).max()) < 1e-5 and float(np.abs(model.P).max()) > 0.01
