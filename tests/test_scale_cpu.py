"""Host logic of ScaleBPR (f-3) and the batched evaluation (f-1) without a GPU: the device is stubbed to
'cpu' and the kernels are replaced by oracle stand-ins (Philox negatives + the sequential step, a dense
product, a numpy mask), so the bookkeeping around them -- loss assembly, the reference's learning-rate /
convergence rules, id spaces, the rated-items-score-0 quirk, metric assembly -- is checked on the CPU."""
import numpy as np
import pytest

from qrec_b200.data.interactions import InteractionTable


def _stub(monkeypatch, calls):
    import torch
    from oracle import bpr_oracle as O
    from qrec_b200 import engine as E
    from qrec_b200.scale import ScaleBPR

    def epoch(P, Q, rowptr, i, rated_rowptr, rated_cols, num_items, seed, epoch, lr, reg_u, reg_i, loss, j_out=None):
        rp, rr, rc = rowptr.numpy(), rated_rowptr.numpy(), rated_cols.numpy()
        users = np.repeat(np.arange(len(rp) - 1), np.diff(rp)).astype(np.int32)
        rated = [set(rc[rr[k]:rr[k + 1]].tolist()) for k in range(len(rr) - 1)]
        j = O.sample_neg_philox(users, rated, num_items, seed, epoch)
        calls.append(('epoch', epoch, lr))
        t = np.stack([users, i.numpy(), j], 1)
        loss += O.bpr_sgd_sequential(P.numpy(), Q.numpy(), t, np.float32(lr), np.float32(reg_u), np.float32(reg_i))

    def sumsq(x, out):
        out += float((x.double() * x.double()).sum())

    def sgemm(A, B, C, trans_a=False, trans_b=False, alpha=1.0, beta=0.0):
        C.copy_(alpha * (A.t() if trans_a else A) @ (B.t() if trans_b else B))

    def mask_rated(scores, users, rowptr, cols, value=0.0):
        rp, co = rowptr.numpy(), cols.numpy()
        for row, u in enumerate(users.numpy().tolist()):
            scores[row, torch.from_numpy(co[rp[u]:rp[u + 1]].astype(np.int64))] = value

    def score_topn(U, V, user_ids, rated_rowptr, rated_cols, N, rated_value=0.0, out_ids=None, out_scores=None):
        # K8 stand-in: the reference's per-user flow (recommender.py:143-152), ties by ascending item id
        rp, co = rated_rowptr.numpy(), rated_cols.numpy()
        ids = torch.empty(len(user_ids), N, dtype=torch.int32)
        val = torch.empty(len(user_ids), N, dtype=torch.float32)
        for row, u in enumerate(user_ids.numpy().tolist()):
            sc = (V @ U[u]).numpy().copy()
            sc[co[rp[u]:rp[u + 1]]] = rated_value
            top = np.argsort(-sc, kind='stable')[:N]
            ids[row] = torch.from_numpy(top.astype(np.int32)); val[row] = torch.from_numpy(sc[top])
        return ids, val

    monkeypatch.setattr(E, 'score_topn', score_topn)
    monkeypatch.setattr(ScaleBPR, '_make_device', staticmethod(lambda index: torch.device('cpu')))
    for name, fn in (('bpr_epoch_usermajor', epoch), ('sumsq', sumsq), ('sgemm', sgemm), ('mask_rated', mask_rated)):
        monkeypatch.setattr(E, name, fn)


def test_scale_bpr_bookkeeping_and_evaluation(golden_bpr, monkeypatch):
    from oracle import bpr_oracle as O
    from qrec_b200.scale import ScaleBPR
    from qrec_b200.util.measure import Measure
    g = golden_bpr
    calls = []
    _stub(monkeypatch, calls)
    n = 6000
    table = InteractionTable.from_records([[u, i, 1.0] for u, i in zip(g['train_users'][:n].tolist(), g['train_items'][:n].tolist())])
    np.random.seed(3)
    model = ScaleBPR(table, emb_size=10, lr=0.05, reg_u=0.01, reg_i=0.01, seed=9)      # d = 10 -> padded to 12
    assert model.P.shape[1] == 12 and float(model.P[:, 10:].abs().sum()) == 0.0
    model.fit(4)
    # epoch bookkeeping: BPR.py:40,53 loss, iterativeRecommender.py:56-63 learning rate
    assert [c[1] for c in calls] == [0, 1, 2, 3]
    lr, last = 0.05, 0.0
    for (ep, loss, delta, used_lr), call in zip(model.history, calls):
        assert used_lr == lr == call[2] and delta == last - loss
        if not abs(delta) < 1e-3:
            lr = O.update_learning_rate(lr, 1.0, ep, last, loss)
        last = loss
    assert model.history[-1][1] < model.history[0][1]                      # the loss falls
    assert float(model.P[:, 10:].abs().sum()) == 0.0                        # padding columns stay zero
    P, Q = model.tables()
    assert P.shape == (table.num_users, 10)
    # batched top-N == per-user host flow of the reference (rated items scored 0, recommender.py:147-149)
    users = list(range(0, table.num_users, 7))
    ids, vals = model.top_n(users, N=5, block=16)
    rated = model.csr
    for row, u in enumerate(users):
        s = Q.astype(np.float32) @ P[u].astype(np.float32)
        s[rated.sorted_cols[rated.sorted_rowptr[u]:rated.sorted_rowptr[u + 1]]] = 0.0
        want = np.argsort(-s, kind='stable')[:5]
        assert np.allclose(np.sort(s[want])[::-1], vals[row], atol=1e-6)
        assert set(ids[row].tolist()) == set(want.tolist()) or np.allclose(s[ids[row]], s[want], atol=1e-6)
    # metrics through the reference's definitions
    test = [[u, i, 1.0] for u, i in zip(g['test_users'][:800].tolist(), g['test_items'][:800].tolist())]
    lines = model.evaluate(test, tops=(5,))
    assert lines[0].startswith('Top 5') and all(k in ''.join(lines) for k in ('Precision:', 'Recall:', 'F1:', 'NDCG:'))
    origin, res = {}, {}
    lut_u = {nme: k for k, nme in enumerate(table.user_names.tolist())}
    inames = table.item_names.tolist()
    for u, i, _ in test:
        if u in lut_u:
            origin.setdefault(u, {})[i] = 1
    ids_all, vals_all = model.top_n([lut_u[u] for u in origin], N=5)
    for (u, _), row_ids, row_vals in zip(origin.items(), ids_all, vals_all):
        res[u] = [(inames[k], float(v)) for k, v in zip(row_ids.tolist(), row_vals.tolist())]
    want = Measure.rankingMeasure(origin, res, [5])
    for a, b in zip(lines, want):
        if ':' in a:
            assert a.split(':')[0] == b.split(':')[0] and abs(float(a.split(':')[1]) - float(b.split(':')[1])) < 1e-9


def test_scale_bpr_raises_on_nan(monkeypatch, golden_bpr):
    from qrec_b200 import engine as E
    from qrec_b200.scale import ScaleBPR
    calls = []
    _stub(monkeypatch, calls)
    g = golden_bpr
    table = InteractionTable.from_records([[u, i, 1.0] for u, i in zip(g['train_users'][:200].tolist(), g['train_items'][:200].tolist())])
    model = ScaleBPR(table, emb_size=8)
    monkeypatch.setattr(E, 'bpr_epoch_usermajor', lambda *a, **k: a[12].add_(float('nan')))
    with pytest.raises(FloatingPointError):
        model.run_epoch()
