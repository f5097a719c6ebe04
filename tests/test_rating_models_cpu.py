"""Host logic of the rating-prediction drop-ins (BasicMF / PMF / SVD) without a GPU: the device is
stubbed to 'cpu' and the K9 entry points are replaced by the pinned oracle, so what is checked is
everything AROUND the kernel -- id mapping, the per-epoch visiting order (MT19937 shuffle), loss
assembly, the adaptive learning rate, SVD's never-stopping loop, evalRatings -- against the golden
runs of the unmodified reference (tests/golden/mf_*_filmtrust.npz).  The kernels themselves are
compared with the same oracle in tests/test_gpu_rating.py."""
import os
import random

import numpy as np
import pytest

from qrec_b200.util.config import ModelConf

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def _install_oracle_engine(monkeypatch, calls):
    import torch
    from oracle import mf_oracle as M
    from qrec_b200 import engine as E
    from qrec_b200.model.rating._pointwise import PointwiseMF

    def np_of(t):
        return None if t is None else t.numpy()          # shares memory with the CPU tensor

    def ordered(kind, P, Q, u, i, r, wu, wi, lr, reg_u, reg_i, loss, Bu=None, Bi=None, reg_b=0.0, global_mean=0.0,
                n_warps=0):
        calls.append(('ordered', len(u), n_warps))
        # the wait arrays handed to the kernel must describe this stream
        eu, ei = E.mf_order_prepare(u.numpy(), i.numpy(), P.shape[0], Q.shape[0])
        assert np.array_equal(eu, wu.numpy()) and np.array_equal(ei, wi.numpy())
        loss += M.mf_sgd_sequential(kind, np_of(P), np_of(Q), u.numpy(), i.numpy(), r.numpy(), lr, reg_u, reg_i,
                                    np_of(Bu), np_of(Bi), reg_b, global_mean)

    def batch(kind, P, Q, u, i, r, lr, reg_u, reg_i, loss, Bu=None, Bi=None, reg_b=0.0, global_mean=0.0,
              max_inflight=0):
        # worst case of the Hogwild kernel: every window of `max_inflight` entries is one stale (Jacobi) step
        calls.append(('batch', len(u), max_inflight))
        w = max_inflight if max_inflight > 0 else len(u)
        for b in range(0, len(u), w):
            sl = slice(b, b + w)
            dP, dQ, dBu, dBi, l = M.mf_sgd_jacobi(kind, np_of(P), np_of(Q), u.numpy()[sl], i.numpy()[sl],
                                                  r.numpy()[sl], lr, reg_u, reg_i, np_of(Bu), np_of(Bi), reg_b,
                                                  global_mean)
            P += torch.from_numpy(dP).to(P.dtype); Q += torch.from_numpy(dQ).to(Q.dtype)
            if kind == 2:
                Bu += torch.from_numpy(dBu).to(Bu.dtype); Bi += torch.from_numpy(dBi).to(Bi.dtype)
            loss += l

    def sumsq(x, out):
        out += float((x.double() * x.double()).sum())

    def predict(P, Q, u, i, Bu=None, Bi=None, global_mean=0.0, out=None):
        calls.append(('predict', len(u)))
        s = (P[u.long()] * Q[i.long()]).sum(1)
        return s if Bu is None else s + global_mean + Bi[i.long()] + Bu[u.long()]

    monkeypatch.setattr(PointwiseMF, '_device', lambda self: torch.device('cpu'))
    monkeypatch.setattr(E, 'mf_sgd_ordered', ordered)
    monkeypatch.setattr(E, 'mf_sgd_batch', batch)
    monkeypatch.setattr(E, 'sumsq', sumsq)
    monkeypatch.setattr(E, 'mf_predict_pairs', predict)


def _build(name, monkeypatch, tmp_path, extra=''):
    import importlib
    g = np.load(os.path.join(GOLD, 'mf_%s_filmtrust.npz' % name.lower()))
    monkeypatch.chdir(tmp_path)
    conf = ModelConf.from_string(str(g['conf']) + extra)
    train = [[u, i, r] for u, i, r in zip(g['train_users'].tolist(), g['train_items'].tolist(), g['train_rating'].tolist())]
    test = [[u, i, r] for u, i, r in zip(g['test_users'].tolist(), g['test_items'].tolist(), g['test_rating'].tolist())]
    seed = int(g['seed'])
    random.seed(seed)
    np.random.seed(seed)
    # the reference drew nothing from `random` between seeding and the first epoch's shuffle
    assert np.array_equal(np.array(random.getstate()[1], dtype=np.uint32), g['mt_state_before'])
    cls = getattr(importlib.import_module('qrec_b200.model.rating.' + name), name)
    return g, cls(conf, train, test)


@pytest.mark.parametrize('name', ['BasicMF', 'PMF', 'SVD'])
def test_parity_mode_life_cycle_reproduces_reference_run(name, monkeypatch, tmp_path, capsys):
    calls = []
    _install_oracle_engine(monkeypatch, calls)
    g, model = _build(name, monkeypatch, tmp_path)
    losses, lrs = [], []
    orig = type(model).isConverged

    def spy(self, epoch):
        losses.append(self.loss)
        before = self.lRate
        out = orig(self, epoch)
        lrs.append((before, self.lRate))
        return out
    monkeypatch.setattr(type(model), 'isConverged', spy)
    measure = model.execute()
    assert np.array_equal(model.P, g['P_last']) and np.array_equal(model.Q, g['Q_last'])
    if name == 'SVD':
        assert np.array_equal(model.Bu, g['Bu_last']) and np.array_equal(model.Bi, g['Bi_last'])
    assert losses == g['loss'].tolist()
    assert np.array_equal(np.array(lrs), g['lrate'])
    assert np.array_equal(np.array(random.getstate()[1], dtype=np.uint32), g['mt_state_after_epoch'][-1])
    assert [m.strip() for m in measure] == g['measure'].tolist()
    assert [e[3] for e in model.data.testData] == g['test_pred'].tolist()
    assert [c[0] for c in calls] == ['ordered'] * 3 and all(64 <= c[2] <= 2368 for c in calls)
    out = capsys.readouterr().out
    assert '%s [1] epoch 3: loss = %.4f' % (name, g['loss'][2]) in out


def test_qrec_front_end_resolves_rating_models(monkeypatch, tmp_path):
    from qrec_b200.QRec import _model_class
    for name in ('BasicMF', 'PMF', 'SVD'):
        assert _model_class(name).__module__ == 'qrec_b200.model.rating.' + name
    assert _model_class('BPR').__module__ == 'qrec_b200.model.ranking.BPR'


@pytest.mark.parametrize('name', ['PMF', 'SVD'])
def test_fast_mode_trains_and_scores_on_device_tables(name, monkeypatch, tmp_path):
    """-mode fast: fp32 tables padded to a multiple of 4 columns, one launch per epoch whose in-flight
    window is sized so that the most frequent row is hit ~0.25/lr times in it (the stand-in applies
    each window as one Jacobi step, the worst case of the Hogwild kernel), test pairs scored from the
    resident tables; the run lands near the reference's error."""
    calls = []
    _install_oracle_engine(monkeypatch, calls)
    g, model = _build(name, monkeypatch, tmp_path, extra='engine=-mode fast\n')
    measure = model.execute()
    n = len(g['train_users'])
    launches = [c for c in calls if c[0] == 'batch']
    assert [c[1] for c in launches] == [n] * 3 and all(32 <= c[2] <= 4096 for c in launches)
    assert [c[0] for c in calls].count('predict') == 3      # one device scoring per epoch
    assert model.P.shape == g['P_last'].shape and model.P.dtype == np.float64
    rmse = float(measure[1].strip().split(':')[1])
    ref = float(str(g['measure'][1]).split(':')[1])
    assert abs(rmse - ref) < 0.05


def test_module_entry_point_runs_a_conf_file(monkeypatch, tmp_path, capsys):
    """`python -m qrec_b200 model.conf --seed S` == seeding + QRec(ModelConf(path)).execute(): the PMF
    golden run again, this time from files on disk through the loader."""
    calls = []
    _install_oracle_engine(monkeypatch, calls)
    g = np.load(os.path.join(GOLD, 'mf_pmf_filmtrust.npz'))
    monkeypatch.chdir(tmp_path)
    os.makedirs('dataset/FilmTrust')
    with open('dataset/FilmTrust/trainset.txt', 'w') as f:
        for u, i, r in zip(g['train_users'].tolist(), g['train_items'].tolist(), g['train_rating'].tolist()):
            f.write('%s %s %s\n' % (u, i, r))
    with open('dataset/FilmTrust/testset.txt', 'w') as f:
        for u, i, r in zip(g['test_users'].tolist(), g['test_items'].tolist(), g['test_rating'].tolist()):
            f.write('%s %s %s\n' % (u, i, r))
    with open('pmf.conf', 'w') as f:
        f.write(str(g['conf']))
    from qrec_b200.__main__ import main
    measure = main(['pmf.conf', '--seed', str(int(g['seed']))])
    assert [m.strip() for m in measure] == g['measure'].tolist()
    assert 'Running time:' in capsys.readouterr().out
