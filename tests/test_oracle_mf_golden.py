"""Pins oracle/mf_oracle.py (and the C restatement) to the golden runs of the unmodified reference's
BasicMF / PMF / SVD on FilmTrust (tests/golden/mf_*_filmtrust.npz, oracle/gen_golden.py mf)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import bpr_oracle as O            # noqa: E402
from oracle import mf_oracle as M             # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')
REG = dict(reg_u=0.01, reg_i=0.02, reg_b=0.03)


def load(name):
    g = np.load(os.path.join(GOLD, 'mf_%s_filmtrust.npz' % name.lower()))
    users = {n: k for k, n in enumerate(g['user_names'].tolist())}
    items = {n: k for k, n in enumerate(g['item_names'].tolist())}
    u0 = np.array([users[x] for x in g['train_users'].tolist()], np.int32)
    i0 = np.array([items[x] for x in g['train_items'].tolist()], np.int32)
    return g, u0, i0, users, items


def replay(name, g, u0, i0, sgd):
    kind = M.KINDS[name]
    P, Q = g['P0'].copy(), g['Q0'].copy()
    Bu = g['Bu0'].copy() if kind == M.SVD else None
    Bi = g['Bi0'].copy() if kind == M.SVD else None
    gm = float(g['global_mean'])
    lr = float(g['lrate'][0][0])
    last, losses, lrs = 0.0, [], []
    first = None
    for e in range(g['order_epoch'].shape[0]):
        o = g['order_epoch'][e]
        sq = sgd(kind, P, Q, u0[o], i0[o], g['train_rating'][o], lr, REG['reg_u'], REG['reg_i'], Bu, Bi,
                 REG['reg_b'], gm)
        loss = M.epoch_loss(kind, sq, P, Q, REG['reg_u'], REG['reg_i'], Bu, Bi, REG['reg_b'])
        losses.append(loss)
        before = lr
        if not abs(last - loss) < 1e-3:
            lr = O.update_learning_rate(lr, 1.0, e + 1, last, loss)
        lrs.append((before, lr))
        last = loss
        if e == 0:
            first = (P.copy(), Q.copy())
    return P, Q, Bu, Bi, losses, lrs, first


@pytest.mark.parametrize('name', ['BasicMF', 'PMF', 'SVD'])
def test_numpy_oracle_reproduces_reference_bits(name):
    g, u0, i0, _, _ = load(name)
    P, Q, Bu, Bi, losses, lrs, first = replay(name, g, u0, i0, M.mf_sgd_sequential)
    assert np.array_equal(P, g['P_last']) and np.array_equal(Q, g['Q_last'])
    if name == 'SVD':
        assert np.array_equal(Bu, g['Bu_last']) and np.array_equal(Bi, g['Bi_last'])
    assert losses == g['loss'].tolist()
    assert np.array_equal(np.array(lrs), g['lrate'])
    assert np.allclose(first[0], g['P_epoch1'], rtol=1e-6, atol=1e-7)
    assert np.allclose(first[1], g['Q_epoch1'], rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize('name', ['BasicMF', 'PMF', 'SVD'])
def test_epoch_order_is_the_mt19937_shuffle(name):
    """isConverged shuffles trainingData after every epoch (iterativeRecommender.py:101): the golden
    visiting orders are random.shuffle replayed from the recorded generator state."""
    g, _, _, _, _ = load(name)
    n = g['order_epoch'].shape[1]
    assert np.array_equal(g['order_epoch'][0], np.arange(n))
    rng = O.make_rng(state625=g['mt_state_before'])
    order = list(range(n))
    for e in range(g['order_epoch'].shape[0]):
        assert order == g['order_epoch'][e].tolist()
        rng.shuffle(order)
        assert np.array_equal(O.rng_state(rng), g['mt_state_after_epoch'][e])


@pytest.mark.parametrize('name', ['BasicMF', 'PMF', 'SVD'])
def test_final_measure_from_tables(name):
    """MAE / RMSE of evalRatings (base/recommender.py:95-125) recomputed from the golden tables."""
    g, _, _, users, items = load(name)
    kind = M.KINDS[name]
    gm = float(g['global_mean'])
    Bu = g['Bu_last'] if kind == M.SVD else None
    Bi = g['Bi_last'] if kind == M.SVD else None
    preds = []
    for k, (un, it) in enumerate(zip(g['test_users'].tolist(), g['test_items'].tolist())):
        if un in users and it in items:
            preds.append(M.predict_rating(kind, g['P_last'], g['Q_last'], users[un], items[it], Bu, Bi, gm))
        else:
            preds.append(g['test_pred'][k])        # cold rows: user / item / global mean (checked elsewhere)
    lo, hi = 0.5, 4.0                              # FilmTrust's rating scale; recommender.py:84-90
    preds = np.array([hi if x > hi else lo if x < lo else round(float(x), 3) for x in preds])
    warm = np.array([un in users and it in items for un, it in zip(g['test_users'].tolist(), g['test_items'].tolist())])
    assert np.array_equal(preds[warm], g['test_pred'][warm])
    err = np.abs(g['test_rating'] - g['test_pred'])
    assert 'MAE:' + str(float(err.sum()) / len(err)) == str(g['measure'][0]) or \
        abs(float(str(g['measure'][0])[4:]) - err.mean()) < 1e-12
    assert abs(float(str(g['measure'][1])[5:]) - np.sqrt((err ** 2).mean())) < 1e-12


@pytest.mark.parametrize('name', ['BasicMF', 'PMF', 'SVD'])
def test_c_restatement_tracks_reference(name):
    """oracle/mf_ref.c: same loop with a left-to-right dot product -- last-bit differences in the
    dot, amplified over 3 x 33 750 sequential steps, stay below 1e-9."""
    from oracle import c_oracle
    g, u0, i0, _, _ = load(name)
    P, Q, Bu, Bi, losses, _, _ = replay(name, g, u0, i0, c_oracle.mf_sgd_sequential)
    assert np.allclose(P, g['P_last'], rtol=1e-9, atol=1e-11) and np.allclose(Q, g['Q_last'], rtol=1e-9, atol=1e-11)
    if name == 'SVD':
        assert np.allclose(Bu, g['Bu_last'], rtol=1e-9, atol=1e-11)
    assert np.allclose(losses, g['loss'], rtol=1e-11)


def test_jacobi_equals_sequential_without_row_reuse():
    rng = np.random.default_rng(5)
    P = rng.random((40, 8)); Q = rng.random((50, 8)); Bu = rng.random(40); Bi = rng.random(50)
    u = rng.permutation(40)[:30].astype(np.int32); i = rng.permutation(50)[:30].astype(np.int32)
    r = rng.integers(1, 9, 30) / 2.0
    for kind in (M.BASIC, M.PMF, M.SVD):
        dP, dQ, dBu, dBi, l0 = M.mf_sgd_jacobi(kind, P, Q, u, i, r, 0.05, 0.01, 0.02, Bu, Bi, 0.03, 2.5)
        P1, Q1, Bu1, Bi1 = P.copy(), Q.copy(), Bu.copy(), Bi.copy()
        l1 = M.mf_sgd_sequential(kind, P1, Q1, u, i, r, 0.05, 0.01, 0.02, Bu1, Bi1, 0.03, 2.5)
        assert np.allclose(P + dP, P1, atol=1e-14) and np.allclose(Q + dQ, Q1, atol=1e-14)
        if kind == M.SVD:
            assert np.allclose(Bu + dBu, Bu1, atol=1e-14) and np.allclose(Bi + dBi, Bi1, atol=1e-14)
        assert abs(l0 - l1) < 1e-9
