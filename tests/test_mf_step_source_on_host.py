"""Runs K9's arithmetic SOURCE (qrec_b200/csrc/mf_step.cuh, the functions mf_sgd_ordered_kernel and
mf_sgd_batch_kernel call) on the CPU through tests/host_shims/mf_step_host.cpp, which also reproduces the
warp's xor-shuffle reduction order, and pins it to the golden runs of the reference's BasicMF / PMF / SVD."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, 'tests', 'golden')
REG = dict(reg_u=0.01, reg_i=0.02, reg_b=0.03)


@pytest.fixture(scope='module')
def host(tmp_path_factory):
    out = str(tmp_path_factory.mktemp('shim') / 'libmf_step_host.so')
    subprocess.check_call(['g++', '-O2', '-std=c++17', '-ffp-contract=off', '-fPIC', '-shared', '-I',
                           os.path.join(ROOT, 'qrec_b200', 'csrc'),
                           os.path.join(ROOT, 'tests', 'host_shims', 'mf_step_host.cpp'), '-o', out])
    lib = C.CDLL(out)
    i32p = C.POINTER(C.c_int32)
    for name, fp, ft in (('host_mf_ordered_f64', C.POINTER(C.c_double), C.c_double),
                         ('host_mf_ordered_f32', C.POINTER(C.c_float), C.c_float)):
        fn = getattr(lib, name)
        fn.restype = C.c_double
        fn.argtypes = [C.c_int, fp, fp, C.c_int, C.c_int64, i32p, i32p, fp, ft, ft, ft, fp, fp, ft, ft]
    f32p = C.POINTER(C.c_float)
    lib.host_mf_delta_fast.argtypes = [C.c_int, f32p, f32p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, f32p, f32p]
    return lib


def _p(a, t):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def _load(name):
    g = np.load(os.path.join(GOLD, 'mf_%s_filmtrust.npz' % name.lower()))
    users = {n: k for k, n in enumerate(g['user_names'].tolist())}
    items = {n: k for k, n in enumerate(g['item_names'].tolist())}
    u0 = np.array([users[x] for x in g['train_users'].tolist()], np.int32)
    i0 = np.array([items[x] for x in g['train_items'].tolist()], np.int32)
    return g, u0, i0


@pytest.mark.parametrize('name', ['BasicMF', 'PMF', 'SVD'])
def test_device_step_source_replays_the_reference_run(host, name):
    """float64, three epochs in the reference's visiting orders.  Everything except the grouping of the
    dot product's partial sums (lane-strided + shuffle tree here, ddot in numpy) is the reference's
    arithmetic, so the tables agree to ~1e-12 -- the bound the GPU parity test (1e-8) relies on."""
    from oracle import bpr_oracle as O
    from oracle import mf_oracle as M
    g, u0, i0 = _load(name)
    kind = M.KINDS[name]
    P, Q = g['P0'].copy(), g['Q0'].copy()
    Bu = g['Bu0'].copy() if kind == 2 else None
    Bi = g['Bi0'].copy() if kind == 2 else None
    gm = float(g['global_mean'])
    lr, last = float(g['lrate'][0][0]), 0.0
    for e in range(3):
        o = g['order_epoch'][e]
        uu, ii = np.ascontiguousarray(u0[o]), np.ascontiguousarray(i0[o])
        rr = np.ascontiguousarray(g['train_rating'][o])
        sq = host.host_mf_ordered_f64(kind, _p(P, C.c_double), _p(Q, C.c_double), P.shape[1], len(uu), _p(uu, C.c_int32),
                                      _p(ii, C.c_int32), _p(rr, C.c_double), lr, REG['reg_u'], REG['reg_i'],
                                      _p(Bu, C.c_double), _p(Bi, C.c_double), REG['reg_b'], gm)
        loss = M.epoch_loss(kind, sq, P, Q, REG['reg_u'], REG['reg_i'], Bu, Bi, REG['reg_b'])
        assert abs(loss - g['loss'][e]) <= 1e-11 * g['loss'][e]
        if not abs(last - loss) < 1e-3:
            lr = O.update_learning_rate(lr, 1.0, e + 1, last, loss)
        assert lr == g['lrate'][e][1]
        last = loss
    np.testing.assert_allclose(P, g['P_last'], rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(Q, g['Q_last'], rtol=1e-10, atol=1e-13)
    if kind == 2:
        np.testing.assert_allclose(Bu, g['Bu_last'], rtol=1e-10, atol=1e-13)
        np.testing.assert_allclose(Bi, g['Bi_last'], rtol=1e-10, atol=1e-13)


@pytest.mark.parametrize('name', ['BasicMF', 'PMF', 'SVD'])
def test_device_step_source_f32_tracks_f32_oracle(host, name):
    from oracle import c_oracle
    from oracle import mf_oracle as M
    g, u0, i0 = _load(name)
    kind = M.KINDS[name]
    n = 8000
    f32 = np.float32
    P, Q = g['P0'].astype(f32), g['Q0'].astype(f32)
    Bu = g['Bu0'].astype(f32) if kind == 2 else None
    Bi = g['Bi0'].astype(f32) if kind == 2 else None
    Pr, Qr = P.copy(), Q.copy()
    Bur, Bir = (Bu.copy(), Bi.copy()) if kind == 2 else (None, None)
    gm = float(g['global_mean'])
    uu, ii, rr = u0[:n].copy(), i0[:n].copy(), g['train_rating'][:n].astype(f32)
    ref = c_oracle.mf_sgd_sequential(kind, Pr, Qr, uu, ii, rr, 0.02, REG['reg_u'], REG['reg_i'], Bur, Bir, REG['reg_b'], gm)
    got = host.host_mf_ordered_f32(kind, _p(P, C.c_float), _p(Q, C.c_float), P.shape[1], n, _p(uu, C.c_int32),
                                   _p(ii, C.c_int32), _p(rr, C.c_float), 0.02, REG['reg_u'], REG['reg_i'],
                                   _p(Bu, C.c_float), _p(Bi, C.c_float), REG['reg_b'], gm)
    np.testing.assert_allclose(P, Pr, rtol=2e-4, atol=2e-6)          # the GPU test's tolerance
    np.testing.assert_allclose(Q, Qr, rtol=2e-4, atol=2e-6)
    assert abs(got - ref) <= 1e-4 * ref


@pytest.mark.parametrize('kind', [0, 1])
def test_fast_deltas_equal_jacobi_oracle(host, kind):
    from oracle import mf_oracle as M
    rng = np.random.default_rng(kind)
    d = 24
    p = (rng.random(d) / 3).astype(np.float32); q = (rng.random(d) / 3).astype(np.float32)
    r = 3.5
    e = np.float32(r - float(p.astype(np.float64) @ q.astype(np.float64)))
    dp, dq = np.empty(d, np.float32), np.empty(d, np.float32)
    host.host_mf_delta_fast(kind, _p(p, C.c_float), _p(q, C.c_float), d, float(e), 0.01, 0.01, 0.02, _p(dp, C.c_float),
                            _p(dq, C.c_float))
    P, Q = p[None, :].copy(), q[None, :].copy()
    rdP, rdQ, _, _, _ = M.mf_sgd_jacobi(kind, P, Q, np.zeros(1, np.int32), np.zeros(1, np.int32), np.array([r]),
                                        0.01, 0.01, 0.02)
    np.testing.assert_allclose(dp, rdP[0], rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(dq, rdQ[0], rtol=1e-5, atol=1e-8)
