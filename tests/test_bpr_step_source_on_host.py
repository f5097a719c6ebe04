"""Runs K1's arithmetic SOURCE (qrec_b200/csrc/bpr_step.cuh: the parity update of bpr_sgd_ordered_kernel and
the 4-wide steps of the throughput kernels) on the CPU through tests/host_shims/bpr_step_host.cpp and pins
it to the golden run of the reference's BPR (tests/golden/bpr_filmtrust_seed0.npz)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LR, REG = 0.01, 0.001


@pytest.fixture(scope='module')
def host(tmp_path_factory):
    out = str(tmp_path_factory.mktemp('shim') / 'libbpr_step_host.so')
    subprocess.check_call(['g++', '-O2', '-std=c++17', '-ffp-contract=off', '-fPIC', '-shared', '-I',
                           os.path.join(ROOT, 'qrec_b200', 'csrc'),
                           os.path.join(ROOT, 'tests', 'host_shims', 'bpr_step_host.cpp'), '-o', out])
    lib = C.CDLL(out)
    i32p, f32p = C.POINTER(C.c_int32), C.POINTER(C.c_float)
    for name, fp, ft in (('host_bpr_ordered_f64', C.POINTER(C.c_double), C.c_double), ('host_bpr_ordered_f32', f32p, C.c_float)):
        fn = getattr(lib, name)
        fn.restype = C.c_double
        fn.argtypes = [fp, fp, C.c_int, C.c_int64, i32p, i32p, i32p, ft, ft, ft]
    lib.host_bpr_step4.argtypes = [f32p, f32p, f32p, C.c_float, C.c_float, C.c_float, f32p, f32p, f32p]
    lib.host_bpr_step4_inplace.argtypes = [f32p, f32p, f32p, C.c_float, C.c_float, C.c_float, f32p, f32p]
    return lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def test_parity_update_source_replays_the_reference_run(host, golden_bpr, bpr_ids):
    """float64, the reference's three epochs of (u,i,j), its loss (BPR.py:40,53) and learning-rate rule:
    the device update source reproduces P, Q after epoch 1 to 1e-10 (the bound of the GPU parity test)."""
    from oracle import bpr_oracle as O
    g = golden_bpr
    _, _, nu, ni = bpr_ids
    np.random.seed(0)
    P = np.random.rand(nu, 64) / 3
    Q = np.random.rand(ni, 64) / 3
    lr, last = LR, 0.0
    for ep in range(3):
        t = g['triples_epoch'][ep]
        u, i, j = (np.ascontiguousarray(t[:, k]) for k in range(3))
        loss = host.host_bpr_ordered_f64(_p(P, C.c_double), _p(Q, C.c_double), 64, len(u), _p(u, C.c_int32), _p(i, C.c_int32),
                                         _p(j, C.c_int32), lr, REG, REG)
        loss += O.epoch_loss_reg(P, Q, REG, REG)
        assert abs(loss - g['loss'][ep]) <= 1e-9 * g['loss'][ep]
        if ep == 0:
            np.testing.assert_allclose(P, g['P_epoch1'], rtol=1e-10, atol=1e-13)
            np.testing.assert_allclose(Q, g['Q_epoch1'], rtol=1e-10, atol=1e-13)
        assert lr == g['lrate'][ep][0]
        if not abs(last - loss) < 1e-3:
            lr = O.update_learning_rate(lr, 1.0, ep + 1, last, loss)
        last = loss
    np.testing.assert_allclose(P.astype(np.float32), g['P_epoch3'], rtol=1e-5, atol=1e-7)


def test_parity_update_source_f32_tracks_f32_oracle(host, golden_bpr, bpr_ids):
    from oracle import c_oracle
    _, _, nu, ni = bpr_ids
    np.random.seed(0)
    P = (np.random.rand(nu, 64) / 3).astype(np.float32); Q = (np.random.rand(ni, 64) / 3).astype(np.float32)
    Pr, Qr = P.copy(), Q.copy()
    t = golden_bpr['triples_epoch'][0][:10000]
    u, i, j = (np.ascontiguousarray(t[:, k]) for k in range(3))
    ref = c_oracle.bpr_sgd_sequential(Pr, Qr, u, i, j, LR, REG, REG)
    got = host.host_bpr_ordered_f32(_p(P, C.c_float), _p(Q, C.c_float), 64, len(u), _p(u, C.c_int32), _p(i, C.c_int32),
                                    _p(j, C.c_int32), LR, REG, REG)
    np.testing.assert_allclose(P, Pr, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(Q, Qr, rtol=1e-5, atol=1e-6)
    assert abs(got - ref) <= 1e-5 * ref


def test_throughput_step_sources_equal_the_reference_step(host):
    """bpr_step4 (deltas of a private copy) and bpr_step4_inplace (P in registers, decay folded into the
    coefficients) against BPR.optimization applied to one triple in float64."""
    from oracle import bpr_oracle as O
    rng = np.random.default_rng(4)
    for _ in range(200):
        p, qi, qj = (rng.random(4).astype(np.float32) for _ in range(3))
        x = float(p.astype(np.float64) @ (qi.astype(np.float64) - qj.astype(np.float64)))
        g = np.float32(LR * (1.0 - 1.0 / (1.0 + np.exp(-x))))
        P = p[None, :].astype(np.float64); Q = np.stack([qi, qj]).astype(np.float64)
        # one reference step with the same g (4 of the 64 columns: the dot product is an input here)
        pn = P[0] + g * (Q[0] - Q[1]); qin = Q[0] + g * pn; qjn = Q[1] - g * pn
        pn = pn - LR * REG * pn; qin = qin - LR * REG * qin; qjn = qjn - LR * REG * qjn
        dp, dqi, dqj = (np.empty(4, np.float32) for _ in range(3))
        host.host_bpr_step4(_p(p, C.c_float), _p(qi, C.c_float), _p(qj, C.c_float), float(g), LR * REG, LR * REG,
                            _p(dp, C.c_float), _p(dqi, C.c_float), _p(dqj, C.c_float))
        np.testing.assert_allclose(p + dp, pn, rtol=3e-7, atol=1e-8)
        np.testing.assert_allclose(qi + dqi, qin, rtol=3e-7, atol=1e-8)
        np.testing.assert_allclose(qj + dqj, qjn, rtol=3e-7, atol=1e-8)
        p2 = p.copy()
        host.host_bpr_step4_inplace(_p(p2, C.c_float), _p(qi, C.c_float), _p(qj, C.c_float), float(g), LR * REG, LR * REG,
                                    _p(dqi, C.c_float), _p(dqj, C.c_float))
        np.testing.assert_allclose(p2, pn, rtol=3e-7, atol=1e-8)
        np.testing.assert_allclose(qi + dqi, qin, rtol=3e-7, atol=1e-8)
        np.testing.assert_allclose(qj + dqj, qjn, rtol=3e-7, atol=1e-8)
