"""SBPR on the GPU: the score-scaled K3 entry point against a float64 restatement, and the drop-in's minibatch Adam
path (SBPR.py:103-134) against float64 autograd of the stated loss + the oracle's TF1 Adam on the same batches."""
import contextlib
import io
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _restate(U, V, u, i, j, c, eps):
    U, V, c = U.astype(np.float64), V.astype(np.float64), c.astype(np.float64)
    y = c * (U[u] * (V[i] - V[j])).sum(1)
    s = 1.0 / (1.0 + np.exp(-y))
    gy = (-s * (1.0 - s) / (s + eps) * c)[:, None]
    gU, gV = np.zeros_like(U), np.zeros_like(V)
    np.add.at(gU, u, gy * (V[i] - V[j]))
    np.add.at(gV, i, gy * U[u])
    np.add.at(gV, j, -gy * U[u])
    return float(-np.log(s + eps).sum()), gU, gV


@pytest.mark.parametrize('d', [12, 52, 64, 160])
def test_grad_scatter_scaled_vs_restatement(d):
    import torch
    from qrec_b200 import engine as E
    rng = np.random.default_rng(d)
    nu, ni, n = 700, 900, 5003
    U = (rng.standard_normal((nu, d)) * 0.3).astype(np.float32)
    V = (rng.standard_normal((ni, d)) * 0.3).astype(np.float32)
    u, i, j = (rng.integers(0, hi, n).astype(np.int32) for hi in (nu, ni, ni))
    c = (1.0 / (rng.integers(0, 6, n) + 1.0)).astype(np.float32)
    dev = lambda a: torch.from_numpy(a).cuda()                                  # noqa: E731
    gU, gV = torch.zeros(nu, d, device='cuda'), torch.zeros(ni, d, device='cuda')
    loss = torch.zeros(1, dtype=torch.float64, device='cuda')
    E.bpr_grad_scatter_scaled(dev(U), dev(V), dev(u), dev(i), dev(j), dev(c), 1e-6, 0.0, gU, gV, loss)
    l, a, b = _restate(U, V, u, i, j, c, 1e-6)
    assert abs(loss.item() - l) <= 1e-5 * abs(l)
    for got, ref in ((gU, a), (gV, b)):
        assert float(np.abs(got.cpu().numpy() - ref).max()) <= 2e-5 * float(np.abs(ref).max())
    # scale 1 everywhere is the unscaled entry point (same arithmetic; float atomics: order of the sums differs)
    g1, h1, l1 = torch.zeros_like(gU), torch.zeros_like(gV), torch.zeros_like(loss)
    g2, h2, l2 = torch.zeros_like(gU), torch.zeros_like(gV), torch.zeros_like(loss)
    E.bpr_grad_scatter_scaled(dev(U), dev(V), dev(u), dev(i), dev(j), torch.ones(n, device='cuda'), 1e-6, 0.01, g1, h1, l1)
    E.bpr_grad_scatter(dev(U), dev(V), dev(u), dev(i), dev(j), 1e-6, 0.01, g2, h2, l2)
    assert abs(l1.item() - l2.item()) <= 1e-6 * abs(l2.item())
    assert float((g1 - g2).abs().max()) <= 1e-5 * float(g2.abs().max()) and float((h1 - h2).abs().max()) <= 1e-5 * float(h2.abs().max())


def test_sbpr_dropin_tf_path_equals_autograd_restatement(golden_bpr, tmp_path, monkeypatch):
    import torch
    from oracle import bpr_oracle as O, tf_models as T
    from test_sbpr_cpu import _model
    monkeypatch.chdir(tmp_path)
    m, train, test, rel = _model(golden_bpr)
    random.seed(21); torch.manual_seed(5)
    with contextlib.redirect_stdout(io.StringIO()):
        m.trainModel_tf()
    torch.manual_seed(5)
    d, nu, ni = m.emb_size, m.num_users, m.num_items
    dev = torch.device('cuda', m.engine_device)
    U = torch.nn.init.trunc_normal_(torch.empty(nu, d, device=dev), std=0.005, a=-0.01, b=0.01).cpu().numpy().copy()
    V = torch.nn.init.trunc_normal_(torch.empty(ni, d, device=dev), std=0.005, a=-0.01, b=0.01).cpu().numpy().copy()
    mU, vU, mV, vV = (np.zeros_like(x) for x in (U, U, V, V))
    random.seed(21)
    t = 0
    for epoch in range(2):
        for u, i, k, j, w in m.next_batch():
            t += 1
            _, gU, gV = T.sbpr_loss_and_grad(U, V, u, i, k, j, w)           # oracle/tf_models.py (SBPR.py:103-115)
            O.adam_tf1(U, mU, vU, gU.astype(np.float32), m.lRate, t)
            O.adam_tf1(V, mV, vV, gV.astype(np.float32), m.lRate, t)
    # Adam's first steps turn a tiny gradient difference into a visible one wherever |g| ~ eps: absolute floor
    np.testing.assert_allclose(m.P, U, rtol=2e-3, atol=2e-4)
    np.testing.assert_allclose(m.Q, V, rtol=2e-3, atol=2e-4)
    assert float(np.abs(m.P).max()) > 0.01
    # the trained model ranks through the ordinary evaluation path
    scores = m.predictForRanking(train[0][0])
    assert len(scores) == ni and np.isfinite(scores).all()
