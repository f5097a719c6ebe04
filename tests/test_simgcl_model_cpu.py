"""Host logic of the SimGCL drop-in (a18) without a GPU: device stubbed to 'cpu', every kernel replaced by a
restatement of its documented contract.  Checks the class's own composition -- three encoders whose LAST layer
(product and noise) is evaluated on the batch's rows only, InfoNCE on the unique users / items, the collapsed backward
pass with a scattered first layer, dense Adam, padded width 10 -> 12 -- against the float64 AUTOGRAD restatement of
model/ranking/SimGCL.py:22-38,60-78,92-108 (oracle/tf_models.simgcl_loss_and_grad), which propagates every row."""
import contextlib
import io

import numpy as np

from conftest import row_list_kernel_stand_ins
from qrec_b200.util.config import ModelConf


def test_simgcl_step_equals_autograd_restatement(golden_graph, monkeypatch, tmp_path):
    import torch
    from oracle import tf_models
    from qrec_b200.model.ranking.SimGCL import SimGCL
    import test_sgl_model_cpu as S
    g = golden_graph
    S._stub(monkeypatch)
    calls = row_list_kernel_stand_ins(monkeypatch)
    monkeypatch.chdir(tmp_path)
    n_tr = 6000
    train = [[u, i, 1.0] for u, i in zip(g['train_users'][:n_tr].tolist(), g['train_items'][:n_tr].tolist())]
    conf = ModelConf.from_string(str(g['conf']).replace('model.name=LightGCN', 'model.name=SimGCL').replace('num.factors=64', 'num.factors=10')
                                 + 'SimGCL=-n_layer 2 -lambda 0.5 -eps 0.1\n')
    m = SimGCL(conf, train, [])
    with contextlib.redirect_stdout(io.StringIO()):
        m.readConfiguration()
        m.initModel()
    N, d = m.num_users + m.num_items, m.emb_size
    assert m.ego.shape == (N, 12) and float(m.ego[:, 10:].abs().sum()) == 0.0
    m.ego.mul_(20.0)                                       # xavier-on-rows is tiny: make the contrastive term visible
    adj = m.create_joint_sparse_adjaceny().tocsr()
    u_all, i_all, _ = m.data.training_ids()
    rng = np.random.default_rng(2)
    for step in (1, 2):
        ego0 = m.ego[:, :d].numpy().astype(np.float64).copy()
        pick = rng.choice(len(u_all), 512, replace=False)
        u, i = u_all[pick].astype(np.int32), i_all[pick].astype(np.int32)
        j = rng.integers(0, m.num_items, 512).astype(np.int32)
        del calls[:]
        m.train_step(torch.from_numpy(u), torch.from_numpy(i), torch.from_numpy(j))
        total, rec, cl = m.losses()
        noise = [[tf_models.philox_uniform(N, 12, m.noise_seed, e * 16 + k, step)[:, :d] for k in range(2)] for e in (1, 2)]
        rrec, rcl, rgrad = tf_models.simgcl_loss_and_grad(adj, ego0, m.num_users, u, i, j, 2, 0.1, 0.5, m.regU, noise)
        assert abs(rec - rrec) <= 1e-5 * abs(rrec) and abs(cl - rcl) <= 1e-5 * abs(rcl), (step, rec, rrec, cl, rcl)
        got = m._total[:, :d].numpy()
        assert np.abs(got - rgrad).max() <= 1e-4 * np.abs(rgrad).max(), step
        assert float(m._total[:, d:].abs().sum()) == 0.0 and float(m.ego[:, d:].abs().sum()) == 0.0
        # the three last layers ran on the row list (the perturbed ones with their noise), the first backward layer as a scatter
        assert calls == ['rows', 'rows', 'perturb_listed', 'rows', 'perturb_listed', 'scatter_rows'], calls
    # the exported tables use the full propagation (every row is read)
    del calls[:]
    m.saveModel()
    assert calls == [] and m.bestU.shape == (m.num_users, d)
