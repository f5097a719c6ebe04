"""End-to-end parity of the drop-in classes on the GPU: the reference's own FilmTrust run
(seeded, 3 epochs) replayed through qrec_b200.model.ranking.BPR, and LightGCN against the
restatement of its TF graph."""
import contextlib
import io
import os
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _lists(g):
    train = [[u, i, r] for u, i, r in zip(g['train_users'].tolist(), g['train_items'].tolist(), g['train_rating'].tolist())]
    test = [[u, i, r] for u, i, r in zip(g['test_users'].tolist(), g['test_items'].tolist(), g['test_rating'].tolist())]
    return train, test


def _run_bpr(g, tmp_path, extra=''):
    from qrec_b200.util.config import ModelConf
    from qrec_b200.model.ranking.BPR import BPR
    os.chdir(tmp_path)
    conf = ModelConf.from_string(str(g['conf']) + extra)
    train, test = _lists(g)
    random.setstate((3, tuple(int(x) for x in g['mt_state_after_split']), None))
    np.random.seed(0)
    model = BPR(conf, train, test)
    losses, lrs = [], []
    orig = model.isConverged

    def spy(epoch):
        losses.append(model.loss)
        before = model.lRate
        r = orig(epoch)
        lrs.append((before, model.lRate))
        return r
    model.isConverged = spy
    out = io.StringIO()
    with contextlib.redirect_stdout(out):
        measure = model.execute()
    return model, measure, losses, lrs, out.getvalue()


def test_bpr_dropin_reproduces_reference_run(golden_bpr, tmp_path):
    g = golden_bpr
    model, measure, losses, lrs, log = _run_bpr(g, tmp_path)
    # epoch losses, learning-rate schedule, final tables, generator state, metrics
    np.testing.assert_allclose(losses, g['loss'], rtol=1e-10)
    assert np.array_equal(np.array(lrs), g['lrate'])
    assert model.P.dtype == np.float64
    np.testing.assert_allclose(model.P, g['P_epoch3'], rtol=2e-7, atol=1e-8)
    np.testing.assert_allclose(model.Q, g['Q_epoch3'], rtol=2e-7, atol=1e-8)
    assert np.array_equal(np.array(random.getstate()[1], dtype=np.uint32), g['mt_state_after_epoch'][2])
    assert [m.strip() for m in measure] == g['measure'].tolist()
    # top-10 lists of the first 64 test users: same items, same order, same '*' marks
    for mine, ref in zip(model.recOutput[1:65], g['rec_lines'].tolist()):
        strip = lambda line: [(p.split(',')[0], p.endswith('*')) for p in line.strip().split(' (')[1:]]  # noqa: E731
        assert mine.split(':')[0] == ref.split(':')[0] and strip(mine) == strip(ref)
    assert 'BPR [1] epoch 3: loss = 4997.0631' in log


def test_bpr_dropin_fp32_parity_mode(golden_bpr, tmp_path):
    g = golden_bpr
    model, measure, losses, _, _ = _run_bpr(g, tmp_path, 'engine=-mode parity -precision f32\n')
    assert model.P.dtype == np.float32
    np.testing.assert_allclose(losses, g['loss'], rtol=2e-6)
    assert np.abs(model.P - g['P_epoch3']).max() <= 2e-5 * np.abs(g['P_epoch3']).max()
    ref = {m.split(':')[0]: float(m.split(':')[1]) for m in g['measure'].tolist()[1:]}
    got = {m.split(':')[0]: float(m.split(':')[1]) for m in measure[1:]}
    for k in ref:
        assert abs(got[k] - ref[k]) <= 0.01


def test_bpr_dropin_fast_mode_quality(golden_bpr, tmp_path):
    """Throughput kernel through the same class surface: 30 epochs reach the reference's ranking
    quality band (the reference after 3 sequential epochs scores P@10 = 0.335)."""
    g = golden_bpr
    conf = str(g['conf']).replace('num.max.epoch=3', 'num.max.epoch=30')
    from qrec_b200.util.config import ModelConf
    from qrec_b200.model.ranking.BPR import BPR
    os.chdir(tmp_path)
    train, test = _lists(g)
    random.seed(3); np.random.seed(3)
    with contextlib.redirect_stdout(io.StringIO()):
        measure = BPR(ModelConf.from_string(conf + 'engine=-mode fast\n'), train, test).execute()
    got = {m.split(':')[0]: float(m.split(':')[1]) for m in measure[1:]}
    assert got['Precision'] > 0.30 and got['NDCG'] > 0.45


def test_lightgcn_steps_match_restatement(golden_graph, tmp_path):
    """3 minibatches of LightGCN.trainModel's step against oracle.lightgcn_step (numpy/scipy
    restatement of model/ranking/LightGCN.py:11-39 + TF1 Adam), same initial tables, the
    reference's own shuffled batches and negatives."""
    import torch
    from oracle import bpr_oracle as O
    from qrec_b200.util.config import ModelConf
    from qrec_b200.model.ranking.LightGCN import LightGCN
    g = golden_graph
    os.chdir(tmp_path)
    train = [[u, i, 1.0] for u, i in zip(g['train_users'].tolist(), g['train_items'].tolist())]
    model = LightGCN(ModelConf.from_string(str(g['conf'])), train, [])
    with contextlib.redirect_stdout(io.StringIO()):
        model.readConfiguration()
        model.initModel()
    nu, ni = model.num_users, model.num_items
    # adjacency: identical to the reference's scipy matrix
    adj = model.create_joint_sparse_adjaceny().tocsr(); adj.sort_indices()
    assert np.array_equal(adj.indptr, g['adj_indptr']) and np.array_equal(adj.indices, g['adj_indices'])
    np.testing.assert_allclose(adj.data, g['adj_data'], rtol=1e-6)
    assert model.n_layers == 3
    U = model.user_embeddings.cpu().numpy().copy(); V = model.item_embeddings.cpu().numpy().copy()
    assert abs(U.std() - 0.005 * 0.88) < 5e-4 and np.abs(U).max() <= 0.01 + 1e-7     # truncated at 2 sigma
    mU, vU, mV, vV = (np.zeros_like(x) for x in (U, U, V, V))
    bs = 2048
    su, si, sj = g['shuffled_u'], g['shuffled_i'], g['pair_all_j']
    for step in range(3):
        sl = slice(step * bs, (step + 1) * bs)
        ref_loss = O.lightgcn_step(adj, U, V, mU, vU, mV, vV, su[sl], si[sl], sj[sl], 3, model.lRate, model.regU, step + 1)
        loss = model.train_step(*(torch.from_numpy(np.ascontiguousarray(x[sl])).cuda() for x in (su, si, sj)))
        assert abs(loss.item() - ref_loss) <= 1e-5 * abs(ref_loss)
        np.testing.assert_allclose(model.user_embeddings.cpu().numpy(), U, rtol=2e-3, atol=2e-6)
        np.testing.assert_allclose(model.item_embeddings.cpu().numpy(), V, rtol=2e-3, atol=2e-6)
    Ue, Ve = model.propagate()
    rUe, rVe, _ = O.lightgcn_forward(adj, U, V, 3)
    np.testing.assert_allclose(Ue.cpu().numpy(), rUe, rtol=2e-3, atol=2e-6)


def test_lightgcn_dropin_trains_and_ranks(golden_bpr, tmp_path):
    """Whole life cycle through execute(): 8 epochs of LightGCN on FilmTrust beat the popularity
    floor by a wide margin (sanity of the composed step + samplers + evaluation)."""
    from qrec_b200.util.config import ModelConf
    from qrec_b200.model.ranking.LightGCN import LightGCN
    g = golden_bpr
    os.chdir(tmp_path)
    conf = (str(g['conf']).replace('model.name=BPR', 'model.name=LightGCN').replace('num.max.epoch=3', 'num.max.epoch=8')
            .replace('learnRate=-init 0.01', 'learnRate=-init 0.005') + 'LightGCN=-n_layer 2\n')
    train, test = _lists(g)
    random.seed(1); np.random.seed(1)
    with contextlib.redirect_stdout(io.StringIO()):
        measure = LightGCN(ModelConf.from_string(conf), train, test).execute()
    got = {m.split(':')[0]: float(m.split(':')[1]) for m in measure[1:]}
    assert got['Precision'] > 0.25 and got['Recall'] > 0.4


@pytest.mark.parametrize('model,extra', [('BPR', 'engine=-mode fast\n'), ('BPR', ''), ('LightGCN', 'LightGCN=-n_layer 2\n'),
                                         ('SimGCL', 'SimGCL=-n_layer 2 -lambda 0.5 -eps 0.1\n')])
def test_shipped_width_num_factors_50(golden_bpr, tmp_path, model, extra):
    """The reference's shipped confs use num.factors=50 (config/BPR.conf:6, LightGCN.conf:6): rows of
    200 bytes.  Parity mode takes any d; the 16-byte row kernels run on tables padded to 52 columns
    whose two extra columns must stay exactly zero and never leak into the exported tables."""
    import importlib
    from qrec_b200.util.config import ModelConf
    g = golden_bpr
    os.chdir(tmp_path)
    cls = getattr(importlib.import_module('qrec_b200.model.ranking.' + model), model)
    conf = (str(g['conf']).replace('model.name=BPR', 'model.name=' + model).replace('num.factors=64', 'num.factors=50')
            .replace('num.max.epoch=3', 'num.max.epoch=2') + extra)
    train, test = _lists(g)
    random.seed(6); np.random.seed(6)
    m = cls(ModelConf.from_string(conf), train, test)
    with contextlib.redirect_stdout(io.StringIO()):
        measure = m.execute()
    U, V = (m.P, m.Q) if model == 'BPR' else (m.U, m.V)
    assert U.shape == (m.num_users, 50) and V.shape == (m.num_items, 50)
    assert np.isfinite(U).all() and np.isfinite(V).all()
    if model != 'BPR':
        assert m.ego.shape[1] == 52 and float(m.ego[:, 50:].abs().max()) == 0.0
    assert len(measure) == 5 and measure[0].startswith('Top 10')


def test_batched_device_evaluation_equals_host_flow(golden_bpr, tmp_path):
    """`-eval gpu` (scores = sgemm, rated -> 0, top-N) gives the reference's metrics and the same
    top-10 lists as the per-user host flow on the reference's own 3-epoch FilmTrust model."""
    g = golden_bpr
    model, measure, _, _, _ = _run_bpr(g, tmp_path, 'engine=-mode parity -precision f64 -eval gpu\n')
    assert [m.strip() for m in measure] == g['measure'].tolist()
    for mine, ref in zip(model.recOutput[1:65], g['rec_lines'].tolist()):
        strip = lambda line: [(p.split(',')[0], p.endswith('*')) for p in line.strip().split(' (')[1:]]  # noqa: E731
        assert strip(mine) == strip(ref)
    # the helper on its own: rated items score exactly 0, order is descending
    import torch
    from qrec_b200.evaluate import batched_top_n
    U, V = model.device_tables()
    ids, vals = batched_top_n(U, V, np.arange(50), model.data.rated_csr(), 10, block=16)
    assert np.all(np.diff(vals, axis=1) <= 0)
    csr = model.data.rated_csr()
    for r in range(50):
        rated = set(csr.sorted_cols[csr.sorted_rowptr[r]:csr.sorted_rowptr[r + 1]].tolist())
        assert all((k not in rated) or v == 0.0 for k, v in zip(ids[r].tolist(), vals[r].tolist()))


def test_user_sharded_lightgcn_world1_equals_dropin_step(golden_graph, tmp_path):
    """parallel.UserShardedLightGCN (users partitioned / items replicated; here world = 1, so the
    all-reduces are no-ops) against the drop-in LightGCN class on the reference's FilmTrust graph:
    same losses, same gradients, same tables -- this exercises the CUDA kernels of the sharded path,
    including the sparse first backward layer through the A_iu / A_ui edge lists."""
    import torch
    from qrec_b200 import parallel
    from qrec_b200.util.config import ModelConf
    from qrec_b200.model.ranking.LightGCN import LightGCN
    g = golden_graph
    os.chdir(tmp_path)
    train = [[u, i, 1.0] for u, i in zip(g['train_users'].tolist(), g['train_items'].tolist())]
    ref = LightGCN(ModelConf.from_string(str(g['conf'])), train, [])
    with contextlib.redirect_stdout(io.StringIO()):
        ref.readConfiguration()
        ref.initModel()
    U, I = ref.num_users, ref.num_items
    adj = ref.norm_adj
    A_ui, A_iu, (lo, hi) = parallel.shard_bipartite_by_user(adj.rowptr, adj.cols, adj.vals, U, I, 0, 1)
    assert (lo, hi) == (0, U)
    m = parallel.UserShardedLightGCN(A_ui, A_iu, ref.ego[:U].clone(), ref.ego[U:].clone(), ref.n_layers, ref.lRate, ref.regU, 0)
    su, si, sj = g['shuffled_u'], g['shuffled_i'], g['pair_all_j']
    for step in range(3):
        sl = slice(step * 2048, (step + 1) * 2048)
        b = [torch.from_numpy(np.ascontiguousarray(x[sl])).cuda() for x in (su, si, sj)]
        l_ref = ref.train_step(*b).item()
        l = m.train_step(*b).item()
        assert abs(l - l_ref) <= 1e-5 * abs(l_ref)
        gref = ref._total
        gtot = torch.cat([m.tot_u, m.tot_i])
        assert float((gtot - gref).abs().max()) <= 2e-3 * float(gref.abs().max())
        torch.testing.assert_close(torch.cat([m.Eu, m.Ei]), ref.ego, rtol=2e-3, atol=2e-4)


def test_user_sharded_lightgcn_graph_replay_equals_dropin_step(golden_graph, tmp_path):
    """train_step_graphed (the step replayed from a CUDA graph: first call eager, second captures and replays, later
    ones replay; Adam's step factor read from device memory) against the drop-in class over five different
    minibatches -- the replayed launch sequence must consume the NEW minibatch and the NEW Adam step each time."""
    import torch
    from qrec_b200 import parallel
    from qrec_b200.util.config import ModelConf
    from qrec_b200.model.ranking.LightGCN import LightGCN
    g = golden_graph
    os.chdir(tmp_path)
    train = [[u, i, 1.0] for u, i in zip(g['train_users'].tolist(), g['train_items'].tolist())]
    ref = LightGCN(ModelConf.from_string(str(g['conf'])), train, [])
    with contextlib.redirect_stdout(io.StringIO()):
        ref.readConfiguration()
        ref.initModel()
    U, I = ref.num_users, ref.num_items
    adj = ref.norm_adj
    A_ui, A_iu, _ = parallel.shard_bipartite_by_user(adj.rowptr, adj.cols, adj.vals, U, I, 0, 1)
    m = parallel.UserShardedLightGCN(A_ui, A_iu, ref.ego[:U].clone(), ref.ego[U:].clone(), ref.n_layers, ref.lRate, ref.regU, 0)
    su, si, sj = g['shuffled_u'], g['shuffled_i'], g['pair_all_j']
    for step in range(5):
        sl = slice(step * 2048, (step + 1) * 2048)
        b = [torch.from_numpy(np.ascontiguousarray(x[sl])).cuda() for x in (su, si, sj)]
        l_ref = ref.train_step(*b).item()
        l = m.train_step_graphed(*b).item()
        assert m.graph_error is None, m.graph_error
        assert abs(l - l_ref) <= 1e-5 * abs(l_ref)
        gref = ref._total
        gtot = torch.cat([m.tot_u, m.tot_i])
        assert float((gtot - gref).abs().max()) <= 2e-3 * float(gref.abs().max())
        torch.testing.assert_close(torch.cat([m.Eu, m.Ei]), ref.ego, rtol=2e-3, atol=2e-4)
    assert m.step == 5 and m._graphs[2048]['graph'] is not None


def test_scale_bpr_from_interaction_table(golden_bpr, tmp_path):
    """f-3 end to end: InteractionTable -> ScaleBPR (fused user-major epochs, device negatives) trains
    FilmTrust to the quality band of the drop-in class, follows the reference's lr schedule rules, and
    ranks through the batched device evaluation."""
    from qrec_b200.data.interactions import InteractionTable
    from qrec_b200.scale import ScaleBPR
    from qrec_b200.util.measure import Measure
    g = golden_bpr
    train, test = _lists(g)
    table = InteractionTable.from_records(train)
    assert table.user_names.tolist() == g['user_names'].tolist() and table.item_names.tolist() == g['item_names'].tolist()
    np.random.seed(0)
    m = ScaleBPR(table, emb_size=64, lr=0.01, reg_u=0.001, reg_i=0.001, seed=5).fit(30)
    P0 = np.random.RandomState(0).rand(len(g['user_names']), 64) / 3
    losses = [h[1] for h in m.history]
    assert losses[0] > losses[5] > losses[-1] and abs(losses[0] - g['loss'][0]) < 0.15 * g['loss'][0]
    for (e0, l0, d0, lr0), (e1, l1, d1, lr1) in zip(m.history, m.history[1:]):
        expect = lr0 if e0 == 1 else lr0 * (1.05 if d0 > 0 else 0.5)        # rule applied after epoch e0
        assert abs(lr1 - min(expect, 1.0)) < 1e-12
    P, Q = m.tables()
    assert P.shape == P0.shape and np.isfinite(P).all() and np.isfinite(Q).all()
    # evaluate on the test users that exist in training
    lut_u = {n: k for k, n in enumerate(table.user_names.tolist())}
    origin = {}
    for u, i, r in test:
        if u in lut_u:
            origin.setdefault(u, {})[i] = r
    users = list(origin)
    ids, vals = m.top_n([lut_u[u] for u in users], 10)
    names = table.item_names
    res = {u: [(names[k], float(v)) for k, v in zip(ids[r], vals[r])] for r, u in enumerate(users)}
    out = Measure.rankingMeasure(origin, res, [10])
    got = {x.split(':')[0]: float(x.split(':')[1]) for x in out[1:]}
    assert got['Precision'] > 0.30 and got['NDCG'] > 0.45
    # the vectorised evaluation path gives the same numbers as the dict-based Measure
    fast = m.evaluate(test, tops=(10,))
    for a, b in zip(fast[1:], out[1:]):
        assert a.split(':')[0] == b.split(':')[0] and abs(float(a.split(':')[1]) - float(b.split(':')[1])) < 1e-9


def test_user_sharded_simgcl_world1_equals_dropin_step(golden_graph, tmp_path):
    """parallel.UserShardedSimGCL (BASELINE config 5's decomposition: user rows sharded, item rows replicated;
    world = 1 here, so the collectives are no-ops) against the drop-in SimGCL class -- itself checked against the
    float64 autograd restatement of model/ranking/SimGCL.py -- on the reference's FilmTrust graph: same Philox
    noise (function of the global row), same losses, same collapsed backward pass, same Adam step."""
    import torch
    from qrec_b200 import parallel
    from qrec_b200.util.config import ModelConf
    from qrec_b200.model.ranking.SimGCL import SimGCL
    g = golden_graph
    os.chdir(tmp_path)
    train = [[u, i, 1.0] for u, i in zip(g['train_users'].tolist(), g['train_items'].tolist())]
    conf = str(g['conf']).replace('model.name=LightGCN', 'model.name=SimGCL') + 'SimGCL=-n_layer 2 -lambda 0.5 -eps 0.1\n'
    ref = SimGCL(ModelConf.from_string(conf), train, [])
    with contextlib.redirect_stdout(io.StringIO()):
        ref.readConfiguration()
        ref.initModel()
    U, I = ref.num_users, ref.num_items
    adj = ref.norm_adj
    A_ui, A_iu, _ = parallel.shard_bipartite_by_user(adj.rowptr, adj.cols, adj.vals, U, I, 0, 1)
    m = parallel.UserShardedSimGCL(A_ui, A_iu, ref.ego[:U].clone(), ref.ego[U:].clone(), ref.n_layers, ref.lRate, ref.regU, 0,
                                   U, ref.cl_rate, ref.eps, noise_seed=ref.noise_seed, d_valid=ref.emb_size)
    su, si, sj = g['shuffled_u'], g['shuffled_i'], g['pair_all_j']
    for step in range(3):
        sl = slice(step * 2048, (step + 1) * 2048)
        b = [torch.from_numpy(np.ascontiguousarray(x[sl])).cuda() for x in (su, si, sj)]
        ref.train_step(*b)
        m.train_step(*b)
        t_ref, rec_ref, cl_ref = ref.losses()
        t, rec, cl = m.losses()
        assert abs(rec - rec_ref) <= 1e-5 * abs(rec_ref) and abs(cl - cl_ref) <= 1e-5 * abs(cl_ref)
        gref = ref._total
        gtot = torch.cat([m.tot_u, m.tot_i])
        assert float((gtot - gref).abs().max()) <= 2e-3 * float(gref.abs().max())
        torch.testing.assert_close(torch.cat([m.Eu, m.Ei]), ref.ego, rtol=2e-3, atol=2e-4)


def test_tbpr_dropin_device_run_equals_oracle_run(golden_bpr, monkeypatch, tmp_path):
    """f-4 sibling model TBPR (model/ranking/TBPR.py): the same seeded life cycle twice in this process -- once with
    the K1 parity kernel on the GPU (float64), once with the kernel replaced by the pinned oracle on the CPU (the
    configuration tests/test_tbpr_cpu.py proves equal to the unmodified reference class) -- gives the same losses,
    learning rates and tables; fast mode (one user-major launch per epoch) lands in the same quality band."""
    import test_tbpr_cpu as T
    from test_bpr_model_cpu import _stub_engine
    from qrec_b200.model.ranking.TBPR import TBPR
    os.chdir(tmp_path)
    train, test, rel = T._data(golden_bpr)
    m_gpu, losses_gpu, measure_gpu = T._run(TBPR, train, test, rel, T.CONF)
    m_fast, losses_fast, measure_fast = T._run(TBPR, train, test, rel, T.CONF.replace('num.max.epoch=3', 'num.max.epoch=12')
                                               + 'engine=-mode fast\n')
    with monkeypatch.context() as mp:
        _stub_engine(mp, [])
        m_cpu, losses_cpu, measure_cpu = T._run(TBPR, train, test, rel, T.CONF)
    assert [l[1] for l in losses_gpu] == [l[1] for l in losses_cpu]
    np.testing.assert_allclose([l[0] for l in losses_gpu], [l[0] for l in losses_cpu], rtol=1e-10)
    np.testing.assert_allclose(m_gpu.P, m_cpu.P, rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(m_gpu.Q, m_cpu.Q, rtol=1e-10, atol=1e-13)
    assert [x.strip() for x in measure_gpu] == [x.strip() for x in measure_cpu]
    prec = lambda meas: float([x for x in meas if x.startswith('Precision')][0].split(':')[1])   # noqa: E731
    assert prec(measure_fast) >= 0.8 * prec(measure_cpu)
    assert losses_fast[-1][0] < losses_fast[0][0]
