"""SimGCL / NGCF on the engine against torch-autograd restatements of the reference's TF graphs
(oracle/tf_models.py), plus the K6 / dense helper kernels on their own.  Needs a GPU."""
import contextlib
import io
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def torch():
    import torch
    assert torch.cuda.is_available()
    return torch


@pytest.fixture(scope='module')
def E():
    from qrec_b200 import engine
    return engine


def _dev(torch, a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize('M,N,K,ta,tb', [(100, 64, 64, 0, 0), (64, 64, 5000, 1, 0), (333, 70, 17, 0, 1),
                                          (65, 129, 33, 1, 1), (2048, 2048, 64, 0, 1), (64, 64, 70001, 1, 0),
                                          (50000, 64, 64, 0, 0), (50001, 64, 64, 0, 1), (9000, 52, 52, 0, 0), (4097, 8, 64, 0, 1)])
def test_sgemm(torch, E, M, N, K, ta, tb):
    g = torch.Generator(device='cuda'); g.manual_seed(M + N + K)
    A = torch.randn((K, M) if ta else (M, K), device='cuda', generator=g)
    B = torch.randn((N, K) if tb else (K, N), device='cuda', generator=g)
    C0 = torch.randn(M, N, device='cuda', generator=g)
    C = C0.clone()
    E.sgemm(A, B, C, trans_a=bool(ta), trans_b=bool(tb), alpha=0.5, beta=-1.5)
    ref = 0.5 * ((A.t() if ta else A).double() @ (B.t() if tb else B).double()) - 1.5 * C0.double()
    scale = ref.abs().max().item()
    assert (C.double() - ref).abs().max().item() <= 2e-5 * scale + 1e-5
    C2 = torch.full((M, N), float('nan'), device='cuda')
    E.sgemm(A, B, C2, trans_a=bool(ta), trans_b=bool(tb))          # beta = 0 must not read C
    assert bool(torch.isfinite(C2).all())


def test_perturb_matches_definition(torch, E):
    from oracle import tf_models
    n, d, eps = 300, 64, 0.1
    rng = np.random.default_rng(0)
    X = rng.standard_normal((n, d)).astype(np.float32)
    X[5, :8] = 0.0                                                  # sign(0) = 0
    acc0 = rng.standard_normal((n, d)).astype(np.float32)
    Xd, acc = _dev(torch, X), _dev(torch, acc0)
    E.simgcl_perturb(Xd, eps, 0xabcdef0123, 17, 3, acc=acc, acc_scale=0.5)
    noise = tf_models.philox_uniform(n, d, 0xabcdef0123, 17, 3)
    nrm = noise / np.sqrt(np.maximum((noise ** 2).sum(1, keepdims=True), 1e-12))
    ref = X + np.sign(X) * nrm * eps
    np.testing.assert_allclose(Xd.cpu().numpy(), ref, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(acc.cpu().numpy(), acc0 + 0.5 * ref, rtol=1e-5, atol=1e-6)
    # the perturbation has norm exactly eps on rows without zeros and keeps the orthant
    delta = Xd.cpu().numpy() - X
    np.testing.assert_allclose(np.linalg.norm(delta[10:], axis=1), eps, rtol=1e-4)
    assert np.all(np.sign(Xd.cpu().numpy()) == np.sign(X))
    # the listed-rows form (compact block + row list with padding, rows of a row-sharded table at offset 100):
    # bit-identical to the whole-table call on those rows, acc touched at the listed rows only
    rows = np.array([7, -1, 150, 3, -1, 199], np.int32)
    Xs = X[100:]                                                      # the "shard": global rows 100..299
    whole, acc_w = _dev(torch, Xs), _dev(torch, acc0[100:])
    E.simgcl_perturb(whole, eps, 0xabcdef0123, 17, 3, acc=acc_w, acc_scale=0.5, row_offset=100)
    block = np.zeros((len(rows), d), np.float32)
    block[rows >= 0] = Xs[rows[rows >= 0]]
    blk, acc_l = _dev(torch, block), _dev(torch, acc0[100:])
    E.simgcl_perturb_listed(blk, _dev(torch, rows), eps, 0xabcdef0123, 17, 3, acc=acc_l, acc_scale=0.5, row_offset=100)
    sel = rows[rows >= 0]
    assert np.array_equal(blk.cpu().numpy()[rows >= 0], whole.cpu().numpy()[sel])
    assert np.all(blk.cpu().numpy()[rows < 0] == 0.0)
    assert np.array_equal(acc_l.cpu().numpy()[sel], acc_w.cpu().numpy()[sel])
    other = np.setdiff1d(np.arange(200), sel)
    assert np.array_equal(acc_l.cpu().numpy()[other], acc0[100:][other])


def test_infonce_block_vs_autograd(torch, E):
    g = torch.Generator(device='cuda'); g.manual_seed(2)
    N, d, b = 500, 64, 257
    T1 = torch.randn(N, d, device='cuda', generator=g, requires_grad=True)
    T2 = torch.randn(N, d, device='cuda', generator=g, requires_grad=True)
    idx = torch.randperm(N, device='cuda', generator=g)[:b]
    z1 = torch.nn.functional.normalize(T1[idx], dim=1)
    z2 = torch.nn.functional.normalize(T2[idx], dim=1)
    ref = -(torch.log(torch.exp((z1 * z2).sum(1) / 0.2) / torch.exp(z1 @ z2.t() / 0.2).sum(1))).sum()
    (0.5 * ref).backward()
    Z1, Z2 = torch.empty(b, d, device='cuda'), torch.empty(b, d, device='cuda')
    n1, n2 = torch.empty(b, device='cuda'), torch.empty(b, device='cuda')
    i32 = idx.int()
    E.gather_normalize(T1.detach(), i32, Z1, n1); E.gather_normalize(T2.detach(), i32, Z2, n2)
    S = torch.empty(b, b, device='cuda')
    E.sgemm(Z1, Z2, S, trans_b=True)
    loss = torch.zeros(1, dtype=torch.float64, device='cuda')
    E.infonce_rows(S, 0.2, loss)
    dZ1, dZ2 = torch.empty_like(Z1), torch.empty_like(Z2)
    E.sgemm(S, Z2, dZ1); E.sgemm(S, Z1, dZ2, trans_a=True)
    G1, G2 = torch.zeros(N, d, device='cuda'), torch.zeros(N, d, device='cuda')
    E.normalize_bwd_scatter(dZ1, Z1, n1, i32, 0.5, G1)
    E.normalize_bwd_scatter(dZ2, Z2, n2, i32, 0.5, G2)
    assert abs(loss.item() - ref.item()) <= 1e-5 * abs(ref.item())
    torch.testing.assert_close(G1, T1.grad, rtol=1e-3, atol=1e-5)
    torch.testing.assert_close(G2, T2.grad, rtol=1e-3, atol=1e-5)


def _graph_model(cls, golden_graph, tmp_path, extra):
    from qrec_b200.util.config import ModelConf
    g = golden_graph
    os.chdir(tmp_path)
    conf = str(g['conf']).replace('model.name=LightGCN', 'model.name=' + cls.__name__) + extra
    train = [[u, i, 1.0] for u, i in zip(g['train_users'].tolist(), g['train_items'].tolist())]
    model = cls(ModelConf.from_string(conf), train, [])
    with contextlib.redirect_stdout(io.StringIO()):
        model.readConfiguration()
        model.initModel()
    return model


def test_simgcl_step_vs_autograd(torch, golden_graph, tmp_path):
    """Losses and d(total)/d(ego) of one SimGCL minibatch (3 encoders, noise, InfoNCE, BPR) against
    the float64 autograd restatement fed with the same Philox noise."""
    from oracle import tf_models
    from qrec_b200.model.ranking.SimGCL import SimGCL
    g = golden_graph
    m = _graph_model(SimGCL, g, tmp_path, 'SimGCL=-n_layer 2 -lambda 0.5 -eps 0.1\n')
    # xavier bound (SimGCL.py:42-44): fan_in = rows, fan_out = d
    bound = (6.0 / (m.num_users + m.emb_size)) ** 0.5
    assert float(m.user_embeddings.abs().max()) <= bound and float(m.user_embeddings.abs().max()) > 0.9 * bound
    adj = m.create_joint_sparse_adjaceny().tocsr()
    N, d = m.num_users + m.num_items, m.emb_size
    ego0 = m.ego.cpu().numpy().astype(np.float64)
    sl = slice(0, 2048)
    u, i, j = g['shuffled_u'][sl], g['shuffled_i'][sl], g['pair_all_j'][sl]
    m.train_step(*(_dev(torch, x) for x in (u, i, j)))
    total, rec, cl = m.losses()
    noise = [[tf_models.philox_uniform(N, d, m.noise_seed, e * 16 + k, 1) for k in range(2)] for e in (1, 2)]
    rrec, rcl, rgrad = tf_models.simgcl_loss_and_grad(adj, ego0, m.num_users, u, i, j, 2, 0.1, 0.5, m.regU, noise)
    assert abs(rec - rrec) <= 1e-4 * abs(rrec) and abs(cl - rcl) <= 1e-4 * abs(rcl)
    got = m._total.cpu().numpy()
    assert np.abs(got - rgrad).max() <= 2e-3 * np.abs(rgrad).max()
    # the ego table moved by one Adam step of size ~lr in the direction of -sign(grad)
    moved = m.ego.cpu().numpy() - ego0
    big = np.abs(rgrad) > 1e-3 * np.abs(rgrad).max()
    assert np.all(np.sign(moved[big]) == -np.sign(rgrad[big]))
    np.testing.assert_allclose(np.abs(moved[big]), m.lRate, rtol=1e-2)


def test_ngcf_step_vs_autograd(torch, golden_graph, tmp_path):
    from oracle import tf_models
    from qrec_b200.model.ranking.NGCF import NGCF, KEEP_PROB
    g = golden_graph
    m = _graph_model(NGCF, g, tmp_path, '')
    adj = m.create_joint_sparse_adjaceny().tocsr()
    N, d = m.num_users + m.num_items, m.emb_size
    # larger embeddings than the 0.005 init so that the activations are not all tiny
    m.ego.mul_(40.0)
    ego0 = m.ego.cpu().numpy().astype(np.float64)
    W1 = [m.weights['W_%d_1' % k].cpu().numpy().astype(np.float64) for k in range(2)]
    W2 = [m.weights['W_%d_2' % k].cpu().numpy().astype(np.float64) for k in range(2)]
    sl = slice(0, 2048)
    u, i, j = g['shuffled_u'][sl], g['shuffled_i'][sl], g['pair_all_j'][sl]
    # inference forward (no dropout) first
    Ue, Ve = m.forward(0)
    _, _, _, _, ref_all = tf_models.ngcf_loss_and_grad(adj, ego0, W1, W2, m.num_users, u, i, j, m.regU)
    np.testing.assert_allclose(np.concatenate([Ue.cpu().numpy(), Ve.cpu().numpy()]), ref_all, rtol=2e-3, atol=2e-5)
    # training step with the engine's Philox dropout masks replayed in the restatement
    loss = m.train_step(*(_dev(torch, x) for x in (u, i, j)))
    masks = [(tf_models.philox_uniform(N, d, m.noise_seed, k, 1).astype(np.float32) < np.float32(KEEP_PROB)).astype(np.float64)
             for k in range(2)]
    assert 0.88 < masks[0].mean() < 0.92
    rl, rgE, rgW1, rgW2, _ = tf_models.ngcf_loss_and_grad(adj, ego0, W1, W2, m.num_users, u, i, j, m.regU, masks, KEEP_PROB)
    assert abs(loss.item() - rl) <= 1e-4 * abs(rl)
    assert np.abs(m._dego.cpu().numpy() - rgE).max() <= 3e-3 * np.abs(rgE).max()
    for k in range(2):
        assert np.abs(m._gw['W_%d_1' % k].cpu().numpy() - rgW1[k]).max() <= 3e-3 * np.abs(rgW1[k]).max()
        assert np.abs(m._gw['W_%d_2' % k].cpu().numpy() - rgW2[k]).max() <= 3e-3 * np.abs(rgW2[k]).max()


@pytest.mark.parametrize('name,extra', [('SimGCL', 'SimGCL=-n_layer 2 -lambda 0.5 -eps 0.1\n'), ('NGCF', ''),
                                        ('SGL', 'SGL=-n_layer 2 -lambda 0.01 -droprate 0.1 -augtype 1 -temp 0.2\n')])   # lambda 0.1 (the yelp2018 conf) lets
                                       # the InfoNCE term swamp BPR on a graph this small: P@10 0.006 vs 0.34 (also with the CPU stand-ins)
def test_graph_models_full_lifecycle(golden_bpr, tmp_path, name, extra):
    """execute() end to end on FilmTrust: trains, evaluates, and lands in a sane quality band."""
    import importlib
    import random
    from qrec_b200.util.config import ModelConf
    g = golden_bpr
    os.chdir(tmp_path)
    cls = getattr(importlib.import_module('qrec_b200.model.ranking.' + name), name)
    conf = (str(g['conf']).replace('model.name=BPR', 'model.name=' + name).replace('num.max.epoch=3', 'num.max.epoch=6')
            .replace('learnRate=-init 0.01', 'learnRate=-init 0.005') + extra)
    train = [[u, i, r] for u, i, r in zip(g['train_users'].tolist(), g['train_items'].tolist(), g['train_rating'].tolist())]
    test = [[u, i, r] for u, i, r in zip(g['test_users'].tolist(), g['test_items'].tolist(), g['test_rating'].tolist())]
    random.seed(2); np.random.seed(2)
    with contextlib.redirect_stdout(io.StringIO()):
        measure = cls(ModelConf.from_string(conf), train, test).execute()
    got = {m.split(':')[0]: float(m.split(':')[1]) for m in measure[1:]}
    # popularity-level ranking on this split is P@10 ~ 0.05; a few epochs must be far above it
    assert got['Precision'] > 0.12 and got['Recall'] > 0.25


@pytest.mark.parametrize('mode', [0, 1, 2])
def test_neumf_step_vs_autograd(torch, golden_graph, tmp_path, mode):
    """Loss, predictions and every parameter gradient of one NeuMF minibatch (reference pointwise
    batch: 1 positive + 4 negatives) against the float64 autograd restatement.  The MLP runs on the
    TF32 tensor-core path (6 chained TF32 products, operands rounded to nearest): every gradient within 2e-2 in
    the Frobenius norm (1.3 % observed for the MLP user table, 0.1-0.5 % for the weights) and its largest single
    entry within 4e-2 of the largest gradient -- a hidden unit whose pre-activation is within TF32 rounding of zero
    switches its ReLU mask, which moves individual entries of the embedding gradients by more than the rounding itself
    (2.1 % was observed for QM with one draw of the biases)."""
    from oracle import tf_models
    from qrec_b200.model.ranking.NeuMF import NeuMF
    g = golden_graph
    m = _graph_model(NeuMF, g, tmp_path, '')
    # make the MLP do something visible: scale the embedding tables up from xavier-on-rows
    for k in ('PG', 'QG', 'PM', 'QM'):
        m.params[k].mul_(8.0)
    torch.manual_seed(20260923)                       # the draw must not depend on which tests ran before
    m.params['b1'].normal_(0, 0.05); m.params['b2'].normal_(0, 0.05); m.params['b3'].normal_(0, 0.05)
    before = {k: v.cpu().numpy().astype(np.float64) for k, v in m.params.items()}
    u, i, y = g['point_b0']
    r = y.astype(np.float32)
    loss = m.train_step(mode, _dev(torch, u), _dev(torch, i), _dev(torch, r))
    ref_loss, ref_g, ref_y = tf_models.neumf_loss_and_grad(before, mode, u, i, r, m.regU)
    # loss reported by the kernel excludes the h-vector regularisers; loss_value() adds them, but
    # uses the post-update parameters -> compare against the restatement minus those terms
    hreg = 0.0
    if mode != 1:
        hreg += m.regU * 0.5 * (before['h_mf'] ** 2).sum()
    if mode == 2:
        hreg += m.regU * 0.5 * 0.25 * ((before['h_mf'] ** 2).sum() + (before['h_mlp'] ** 2).sum())
    assert abs(loss.item() - (ref_loss - hreg)) <= 2e-3 * abs(ref_loss)
    np.testing.assert_allclose(m._y[:len(u)].cpu().numpy(), ref_y, rtol=5e-3, atol=2e-3)
    for k in m.opt_vars[mode]:
        got = m.grads[k].cpu().numpy()
        assert np.linalg.norm(got - ref_g[k]) <= 2e-2 * np.linalg.norm(ref_g[k]) + 1e-7, k
        assert np.abs(got - ref_g[k]).max() <= 4e-2 * np.abs(ref_g[k]).max() + 1e-7, k
    # variables outside this phase's optimiser did not move
    for k in set(m.params) - set(m.opt_vars[mode]):
        assert np.array_equal(m.params[k].cpu().numpy().astype(np.float64), before[k]), k
    # prediction path (all items of one user) agrees with the training-time forward
    pred = {0: m.predict_mf, 1: m.predict_mlp, 2: m.predict_neu}[mode](3)
    assert pred.shape == (m.num_items,) and np.all((pred > 0) & (pred < 1))


def test_neumf_full_lifecycle(golden_bpr, tmp_path):
    import random
    from qrec_b200.util.config import ModelConf
    from qrec_b200.model.ranking.NeuMF import NeuMF
    g = golden_bpr
    os.chdir(tmp_path)
    conf = (str(g['conf']).replace('model.name=BPR', 'model.name=NeuMF').replace('num.max.epoch=3', 'num.max.epoch=10')
            .replace('learnRate=-init 0.01', 'learnRate=-init 0.002'))
    train = [[u, i, r] for u, i, r in zip(g['train_users'].tolist(), g['train_items'].tolist(), g['train_rating'].tolist())]
    test = [[u, i, r] for u, i, r in zip(g['test_users'].tolist(), g['test_items'].tolist(), g['test_rating'].tolist())]
    random.seed(4); np.random.seed(4)
    with contextlib.redirect_stdout(io.StringIO()):
        measure = NeuMF(ModelConf.from_string(conf), train, test).execute()
    got = {m.split(':')[0]: float(m.split(':')[1]) for m in measure[1:]}
    assert got['Precision'] > 0.12 and got['Recall'] > 0.25


@pytest.mark.parametrize('aug', [1, 2, 0])
def test_sgl_step_vs_autograd(torch, golden_graph, tmp_path, aug):
    """One SGL minibatch (f-4 sibling model; model/ranking/SGL.py): three LightGCN encoders over the full graph and
    the epoch's two augmented views (rebuilt on the device by the f-2 kernels), BPR + merged user/item InfoNCE,
    per-view Horner backward -- against the float64 autograd restatement fed with the SAME view matrices (read back
    from the device; their construction has its own test, test_gpu_adjacency.py)."""
    import scipy.sparse as sp
    from oracle import tf_models
    from qrec_b200.model.ranking.SGL import SGL
    g = golden_graph
    m = _graph_model(SGL, g, tmp_path, 'SGL=-n_layer 2 -lambda 0.1 -droprate 0.3 -augtype %d -temp 0.2\n' % aug)
    N, d = m.num_users + m.num_items, m.emb_size
    m.ego.mul_(20.0)                                   # make the InfoNCE term visible next to the 0.005-sigma init
    views = m.build_views(0)
    assert len(views) == 2 and all(len(v) == 2 for v in views)
    to_sp = lambda a: sp.csr_matrix((a.vals.cpu().numpy().astype(np.float64), a.cols.cpu().numpy(), a.rowptr.cpu().numpy()), shape=(N, N))  # noqa: E731
    sp_views = [[to_sp(a) for a in v] for v in views]
    full = to_sp(m.norm_adj)
    for v in sp_views:                                 # a view keeps ~70 % of the edges (node dropout: ~49 %), never more
        for a in v:
            assert 0.3 * full.nnz < a.nnz < (0.85 if aug else 0.7) * full.nnz
    if aug == 2:
        assert views[0][0].nnz != views[0][1].nnz or not torch.equal(views[0][0].cols, views[0][1].cols)   # a fresh draw per layer
    else:
        assert views[0][0] is views[0][1]
    ego0 = m.ego[:, :d].cpu().numpy().astype(np.float64)
    sl = slice(0, 2048)
    u, i, j = g['shuffled_u'][sl], g['shuffled_i'][sl], g['pair_all_j'][sl]
    m.train_step(*(_dev(torch, x) for x in (u, i, j)))
    rec, ssl = m.losses()
    rrec, rssl, rgrad = tf_models.sgl_loss_and_grad(full, sp_views, ego0, m.num_users, u, i, j, 2, 0.1, 0.2, m.regU)
    assert abs(rec - rrec) <= 1e-4 * abs(rrec) and abs(ssl - rssl) <= 1e-4 * abs(rssl)
    got = m._total[:, :d].cpu().numpy()
    assert np.abs(got - rgrad).max() <= 2e-3 * np.abs(rgrad).max()
    moved = m.ego[:, :d].cpu().numpy() - ego0
    big = np.abs(rgrad) > 1e-3 * np.abs(rgrad).max()
    assert np.all(np.sign(moved[big]) == -np.sign(rgrad[big]))


@pytest.mark.parametrize('rows,cols', [(10240, 320), (10240, 64), (77, 5), (1, 128), (3000, 300), (0, 16)])
def test_gemv_t(torch, E, rows, cols):
    """qrec_gemv_t_f32: A^T v and plain column sums (NeuMF's head-vector / bias gradients), on a row slice of a wider
    workspace too, overwrite and accumulate."""
    g = torch.Generator(device='cuda'); g.manual_seed(rows + cols)
    W = torch.randn(rows + 3, cols, device='cuda', generator=g)
    A = W[:rows]
    v = torch.randn(rows, device='cuda', generator=g)
    out = torch.full((cols,), float('nan'), device='cuda')
    E.gemv_t(A, v, out, alpha=0.5)
    ref = 0.5 * (A.double().t() @ v.double())
    assert float((out.double() - ref).abs().max()) <= 1e-4 * max(1.0, float(ref.abs().max()))
    E.gemv_t(A, None, out, alpha=2.0, beta=1.0)
    ref = ref + 2.0 * A.double().sum(0)
    assert float((out.double() - ref).abs().max()) <= 1e-4 * max(1.0, float(ref.abs().max()))
