"""Pins the CPU oracle (oracle/) against vectors recorded from the UNMODIFIED reference
(oracle/gen_golden.py -> tests/golden/*.npz).  CPU only."""
import numpy as np

from oracle import bpr_oracle as O
from oracle import c_oracle
from conftest import rows_and_sets

LR, REG = 0.01, 0.001


def _init_tables(nu, ni, d=64):
    # base/iterativeRecommender.py:37-38 -- legacy numpy stream, P first then Q
    np.random.seed(0)
    P = np.random.rand(nu, d) / 3
    Q = np.random.rand(ni, d) / 3
    return P, Q


def test_sampler_stream_matches_reference(golden_bpr, bpr_ids):
    u, i, nu, ni = bpr_ids
    rows, sets = rows_and_sets(u, i, nu)
    rng = O.make_rng(golden_bpr['mt_state_after_split'])
    for ep in range(3):
        t = O.sample_bpr_epoch(rng, rows, sets, ni)
        assert np.array_equal(t, golden_bpr['triples_epoch'][ep])
        # base/iterativeRecommender.py:101 -- the epoch ends with shuffle(trainingData)
        O.shuffle_pairs(rng, u.copy(), i.copy())
        assert np.array_equal(O.rng_state(rng), golden_bpr['mt_state_after_epoch'][ep])


def test_numpy_sgd_matches_reference_epoch1(golden_bpr, bpr_ids):
    _, _, nu, ni = bpr_ids
    P, Q = _init_tables(nu, ni)
    loss = O.bpr_sgd_sequential(P, Q, golden_bpr['triples_epoch'][0], LR, REG, REG)
    loss += O.epoch_loss_reg(P, Q, REG, REG)
    np.testing.assert_allclose(P, golden_bpr['P_epoch1'], rtol=1e-11, atol=1e-14)
    np.testing.assert_allclose(Q, golden_bpr['Q_epoch1'], rtol=1e-11, atol=1e-14)
    assert abs(loss - golden_bpr['loss'][0]) < 1e-7 * golden_bpr['loss'][0]


def test_c_oracle_three_epochs_losses_lr_schedule(golden_bpr, bpr_ids):
    _, _, nu, ni = bpr_ids
    P, Q = _init_tables(nu, ni)
    lr, last = LR, 0.0
    for ep in range(3):
        t = golden_bpr['triples_epoch'][ep]
        assert lr == golden_bpr['lrate'][ep][0]
        loss = c_oracle.bpr_sgd_sequential(P, Q, t[:, 0], t[:, 1], t[:, 2], lr, REG, REG)
        loss += O.epoch_loss_reg(P, Q, REG, REG)
        assert abs(loss - golden_bpr['loss'][ep]) < 1e-9 * golden_bpr['loss'][ep]
        if ep == 0:
            np.testing.assert_allclose(P, golden_bpr['P_epoch1'], rtol=1e-10, atol=1e-13)
            np.testing.assert_allclose(Q, golden_bpr['Q_epoch1'], rtol=1e-10, atol=1e-13)
        lr = O.update_learning_rate(lr, 1.0, ep + 1, last, loss)
        assert lr == golden_bpr['lrate'][ep][1]
        last = loss
    np.testing.assert_allclose(P, golden_bpr['P_epoch3'], rtol=2e-7, atol=1e-8)
    np.testing.assert_allclose(Q, golden_bpr['Q_epoch3'], rtol=2e-7, atol=1e-8)


def test_fp32_sequential_within_1e5_of_reference(golden_bpr, bpr_ids):
    """What the fp32 parity kernel is allowed to lose against the float64 reference."""
    _, _, nu, ni = bpr_ids
    P, Q = _init_tables(nu, ni)
    P, Q = P.astype(np.float32), Q.astype(np.float32)
    t = golden_bpr['triples_epoch'][0]
    c_oracle.bpr_sgd_sequential(P, Q, t[:, 0], t[:, 1], t[:, 2], LR, REG, REG)
    for got, ref in ((P, golden_bpr['P_epoch1']), (Q, golden_bpr['Q_epoch1'])):
        assert np.abs(got - ref).max() <= 1e-5 * np.abs(ref).max()


def test_pairwise_sampler_matches_reference(golden_graph, graph_ids):
    g = golden_graph
    u, i, nu, ni = graph_ids
    _, sets = rows_and_sets(u, i, nu)
    rng = O.make_rng(g['mt_state_before_pairwise'])
    su, si = O.shuffle_pairs(rng, u, i)
    assert np.array_equal(su, g['shuffled_u']) and np.array_equal(si, g['shuffled_i'])
    js = []
    bs = 2048
    for b in range(0, len(su), bs):
        js.append(O.sample_pairwise(rng, su[b:b + bs].tolist(), sets, ni))
    assert len(js) == int(g['pair_num_batches'])
    assert np.array_equal(np.concatenate(js), g['pair_all_j'])
    assert np.array_equal(np.stack([su[:bs], si[:bs], js[0]]), g['pair_first'])
    assert np.array_equal(O.rng_state(rng), g['mt_state_after_pairwise'])


def test_pointwise_sampler_matches_reference(golden_graph, graph_ids):
    g = golden_graph
    u, i, nu, ni = graph_ids
    _, sets = rows_and_sets(u, i, nu)
    rng = O.make_rng(g['mt_state_before_pointwise'])
    su, si = g['shuffled_u'], g['shuffled_i']       # pointwise iterates the list as it stands
    for b, key in ((0, 'point_b0'), (1, 'point_b1')):
        sl = slice(b * 2048, (b + 1) * 2048)
        ou, oi, oy = O.sample_pointwise(rng, su[sl].tolist(), si[sl].tolist(), sets, ni)
        assert np.array_equal(np.stack([ou, oi, oy]), g[key])


def test_norm_adjacency_matches_reference(golden_graph, graph_ids):
    g = golden_graph
    u, i, nu, ni = graph_ids
    adj = O.norm_adjacency(nu, ni, u, i)
    assert tuple(g['adj_shape']) == adj.shape
    assert np.array_equal(adj.indptr, g['adj_indptr'])
    assert np.array_equal(adj.indices, g['adj_indices'])
    np.testing.assert_allclose(adj.data, g['adj_data'], rtol=1e-6)
    # C SpMM restatement against scipy on the golden matrix
    X = np.random.default_rng(0).standard_normal((adj.shape[0], 64)).astype(np.float32)
    np.testing.assert_allclose(c_oracle.spmm_csr(adj.indptr, adj.indices, adj.data, X), adj @ X,
                               rtol=1e-4, atol=1e-5)


def test_philox_known_answer():
    # Random123 kat_vectors: philox4x32-10, ctr = key = 0  -> 6627e8d5 e169c58d bc57ac4c 9b00dbd8
    z = np.zeros(1, np.uint32)
    assert int(O.philox4x32_10_x(z, z, z, z, 0, 0)[0]) == 0x6627e8d5
    f = np.full(1, 0xffffffff, np.uint32)          # ctr = key = ff..ff -> 408f276d ...
    assert int(O.philox4x32_10_x(f, f, f, f, 0xffffffff, 0xffffffff)[0]) == 0x408f276d
