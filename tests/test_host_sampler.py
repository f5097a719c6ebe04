"""K0 compat: the C MT19937 clone in libqrec.so against CPython's `random` (the reference's RNG)
and against the sampler streams recorded from the reference.  CPU only (host code)."""
import random

import numpy as np
import pytest

from qrec_b200 import engine as E
from conftest import rows_and_sets


@pytest.mark.parametrize('seed', [0, 1, 1234, 2**31 - 1, 2**32 + 5, 2**63 + 11])
def test_seed_and_raw_stream(seed):
    r = random.Random(seed)
    m = E.MT19937(seed)
    assert m.getstate() == r.getstate()
    assert [m.getrandbits32() for _ in range(2000)] == [r.getrandbits(32) for _ in range(2000)]
    assert [m.random() for _ in range(500)] == [r.random() for _ in range(500)]
    assert m.getstate() == r.getstate()


def test_randbelow_choice_randint():
    r = random.Random(7)
    m = E.MT19937(7)
    for n in [1, 2, 3, 5, 7, 8, 1891, 2044, 65536, 100000, 2**31 - 1, 2**32 - 1]:
        assert [m.randbelow(n) for _ in range(300)] == [r._randbelow(n) for _ in range(300)]
    seq = list(range(1891))
    assert [m.randbelow(1891) for _ in range(1000)] == [r.choice(seq) for _ in range(1000)]
    assert [m.randbelow(50) for _ in range(1000)] == [r.randint(0, 49) for _ in range(1000)]


@pytest.mark.parametrize('n', [0, 1, 2, 3, 1000, 34437])
def test_shuffle(n):
    r = random.Random(99)
    m = E.MT19937(99)
    x = list(range(n))
    r.shuffle(x)
    a = np.arange(n, dtype=np.int32)
    m.shuffle(a)
    assert a.tolist() == x
    assert m.getstate() == r.getstate()


def test_data_split_matches_python():
    r = random.Random(0)
    m = E.MT19937(0)
    keep = m.data_split(5000, 0.2)
    ref = np.array([not (r.random() < 0.2) for _ in range(5000)])
    assert np.array_equal(keep, ref)
    # out-of-range ratio is reset to 0.3 (util/dataSplit.py:10-11)
    r2, m2 = random.Random(3), E.MT19937(3)
    assert np.array_equal(m2.data_split(100, 1.5), np.array([not (r2.random() < 0.3) for _ in range(100)]))


def test_bpr_epoch_stream_golden(golden_bpr, bpr_ids):
    """Bit-exact (u,i,j) for 3 epochs incl. the end-of-epoch shuffle, from the reference's state."""
    u, i, nu, ni = bpr_ids
    csr = E.RatedCSR(nu, ni, u, i)
    assert csr.num_positives == golden_bpr['triples_epoch'].shape[1]
    m = E.MT19937()
    m.setstate(golden_bpr['mt_state_after_split'])
    su, si = u.copy(), i.copy()
    for ep in range(3):
        tu, ti, tj = m.sample_bpr_epoch(csr)
        assert np.array_equal(np.stack([tu, ti, tj], 1), golden_bpr['triples_epoch'][ep])
        m.shuffle_pairs(su, si)
        assert np.array_equal(m.getstate_array(), golden_bpr['mt_state_after_epoch'][ep])


def test_rated_csr_dict_semantics():
    # duplicates collapse; first-insertion order; rating threshold uses the LAST value
    u = np.array([0, 0, 1, 0, 1, 0])
    i = np.array([3, 1, 2, 3, 0, 2])
    r = np.array([1, 1, 1, 0.5, 1, 1.0])
    c = E.RatedCSR(2, 4, u, i, r)
    assert c.sorted_rowptr.tolist() == [0, 3, 5] and c.sorted_cols.tolist() == [1, 2, 3, 0, 2]
    # (0,3) was overwritten with 0.5 -> not positive; order of the rest = first appearance
    assert c.pos_rowptr.tolist() == [0, 2, 4] and c.pos_cols.tolist() == [1, 2, 2, 0]
    rows, sets = rows_and_sets(u, i, 2)
    assert rows[0] == [3, 1, 2] and sets[1] == {0, 2}


def test_pairwise_and_pointwise_golden(golden_graph, graph_ids):
    g = golden_graph
    u, i, nu, ni = graph_ids
    csr = E.RatedCSR(nu, ni, u, i)
    m = E.MT19937()
    m.setstate(g['mt_state_before_pairwise'])
    su, si = u.copy(), i.copy()
    m.shuffle_pairs(su, si)
    assert np.array_equal(su, g['shuffled_u']) and np.array_equal(si, g['shuffled_i'])
    js = [m.sample_pairwise(csr, su[b:b + 2048]) for b in range(0, len(su), 2048)]
    assert len(js) == int(g['pair_num_batches']) and len(js[-1]) == g['pair_last'].shape[1]
    assert np.array_equal(np.concatenate(js), g['pair_all_j'])
    assert np.array_equal(m.getstate_array(), g['mt_state_after_pairwise'])
    m.setstate(g['mt_state_before_pointwise'])
    for b, key in ((0, 'point_b0'), (1, 'point_b1')):
        sl = slice(b * 2048, (b + 1) * 2048)
        ou, oi, oy = m.sample_pointwise(csr, su[sl], si[sl])
        assert np.array_equal(np.stack([ou, oi, oy]), g[key])
    assert np.array_equal(m.getstate_array(), g['mt_state_after_pointwise'])


def test_sampler_edge_cases():
    m = E.MT19937(5)
    empty = E.RatedCSR(3, 10, np.array([], np.int64), np.array([], np.int64))
    tu, ti, tj = m.sample_bpr_epoch(empty)
    assert len(tu) == 0
    assert len(m.sample_pairwise(empty, np.array([], np.int32))) == 0
    # a user that rated every item cannot be given a negative: error, not an endless loop
    full = E.RatedCSR(1, 3, np.array([0, 0, 0]), np.array([0, 1, 2]))
    with pytest.raises(E.QRecError):
        m.sample_bpr_epoch(full)
    with pytest.raises(E.QRecError):
        m.sample_pairwise(full, np.array([0], np.int32))
    # ragged: users without interactions are skipped
    rag = E.RatedCSR(4, 6, np.array([1, 3, 3]), np.array([2, 0, 5]))
    tu, ti, tj = m.sample_bpr_epoch(rag)
    assert tu.tolist() == [1, 3, 3] and ti.tolist() == [2, 0, 5]
    assert all(j not in s for j, s in zip(tj.tolist(), [{2}, {0, 5}, {0, 5}]))


def test_order_prepare():
    u = np.array([0, 0, 1, 0], np.int32)
    i = np.array([1, 2, 1, 3], np.int32)
    j = np.array([2, 3, 0, 1], np.int32)
    wu, wi, wj = E.bpr_order_prepare(u, i, j, 2, 4)
    assert wu.tolist() == [0, 1, 0, 2]
    assert wi.tolist() == [0, 1, 1, 1]      # Q[1]:0, Q[2]: touched by k=0 as j -> 1, Q[1] again ->1, Q[3]: 1
    assert wj.tolist() == [0, 0, 0, 2]
    with pytest.raises(E.QRecError):
        E.bpr_order_prepare(np.array([0], np.int32), np.array([1], np.int32), np.array([1], np.int32), 1, 2)
    with pytest.raises(E.QRecError):
        E.bpr_order_prepare(np.array([5], np.int32), np.array([1], np.int32), np.array([0], np.int32), 1, 2)


def test_bpr_epoch_with_non_binarised_ratings_uses_positive_set_only():
    """BPR.trainModel builds PositiveSet from ratings >= 1 (model/ranking/BPR.py:21-25), iterates it and
    rejects negatives against IT (not against every rated item).  Restated inline with Python's
    `random`; the C sampler must give the same triples and leave the same generator state."""
    from collections import defaultdict
    from qrec_b200.data.rating import Rating
    from qrec_b200.util.config import ModelConf
    rng = random.Random(3)
    train = [['u%d' % rng.randint(0, 30), 'i%d' % rng.randint(0, 40), float(rng.choice([0.5, 1, 2, 3]))] for _ in range(500)]
    d = Rating(ModelConf.from_string('ratings=x\nevaluation.setup=-ap 0.2\n'), [r[:] for r in train], [])
    positive = defaultdict(dict)
    for user in d.user:
        for item in d.trainSet_u[user]:
            if d.trainSet_u[user][item] >= 1:
                positive[user][item] = 1
    item_list = list(d.item.keys())
    random.seed(5)
    ref = []
    for user in positive:
        for item in positive[user]:
            neg = random.choice(item_list)
            while neg in positive[user]:
                neg = random.choice(item_list)
            ref.append((d.user[user], d.item[item], d.item[neg]))
    state = random.getstate()
    random.seed(5)
    m = E.MT19937()
    m.setstate(random.getstate())
    u, i, j = m.sample_bpr_epoch(d.rated_csr())
    assert np.array_equal(np.stack([u, i, j], 1), np.array(ref)) and m.getstate() == state
    assert any(v < 1 for row in d.trainSet_u.values() for v in row.values())      # the filter mattered


def test_sbpr_and_tbpr_native_samplers_equal_python_random_on_random_structures():
    """qrec_sample_sbpr_batch / qrec_sample_tbpr_epoch against the same loops written with random.Random (the calls
    model/ranking/SBPR.py:84-100 and TBPR.py:131-160 make), on small random structures that hit the corners: users
    without social feedback, one-element pools, missing chain levels, users without positives."""
    import random
    from qrec_b200 import engine as E
    rng = np.random.default_rng(12)
    for trial in range(25):
        nu, ni = int(rng.integers(3, 12)), int(rng.integers(8, 40))
        rated = [sorted(rng.choice(ni, int(rng.integers(0, min(5, ni - 3))), replace=False).tolist()) for _ in range(nu)]
        u_ids = np.array([u for u in range(nu) for _ in rated[u]], np.int32)
        i_ids = np.array([x for u in range(nu) for x in rated[u]], np.int32)
        csr = E.RatedCSR(nu, ni, u_ids, i_ids)

        def pool(p_empty):
            out = []
            for u in range(nu):
                free = [x for x in range(ni) if x not in rated[u]]
                k = 0 if rng.random() < p_empty else int(rng.integers(1, max(2, min(4, len(free) - 2))))
                out.append([int(x) for x in rng.permutation(free)[:k]])
            return out

        def as_csr(lists):
            rp = np.zeros(nu + 1, np.int64)
            rp[1:] = np.cumsum([len(x) for x in lists])
            return rp, np.array([x for row in lists for x in row], np.int32)
        # ---- SBPR rows
        fp = pool(0.4)
        counts = [[int(rng.integers(1, 4)) for _ in row] for row in fp]
        fp_rp, fp_items = as_csr(fp)
        fp_counts = np.array([c for row in counts for c in row], np.int32)
        fp_sorted = np.array([x for row in fp for x in sorted(row)], np.int32)
        rows = rng.integers(0, nu, 60).astype(np.int32)
        seed = int(rng.integers(0, 2 ** 31))
        mt = E.MT19937(seed)
        k, j, w = mt.sample_sbpr_batch(csr, fp_rp, fp_items, fp_counts, fp_sorted, rows)
        ref = random.Random(seed)
        item_list = list(range(ni))
        for r, u in enumerate(rows.tolist()):
            if len(fp[u]) == 0:
                f, wt = ref.choice(item_list), 0
            else:
                f = ref.choice(fp[u]); wt = counts[u][fp[u].index(f)]
            neg = ref.choice(item_list)
            while neg in rated[u] or neg in fp[u]:
                neg = ref.choice(item_list)
            assert (int(k[r]), int(j[r]), int(w[r])) == (f, neg, wt), (trial, r)
        assert mt.getstate() == ref.getstate()
        # ---- TBPR chains (positives = rated: every rating is 1)
        joint, weak, strong = pool(0.5), pool(0.5), pool(0.5)
        order = np.array([u for u in range(nu) if rated[u]], np.int32)
        mt = E.MT19937(seed + 1)
        su, sa, sb, per_user = mt.sample_tbpr_epoch(csr, order, as_csr(joint), as_csr(weak), as_csr(strong))
        ref = random.Random(seed + 1)
        eu, ea, eb, ecount = [], [], [], []
        pos_order = {u: csr.pos_cols[csr.pos_rowptr[u]:csr.pos_rowptr[u + 1]].tolist() for u in range(nu)}
        for u in order.tolist():
            start = len(eu)
            for item in pos_order[u]:
                chain = [item]
                for level in (joint[u], weak[u], strong[u]):
                    if level:
                        chain.append(ref.choice(level))
                neg = ref.choice(item_list)
                while neg in rated[u]:
                    neg = ref.choice(item_list)
                chain.append(neg)
                for a, b in zip(chain[:-1], chain[1:]):
                    eu.append(u); ea.append(a); eb.append(b)
            ecount.append(len(eu) - start)
        assert su.tolist() == eu and sa.tolist() == ea and sb.tolist() == eb and per_user.tolist() == ecount, trial
        assert mt.getstate() == ref.getstate()
