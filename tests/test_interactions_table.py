"""f-3: array-backed interaction table + binary cache against the list/dict pipeline (CPU only)."""
import os
import random

import numpy as np
import pytest

from qrec_b200.data.interactions import InteractionTable
from qrec_b200.data.rating import Rating
from qrec_b200.util.config import ModelConf
from qrec_b200.util.io import FileIO


def _write(tmp_path, rng, n=800, header=False, sep=' '):
    lines = ['user item rating\n'] if header else []
    recs = []
    for _ in range(n):
        u, i, r = 'u%d' % rng.randint(0, 60), 'it%d' % rng.randint(0, 90), rng.choice([0.5, 1, 2, 3.5, 4])
        lines.append(sep.join([u, i, str(r)]) + '\n')
        recs.append([u, i, float(r)])
    p = tmp_path / 'ratings.txt'
    p.write_text(''.join(lines))
    return str(p), recs


@pytest.mark.parametrize('binar', [None, 1.0, 3.0])
@pytest.mark.parametrize('sep,header', [(' ', False), (',', True), ('\t', False)])
def test_table_matches_loader_and_rating(tmp_path, binar, sep, header):
    rng = random.Random(hash((binar, sep, header)) & 0xffff)
    path, _ = _write(tmp_path, rng, header=header, sep=sep)
    setup = '-columns 0 1 2' + (' -header' if header else '')
    conf = ModelConf.from_string('ratings.setup=%s\nevaluation.setup=-ap 0.2\n' % setup)
    recs = FileIO.loadDataSet(conf, path, binarized=binar is not None, threshold=binar or 0)
    data = Rating(conf, [r[:] for r in recs], [])
    t = InteractionTable.from_text(path, header=header, binarize_threshold=binar)
    assert t.user_names.tolist() == [data.id2user[k] for k in range(len(data.user))]
    assert t.item_names.tolist() == [data.id2item[k] for k in range(len(data.item))]
    u, i, r = data.training_ids()
    assert np.array_equal(t.u, u) and np.array_equal(t.i, i) and np.array_equal(t.r, r)
    assert t.to_records() == recs
    a, b = t.rated_csr(), data.rated_csr()
    for f in ('pos_rowptr', 'pos_cols', 'sorted_rowptr', 'sorted_cols', 'possorted_cols'):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f


def test_cache_roundtrip_and_invalidation(tmp_path):
    rng = random.Random(1)
    path, _ = _write(tmp_path, rng)
    t1 = InteractionTable.cached_from_text(path, binarize_threshold=1.0)
    cache = [f for f in os.listdir(tmp_path) if f.endswith('.qrec.npz')]
    assert len(cache) == 1
    t2 = InteractionTable.cached_from_text(path, binarize_threshold=1.0)          # served from the cache
    for f in ('u', 'i', 'r', 'user_names', 'item_names'):
        assert np.array_equal(getattr(t1, f), getattr(t2, f))
    t3 = InteractionTable.cached_from_text(path)                                   # other options: other cache entry
    assert len(t3) > len(t1) and len([f for f in os.listdir(tmp_path) if f.endswith('.qrec.npz')]) == 2
    # touching the source invalidates
    os.utime(path, (os.path.getmtime(path) + 10, os.path.getmtime(path) + 10))
    with open(path, 'a') as fh:
        fh.write('zz yy 4\n')
    os.utime(path, None)
    t4 = InteractionTable.cached_from_text(path, binarize_threshold=1.0)
    assert len(t4) == len(t1) + 1 and t4.user_names[-1] == 'zz'


def test_split_equals_seeded_reference_style_split(tmp_path):
    """table.split(mask) with the C MT19937 data_split == DataSplit.dataSplit + Rating on the lists."""
    from qrec_b200 import engine as E
    from qrec_b200.util.dataSplit import DataSplit
    rng = random.Random(2)
    path, recs = _write(tmp_path, rng)
    t = InteractionTable.from_text(path)
    random.seed(0)
    train, test = DataSplit.dataSplit([r[:] for r in recs], test_ratio=0.2)
    state = random.getstate()
    m = E.MT19937(0)
    keep = m.data_split(len(t), 0.2)
    assert m.getstate() == state
    tr, (tu, ti, tr_r) = t.split(keep)
    assert tr.to_records() == train
    assert [[a, b, float(c)] for a, b, c in zip(tu.tolist(), ti.tolist(), tr_r.tolist())] == test


def test_empty_table(tmp_path):
    p = tmp_path / 'empty.txt'
    p.write_text('')
    t = InteractionTable.from_text(str(p))
    assert len(t) == 0 and t.num_users == 0 and t.rated_csr().num_positives == 0
    assert InteractionTable.from_records([]).to_records() == []
