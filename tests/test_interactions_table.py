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


# --------------------------------------------------------------------------------------------- native reader
def _python_path(path, **kw):
    """InteractionTable.from_text with the native reader switched off (an equivalent delimiter regex)."""
    return InteractionTable.from_text(path, delim='[ ,\t]', **kw)


def _same(a, b):
    return (a.user_names.tolist() == b.user_names.tolist() and a.item_names.tolist() == b.item_names.tolist()
            and np.array_equal(a.u, b.u) and np.array_equal(a.i, b.i) and np.array_equal(a.r, b.r))


@pytest.mark.parametrize('kw', [dict(), dict(header=True), dict(binarize_threshold=2.0), dict(columns=(1, 0, 2)),
                                dict(columns=(0, 1)), dict(columns=(0, 2, 3), binarize_threshold=1.0)])
def test_native_reader_equals_python_loop(tmp_path, kw, monkeypatch):
    rng = random.Random(len(repr(kw)))
    lines = []
    for n in range(3000):
        u, i = 'user%d' % rng.randint(0, 200), 'itém%d' % rng.randint(0, 300)        # non-ASCII names
        r = rng.choice(['0.5', '1', '2.0', '3.5', '4', '1e0', '+2', '.5'])
        sep = rng.choice([' ', ',', '\t'])
        lead, trail = rng.choice(['', ' ', '\t']), rng.choice(['', ' ', '\r', ' \t'])
        lines.append(lead + sep.join([u, i, r, '7']) + trail + '\n')
    lines[-1] = lines[-1].rstrip('\n')                                                   # no final newline
    p = tmp_path / 'r.txt'
    p.write_text(''.join(lines), encoding='utf-8')
    calls = []
    orig = InteractionTable._from_text_native.__func__
    monkeypatch.setattr(InteractionTable, '_from_text_native',
                        classmethod(lambda cls, *a: (calls.append(1), orig(cls, *a))[1]))
    a = InteractionTable.from_text(str(p), **kw)
    assert calls == [1]
    b = _python_path(str(p), **kw)
    assert len(calls) == 1 and _same(a, b) and len(a) > 0


def test_native_reader_keeps_empty_fields_like_re_split(tmp_path):
    """re.split(' |,|\\t') does not merge separators: 'a  b 3' is ['a', '', 'b', '3']."""
    p = tmp_path / 'r.txt'
    p.write_text('a  b 3\nc ,d 4\n')
    t = InteractionTable.from_text(str(p), columns=(0, 2, 3))
    assert t.user_names.tolist() == ['a', 'c'] and t.item_names.tolist() == ['b', 'd'] and t.r.tolist() == [3.0, 4.0]
    t = InteractionTable.from_text(str(p), columns=(0, 1, 3))
    assert t.item_names.tolist() == [''] and _same(t, _python_path(str(p), columns=(0, 1, 3)))


def test_native_reader_declines_and_python_path_decides(tmp_path):
    p = tmp_path / 'r.txt'
    p.write_text('a b 1_0\nc d 2\n')                      # float('1_0') == 10.0 in Python, not a strtod number
    t = InteractionTable.from_text(str(p))
    assert t.r.tolist() == [10.0, 2.0]
    p.write_text('a b 1\nshort\n')                         # the reference fails with IndexError on a short line
    with pytest.raises(IndexError):
        InteractionTable.from_text(str(p))
    with pytest.raises(FileNotFoundError):
        InteractionTable.from_text(str(tmp_path / 'absent.txt'))
    p.write_text('a b nan\n')
    assert np.isnan(InteractionTable.from_text(str(p)).r[0])


def test_native_reader_parses_decimals_like_float(tmp_path):
    """The short-decimal fast path (digits / 10^k, both exact) and the strtod path must give float()'s
    bits for every token."""
    rng = random.Random(11)
    toks = ['0', '5', '4.', '.25', '0.1', '0.30000000000000004', '123456789012345', '1234567.89012345', '9.999999999999999',
            '1e-3', '-2.5', '+7', '00012.5000', '179769313486231', '0.000000000000001']
    for _ in range(3000):
        a, b = rng.randrange(0, 10 ** rng.randrange(1, 9)), rng.randrange(0, 10 ** rng.randrange(1, 8))
        toks.append('%d.%0*d' % (a, rng.randrange(1, 8), b))
    p = tmp_path / 'r.txt'
    p.write_text(''.join('u%d i%d %s\n' % (k, k % 7, t) for k, t in enumerate(toks)))
    t = InteractionTable.from_text(str(p))
    assert t.r.tolist() == [float(x) for x in toks]
