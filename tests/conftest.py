import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a box without a GPU skips the gpu-marked tests instead of failing them."""
    try:
        import torch
        has_cuda = torch.cuda.is_available()
    except Exception:                                    # noqa: BLE001
        has_cuda = False
    if has_cuda:
        # `pytest -x` stops at the first failure: run the suites with the longest hardware record first and the ones
        # added most recently last, so that a regression in new code cannot hide the evidence for the old
        late = ['lifecycle[SGL']
        rank = lambda it: next((k + 1 for k, name in enumerate(late) if name in it.nodeid), 0)   # noqa: E731
        items.sort(key=rank)                         # stable: everything else keeps its order
        return
    skip = pytest.mark.skip(reason='needs a CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session')
def golden_bpr():
    return np.load(os.path.join(GOLDEN, 'bpr_filmtrust_seed0.npz'))


@pytest.fixture(scope='session')
def golden_graph():
    return np.load(os.path.join(GOLDEN, 'sampler_graph_filmtrust_seed1234.npz'))


def ids_from_names(names, vocab):
    lut = {n: k for k, n in enumerate(vocab.tolist())}
    return np.array([lut[n] for n in names.tolist()], dtype=np.int32)


@pytest.fixture(scope='session')
def bpr_ids(golden_bpr):
    """(u_ids, i_ids, num_users, num_items) of the FilmTrust training list in reference order."""
    g = golden_bpr
    u = ids_from_names(g['train_users'], g['user_names'])
    i = ids_from_names(g['train_items'], g['item_names'])
    return u, i, len(g['user_names']), len(g['item_names'])


@pytest.fixture(scope='session')
def graph_ids(golden_graph):
    g = golden_graph
    u = ids_from_names(g['train_users'], g['user_names'])
    i = ids_from_names(g['train_items'], g['item_names'])
    return u, i, len(g['user_names']), len(g['item_names'])


def rows_and_sets(u, i, num_users):
    """dict-of-dict view of the training list: per-user item lists (insertion order, dedup) + sets."""
    rows = [dict() for _ in range(num_users)]
    for uu, ii in zip(u.tolist(), i.tolist()):
        rows[uu][ii] = 1
    return [list(r.keys()) for r in rows], [set(r.keys()) for r in rows]


def adjacency_kernel_stand_ins(monkeypatch):
    """numpy restatements of the two value kernels of csrc/adj_kernels.cu (the GPU suite runs the real ones)."""
    import torch
    from qrec_b200 import engine as E

    def line_weights(line_pair, keep, pair_w):
        k = np.ones(len(line_pair), bool) if keep is None else keep.numpy().astype(bool)
        pair_w.copy_(torch.from_numpy(np.bincount(line_pair.numpy()[k], minlength=len(pair_w)).astype(np.float32)))
        return pair_w

    def normalize(rowptr, cols, pair, pair_w, deg, vals):
        rp, co = rowptr.numpy(), cols.numpy()
        w = pair_w.numpy()[pair.numpy()]
        row = np.repeat(np.arange(len(rp) - 1), np.diff(rp))
        d = np.zeros(len(rp) - 1, np.float32)
        np.add.at(d, row, w)
        with np.errstate(divide='ignore'):
            dinv = np.where(d > 0, (1.0 / np.sqrt(d.astype(np.float64))), 0.0).astype(np.float32)
        deg.copy_(torch.from_numpy(d))
        vals.copy_(torch.from_numpy((dinv[row] * w) * dinv[co]))
        return vals
    monkeypatch.setattr(E, 'adj_line_weights', line_weights)
    monkeypatch.setattr(E, 'adj_normalize', normalize)
