"""Vectorised ranking metrics over id arrays -- the same definitions as util/measure.py:24-138
(`Measure.rankingMeasure`), for evaluations too large for dict-of-dict loops:

    Precision = total hits / (users * N)          Recall = mean over users of hits_u / |test_u|
    F1 = 2PR/(P+R)                                NDCG = mean over users of DCG_u / IDCG_u,
    DCG_u = sum over hit ranks k (0-based) of 1/ln(k+2),  IDCG_u = sum_{k < min(N, |test_u|)} 1/ln(k+2)
"""
import numpy as np


def ranking_measures(top_ids, test_rowptr, test_cols, tops):
    """top_ids: int array [n_users, Nmax] (best first) of the users being evaluated, in the order of
    test_rowptr (int64[n_users+1]) / test_cols (their test items, any order).  tops: list of N values.
    Returns the reference's list of strings ('Top N', 'Precision:..', 'Recall:..', 'F1:..', 'NDCG:..')."""
    top_ids = np.asarray(top_ids)
    n_users, nmax = top_ids.shape
    deg = np.diff(test_rowptr).astype(np.int64)
    # membership of every recommended item in its user's test set, via sorted (user, item) keys
    span = int(max(int(test_cols.max()) + 1 if len(test_cols) else 1, int(top_ids.max()) + 1 if top_ids.size else 1))
    owner = np.repeat(np.arange(n_users, dtype=np.int64), deg)
    keys = np.sort(owner * span + test_cols.astype(np.int64))
    rec_keys = (np.arange(n_users, dtype=np.int64)[:, None] * span + top_ids.astype(np.int64)).ravel()
    pos = np.searchsorted(keys, rec_keys)
    hit = np.zeros(rec_keys.shape[0], dtype=bool)
    ok = pos < len(keys)
    hit[ok] = keys[pos[ok]] == rec_keys[ok]
    hit = hit.reshape(n_users, nmax)
    gains = 1.0 / np.log(np.arange(nmax) + 2.0)
    out = []
    for n in tops:
        h = hit[:, :n]
        hits_u = h.sum(1)
        prec = int(hits_u.sum()) / (n_users * n)
        # per-user terms are accumulated left to right like the reference's Python loops
        rec_terms = hits_u / deg
        recall = _seq_sum(rec_terms) / n_users
        f1 = 2 * prec * recall / (prec + recall) if (prec + recall) != 0 else 0
        dcg = np.zeros(n_users)
        for k in range(min(n, nmax)):
            dcg = dcg + np.where(h[:, k], gains[k], 0.0)
        idcg_table = np.concatenate([[0.0], np.cumsum(gains[:n])])          # cumsum is sequential
        idcg = idcg_table[np.minimum(deg, n)]
        ndcg = _seq_sum(dcg / idcg) / n_users
        out += ['Top ' + str(n) + '\n', 'Precision:' + str(prec) + '\n', 'Recall:' + str(recall) + '\n',
                'F1:' + str(f1) + '\n', 'NDCG:' + str(ndcg) + '\n']
    return out


def _seq_sum(x):
    """Left-to-right float sum (what `total += term` does); numpy's pairwise sum differs in the last bits."""
    return float(np.cumsum(np.asarray(x, dtype=np.float64))[-1]) if len(x) else 0.0
