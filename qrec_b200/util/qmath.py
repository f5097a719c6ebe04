"""Scalar helpers used around the hot path (reference: util/qmath.py:127-146)."""
import heapq
from math import exp

import numpy as np


def sigmoid(val):
    return 1 / (1 + exp(-val))


def _heap_top_k(K, candidates):
    # the reference's algorithm verbatim in behaviour: min-heap of (score, id), strict '>' replace,
    # then a stable sort by score -- only needed when ties make the order algorithm-dependent.
    heap = [(float(s), k) for k, s in enumerate(candidates[:K])]
    heapq.heapify(heap)
    for k in range(K, len(candidates)):
        s = float(candidates[k])
        if s > heap[0][0]:
            heapq.heapreplace(heap, (s, k))
    heap.sort(key=lambda t: t[0], reverse=True)
    return [t[1] for t in heap], [t[0] for t in heap]


def find_k_largest(K, candidates):
    """ids and scores of the K largest entries, highest first (util/qmath.py:134-146)."""
    scores = np.asarray(candidates, dtype=np.float64)
    n = scores.shape[0]
    if n <= K:
        return _heap_top_k(K, scores)
    part = np.argpartition(-scores, K)[:K + 1]
    top = part[np.argsort(-scores[part], kind='stable')]
    vals = scores[top]
    if np.any(vals[1:] == vals[:-1]):          # ties inside / at the edge of the top K
        return _heap_top_k(K, scores)
    return top[:K].tolist(), vals[:K].tolist()
