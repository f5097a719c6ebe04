"""CPU oracle for the QRec hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product (qrec_b200/) never does; it fails loudly without its CUDA
library instead of falling back here.

Each function restates one piece of the reference (Coder-Yu/QRec, file:line cited) in plain
numpy / pure Python.  Pinning: tests/test_oracle_golden.py checks these restatements against the
golden vectors in tests/golden/, which were produced by running the UNMODIFIED reference
(oracle/gen_golden.py) -- (u,i,j) streams, P/Q after epochs 1 and 3, epoch losses, the
learning-rate schedule, sampler batches and the normalised adjacency.

Pinned against the reference: sample_bpr_epoch, sample_pairwise, sample_pointwise,
bpr_sgd_sequential, epoch_loss_reg, update_learning_rate, norm_adjacency.
"Parity unpinned" (TensorFlow 1.14 is absent, the reference's TF graphs cannot run here; these
follow the cited lines + TF1 op semantics): lightgcn_forward,
lightgcn_step, bpr_loss_grad, adam_tf1.
"""
import math
import random

import numpy as np


# ---------------------------------------------------------------------------------------------
# K0: samplers.  CPython's own `random.Random` IS the reference RNG, so the oracle simply uses it.
# ---------------------------------------------------------------------------------------------
def make_rng(state625=None, seed=None):
    r = random.Random()
    if seed is not None:
        r.seed(seed)
    if state625 is not None:
        r.setstate((3, tuple(int(x) for x in state625), None))
    return r


def rng_state(r):
    return np.array(r.getstate()[1], dtype=np.uint32)


def sample_bpr_epoch(rng, pos_rows, rated_sets, num_items):
    """model/ranking/BPR.py:28-38.  pos_rows: list over users of item-id lists (insertion order);
    rated_sets: list of sets.  `choice(itemList)` == item id `_randbelow(num_items)` because
    itemList = list(self.data.item.keys()) is in id order (data/rating.py:48-54)."""
    item_list = list(range(num_items))
    out = []
    for u, items in enumerate(pos_rows):
        for i in items:
            j = rng.choice(item_list)
            while j in rated_sets[u]:
                j = rng.choice(item_list)
            out.append((u, i, j))
    return np.array(out, dtype=np.int32).reshape(-1, 3)


def sample_pairwise(rng, users, rated_sets, num_items):
    """base/deepRecommender.py:44-50 for one batch of user ids."""
    item_list = list(range(num_items))
    js = []
    for u in users:
        j = rng.choice(item_list)
        while j in rated_sets[u]:
            j = rng.choice(item_list)
        js.append(j)
    return np.array(js, dtype=np.int32)


def sample_pointwise(rng, users, items, rated_sets, num_items):
    """base/deepRecommender.py:65-76."""
    ou, oi, oy = [], [], []
    for u, i in zip(users, items):
        ou.append(u); oi.append(i); oy.append(1)
        for _ in range(4):
            j = rng.randint(0, num_items - 1)
            while j in rated_sets[u]:
                j = rng.randint(0, num_items - 1)
            ou.append(u); oi.append(j); oy.append(0)
    return (np.array(ou, np.int32), np.array(oi, np.int32), np.array(oy, np.int32))


def shuffle_pairs(rng, a, b):
    """random.shuffle of a list of (a,b) rows (base/deepRecommender.py:30)."""
    rows = list(zip(a.tolist(), b.tolist()))
    rng.shuffle(rows)
    arr = np.array(rows, dtype=np.int32).reshape(-1, 2)
    return arr[:, 0].copy(), arr[:, 1].copy()



# ---------------------------------------------------------------------------------------------
# per-user item sets (data/rating.py:48-55 dict-of-dicts semantics; model/ranking/BPR.py:22-25)
# ---------------------------------------------------------------------------------------------
def rated_csr_numpy(num_users, num_items, u_ids, i_ids, ratings=None, positive_threshold=1.0):
    """The numpy construction the engine used before qrec_build_rated_csr (three full-length sorts);
    kept as the checker of the native builder.  A repeated (user, item) keeps the position of its first
    occurrence and the value of its last.  Returns a dict of the five arrays."""
    u_ids = np.ascontiguousarray(u_ids, dtype=np.int64)
    i_ids = np.ascontiguousarray(i_ids, dtype=np.int64)
    n = u_ids.shape[0]
    if ratings is None:
        ratings = np.ones(n, dtype=np.float64)
    ratings = np.asarray(ratings, dtype=np.float64)
    key = u_ids * int(num_items) + i_ids
    order = np.argsort(key, kind='stable')
    ks = key[order]
    first = np.ones(n, dtype=bool)
    first[1:] = ks[1:] != ks[:-1]
    last = np.ones(n, dtype=bool)
    last[:-1] = ks[1:] != ks[:-1]
    first_pos = order[first]
    last_rating = ratings[order[last]]
    uniq_u, uniq_i = u_ids[first_pos], i_ids[first_pos]
    sorted_rowptr = np.zeros(num_users + 1, dtype=np.int64)
    sorted_rowptr[1:] = np.cumsum(np.bincount(uniq_u, minlength=num_users))
    keep = last_rating >= positive_threshold
    pu, pi, ppos = uniq_u[keep], uniq_i[keep], first_pos[keep]
    o2 = np.lexsort((ppos, pu))
    pos_rowptr = np.zeros(num_users + 1, dtype=np.int64)
    pos_rowptr[1:] = np.cumsum(np.bincount(pu, minlength=num_users))
    return dict(sorted_rowptr=sorted_rowptr, sorted_cols=uniq_i.astype(np.int32), pos_rowptr=pos_rowptr,
                pos_cols=pi[o2].astype(np.int32), possorted_cols=pi.astype(np.int32))

# ---------------------------------------------------------------------------------------------
# K1: BPR.optimization (model/ranking/BPR.py:45-53) and the epoch bookkeeping around it
# ---------------------------------------------------------------------------------------------
def bpr_sgd_sequential(P, Q, triples, lr, reg_u, reg_i):
    """Applies BPR.optimization to P,Q IN PLACE for each (u,i,j) in order; returns sum(-ln s).
    dtype follows P (float64 = the reference; float32 = what an fp32 engine should produce)."""
    T = P.dtype.type
    lr, reg_u, reg_i = T(lr), T(reg_u), T(reg_i)
    loss = 0.0
    one = T(1)
    for u, i, j in triples:
        x = P[u].dot(Q[i]) - P[u].dot(Q[j])
        s = one / (one + T(math.exp(-x)))                 # util/qmath.py:127-128
        g = lr * (one - s)
        P[u] += g * (Q[i] - Q[j])
        Q[i] += g * P[u]
        Q[j] -= g * P[u]
        P[u] -= lr * reg_u * P[u]
        Q[i] -= lr * reg_i * Q[i]
        Q[j] -= lr * reg_i * Q[j]
        loss += -math.log(s)
    return loss


def bpr_sgd_jacobi(P, Q, triples, lr, reg_u, reg_i):
    """Minibatch reading of the same step: every triple reads the PRE-batch rows, applies
    BPR.py:45-52 to its private copy, and the row deltas are summed.  Equals
    bpr_sgd_sequential when no row is shared inside the batch.  Returns (dP, dQ, loss)."""
    dP = np.zeros(P.shape, dtype=np.float64)
    dQ = np.zeros(Q.shape, dtype=np.float64)
    loss = 0.0
    for u, i, j in triples:
        p = P[u].astype(np.float64); qi = Q[i].astype(np.float64); qj = Q[j].astype(np.float64)
        x = p.dot(qi) - p.dot(qj)
        s = 1.0 / (1.0 + math.exp(-x))
        g = lr * (1.0 - s)
        pn = p + g * (qi - qj)
        qin = qi + g * pn
        qjn = qj - g * pn
        pn = pn - lr * reg_u * pn
        qin = qin - lr * reg_i * qin
        qjn = qjn - lr * reg_i * qjn
        dP[u] += pn - p
        dQ[i] += qin - qi
        dQ[j] += qjn - qj
        loss += -math.log(s)
    return dP, dQ, loss


def epoch_loss_reg(P, Q, reg_u, reg_i):
    """model/ranking/BPR.py:40."""
    return reg_u * float((P * P).sum()) + reg_i * float((Q * Q).sum())


def update_learning_rate(lr, max_lr, epoch, last_loss, loss):
    """base/iterativeRecommender.py:56-63 (called only when not converged, :97-99)."""
    if epoch > 1:
        if abs(last_loss) > abs(loss):
            lr *= 1.05
        else:
            lr *= 0.5
    if lr > max_lr > 0:
        lr = max_lr
    return lr


# ---------------------------------------------------------------------------------------------
# K0 fast: Philox4x32-10 (Salmon et al., SC'11 -- the published algorithm; Random123 constants)
# ---------------------------------------------------------------------------------------------
_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = 0x9E3779B9, 0xBB67AE85


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over uint32 arrays c0..c3; scalar keys.  Returns the four output words."""
    c0 = c0.astype(np.uint64); c1 = c1.astype(np.uint64)
    c2 = c2.astype(np.uint64); c3 = c3.astype(np.uint64)
    mask = np.uint64(0xFFFFFFFF)
    k0 = int(k0) & 0xFFFFFFFF; k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & mask
        hi1, lo1 = p1 >> np.uint64(32), p1 & mask
        n0 = hi1 ^ c1 ^ np.uint64(k0)
        n2 = hi0 ^ c3 ^ np.uint64(k1)
        c0, c1, c2, c3 = n0, lo1, n2, lo0
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def philox4x32_10_x(c0, c1, c2, c3, k0, k1):
    """First output word only (the negative sampler uses just this one)."""
    return philox4x32_10(c0, c1, c2, c3, k0, k1)[0]


def sample_neg_philox(users, rated_sets, num_items, seed, epoch):
    """Restates qrec_sample_neg_philox (engine-defined sampler; no reference counterpart beyond
    the rejection rule of base/deepRecommender.py:47-49)."""
    n = len(users)
    k = np.arange(n, dtype=np.uint64)
    out = np.empty(n, dtype=np.int32)
    pending = np.arange(n)
    attempt = 0
    while pending.size:
        kk = k[pending]
        r = philox4x32_10_x((kk & np.uint64(0xFFFFFFFF)).astype(np.uint32),
                            (kk >> np.uint64(32)).astype(np.uint32),
                            np.full(pending.size, attempt, np.uint32),
                            np.full(pending.size, epoch, np.uint32),
                            seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
        cand = ((r.astype(np.uint64) * np.uint64(num_items)) >> np.uint64(32)).astype(np.int32)
        rej = np.array([c in rated_sets[users[p]] for p, c in zip(pending, cand)], dtype=bool)
        out[pending[~rej]] = cand[~rej]
        pending = pending[rej]
        attempt += 1
    return out


# ---------------------------------------------------------------------------------------------
# K2..K4: graph path (TF restatements -- parity unpinned, see module docstring)
# ---------------------------------------------------------------------------------------------
def norm_adjacency(num_users, num_items, u_ids, i_ids):
    """base/graphRecommender.py:10-29: D^-1/2 (R (+) R^T) D^-1/2 as scipy CSR float32, sorted."""
    import scipy.sparse as sp
    n = num_users + num_items
    ones = np.ones(len(u_ids), dtype=np.float32)
    tmp = sp.csr_matrix((ones, (np.asarray(u_ids), np.asarray(i_ids) + num_users)), shape=(n, n))
    adj = tmp + tmp.T
    rowsum = np.array(adj.sum(1))
    with np.errstate(divide='ignore'):
        d_inv = np.power(rowsum, -0.5).flatten()
    d_inv[np.isinf(d_inv)] = 0.
    dm = sp.diags(d_inv)
    out = dm.dot(adj).dot(dm).tocsr()
    out.sort_indices()
    return out.astype(np.float32)


def lightgcn_forward(adj, U, V, n_layers):
    """model/ranking/LightGCN.py:13-20: mean over {E0..En} of E_{k+1} = A E_k (fp32)."""
    E = np.concatenate([U, V], axis=0).astype(np.float32)
    acc = E.copy()
    layers = [E]
    for _ in range(n_layers):
        E = (adj @ E).astype(np.float32)
        layers.append(E)
        acc += E
    out = acc / np.float32(n_layers + 1)
    return out[:U.shape[0]], out[U.shape[0]:], layers


def bpr_loss_grad(Ue, Ve, u, i, j, eps, reg):
    """util/loss.py:3-6 + batch L2 (LightGCN.py:28-30) and the gradient w.r.t. the propagated
    tables, duplicates summed (tf.nn.embedding_lookup's IndexedSlices gradient)."""
    pu = Ue[u].astype(np.float64); pi = Ve[i].astype(np.float64); pj = Ve[j].astype(np.float64)
    y = (pu * pi).sum(1) - (pu * pj).sum(1)
    s = 1.0 / (1.0 + np.exp(-y))
    loss = -np.log(s + eps).sum() + reg * 0.5 * ((pu * pu).sum() + (pi * pi).sum() + (pj * pj).sum())
    gy = -(s * (1.0 - s) / (s + eps))[:, None]
    gU = np.zeros(Ue.shape, np.float64); gV = np.zeros(Ve.shape, np.float64)
    np.add.at(gU, u, gy * (pi - pj) + reg * pu)
    np.add.at(gV, i, gy * pu + reg * pi)
    np.add.at(gV, j, -gy * pu + reg * pj)
    return float(loss), gU, gV


def adam_tf1(var, m, v, g, lr, t, b1=0.9, b2=0.999, eps=1e-8):
    """tf.train.AdamOptimizer dense update as TF 1.14's ApplyAdam functor evaluates it
    (tensorflow/core/kernels/training_ops.cc, un-vendored third party), all fp32:
        alpha = lr*sqrt(1-b2^t)/(1-b1^t);  m += (g-m)*(1-b1);  v += (g*g-v)*(1-b2);
        var -= (m*alpha)/(sqrt(v)+eps)
    In place."""
    f = np.float32
    b1, b2 = f(b1), f(b2)
    alpha = f(lr) * np.sqrt(f(1) - f(float(b2) ** t)) / (f(1) - f(float(b1) ** t))
    m += (g - m) * (f(1) - b1)
    v += (g * g - v) * (f(1) - b2)
    var -= (m * alpha) / (np.sqrt(v) + f(eps))
    return var


def lightgcn_step(adj, U, V, mU, vU, mV, vV, u, i, j, n_layers, lr, reg, t, eps=1e-7):
    """One minibatch of model/ranking/LightGCN.py:27-39: forward, loss, backward through the
    layer mean and the n SpMMs (A symmetric => A^T = A), dense Adam on U and V.  In place."""
    nu = U.shape[0]
    Ue, Ve, _ = lightgcn_forward(adj, U, V, n_layers)
    loss, gUe, gVe = bpr_loss_grad(Ue, Ve, u, i, j, eps, reg)
    G = np.concatenate([gUe, gVe], axis=0).astype(np.float32) / np.float32(n_layers + 1)
    # d(mean)/dE0 = sum_k A^k G / (n+1)
    total = G.copy()
    cur = G
    for _ in range(n_layers):
        cur = (adj @ cur).astype(np.float32)
        total += cur
    adam_tf1(U, mU, vU, total[:nu], lr, t)
    adam_tf1(V, mV, vV, total[nu:], lr, t)
    return loss
