"""CPU oracle for the rating-prediction MF family -- TEST INFRASTRUCTURE, NOT
PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import it.

Restates, in plain numpy with the reference's own expression order (so float64 results are
bit-identical), the per-entry sequential SGD of
    BasicMF.trainModel   model/rating/BasicMF.py:9-25
    PMF.trainModel       model/rating/PMF.py:9-28
    SVD.trainModel       model/rating/SVD.py:9-36 (+ predictForRating, SVD.py:84-90)
and the epoch bookkeeping they share with BPR (base/iterativeRecommender.py:82-102: loss delta,
adaptive learning rate, `shuffle(trainingData)` after EVERY epoch).

One detail carries the whole parity question: `p = self.P[u]` is a numpy VIEW, so after
`self.P[u] += ...` the item update `self.Q[i] += lr*(error*p - regI*q)` reads the UPDATED user row
(Gauss-Seidel), while `q` -- also a view -- is still the old item row at that point.

Pinned against the reference: tests/test_oracle_mf_golden.py replays tests/golden/mf_*_filmtrust.npz
(made by oracle/gen_golden.py from the unmodified reference): visiting order per epoch, tables after
the last epoch bit-for-bit, epoch losses, learning rates, MT19937 states and the MAE / RMSE lines.
"""
import numpy as np

BASIC, PMF, SVD = 0, 1, 2
KINDS = {'BasicMF': BASIC, 'PMF': PMF, 'SVD': SVD}


def mf_sgd_sequential(kind, P, Q, u, i, r, lr, reg_u=0.0, reg_i=0.0, Bu=None, Bi=None, reg_b=0.0,
                      global_mean=0.0):
    """One pass over the entries (u[k], i[k], r[k]) in order, IN PLACE; returns sum(error^2).
    dtype follows P (float64 = the reference)."""
    T = P.dtype.type
    lr, reg_u, reg_i, reg_b = T(lr), T(reg_u), T(reg_i), T(reg_b)
    gm = T(global_mean)
    loss = 0.0
    for k in range(len(u)):
        uu, ii, rating = int(u[k]), int(i[k]), T(r[k])
        if kind == SVD:
            # SVD.py:88: P[u].dot(Q[i]) + globalMean + Bi[i] + Bu[u]
            error = rating - (P[uu].dot(Q[ii]) + gm + Bi[ii] + Bu[uu])
        else:
            error = rating - P[uu].dot(Q[ii])
        loss += float(error) ** 2
        p, q = P[uu], Q[ii]                         # views
        if kind == BASIC:
            P[uu] += lr * error * q                 # BasicMF.py:22-23: (lr*error)*q
            Q[ii] += lr * error * p
        else:
            if kind == SVD:
                bu, bi = Bu[uu], Bi[ii]             # scalars: copies, unlike p and q
            P[uu] += lr * (error * q - reg_u * p)   # PMF.py:21-22 / SVD.py:27-28
            Q[ii] += lr * (error * p - reg_i * q)
            if kind == SVD:
                Bu[uu] += lr * (error - reg_b * bu)
                Bi[ii] += lr * (error - reg_b * bi)
    return loss


def epoch_loss(kind, sq_err, P, Q, reg_u, reg_i, Bu=None, Bi=None, reg_b=0.0):
    """`self.loss` as the epoch ends: BasicMF has no penalty term (BasicMF.py:12-23), PMF adds
    regU*|P|^2 + regI*|Q|^2 (PMF.py:24), SVD also regB*(|Bu|^2+|Bi|^2) (SVD.py:33-34)."""
    if kind == BASIC:
        return sq_err
    loss = sq_err + (reg_u * (P * P).sum() + reg_i * (Q * Q).sum())
    if kind == SVD:
        loss = sq_err + (reg_u * (P * P).sum() + reg_i * (Q * Q).sum() + reg_b * ((Bu * Bu).sum() + (Bi * Bi).sum()))
    return float(loss)


def predict_rating(kind, P, Q, uu, ii, Bu=None, Bi=None, global_mean=0.0):
    if kind == SVD:
        return P[uu].dot(Q[ii]) + global_mean + Bi[ii] + Bu[uu]
    return P[uu].dot(Q[ii])


def mf_sgd_jacobi(kind, P, Q, u, i, r, lr, reg_u=0.0, reg_i=0.0, Bu=None, Bi=None, reg_b=0.0, global_mean=0.0):
    """Minibatch reading of the same step: every entry reads the PRE-batch rows, applies the update to
    its private copy (the item row still sees the entry's own new user row) and the row deltas are
    summed.  Equals mf_sgd_sequential when no row repeats inside the batch.
    Returns (dP, dQ, dBu, dBi, sum error^2) in float64."""
    dP = np.zeros(P.shape, np.float64); dQ = np.zeros(Q.shape, np.float64)
    dBu = np.zeros(P.shape[0], np.float64); dBi = np.zeros(Q.shape[0], np.float64)
    loss = 0.0
    for k in range(len(u)):
        uu, ii = int(u[k]), int(i[k])
        p = P[uu].astype(np.float64); q = Q[ii].astype(np.float64)
        pred = p.dot(q)
        if kind == SVD:
            pred = pred + global_mean + float(Bi[ii]) + float(Bu[uu])
        e = float(r[k]) - pred
        loss += e * e
        if kind == BASIC:
            pn = p + lr * e * q
            qn = q + lr * e * pn
        else:
            pn = p + lr * (e * q - reg_u * p)
            qn = q + lr * (e * pn - reg_i * q)
        dP[uu] += pn - p
        dQ[ii] += qn - q
        if kind == SVD:
            dBu[uu] += lr * (e - reg_b * float(Bu[uu]))
            dBi[ii] += lr * (e - reg_b * float(Bi[ii]))
    return dP, dQ, dBu, dBi, loss
