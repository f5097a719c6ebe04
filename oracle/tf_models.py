"""Restatements of the reference's TensorFlow-1.14 graphs in torch (CPU, float64, autograd).
TEST INFRASTRUCTURE ONLY.  "Parity unpinned": TensorFlow is absent from this image, so these follow
the cited reference lines and TF1 op semantics instead of recorded TF outputs;
using autograd here makes them an independent check of the engine's hand-derived backward passes.
"""
import numpy as np
import torch

from . import bpr_oracle as O


def philox_uniform(n_rows, d, seed, tag, step):
    """The engine's noise definition (include/qrec.h, qrec_simgcl_perturb_f32): element (r, c) is
    word (c & 3) of Philox4x32-10(key=seed; ctr=(r, c>>2, tag, step)), u = (w >> 8) * 2^-24."""
    nvec = d // 4
    r = np.repeat(np.arange(n_rows, dtype=np.uint32), nvec)
    v = np.tile(np.arange(nvec, dtype=np.uint32), n_rows)
    words = O.philox4x32_10(r, v, np.full(r.shape, tag, np.uint32), np.full(r.shape, step, np.uint32),
                            seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    w = np.stack(words, axis=1).reshape(n_rows, d)
    return (w >> np.uint32(8)).astype(np.float64) * (1.0 / 16777216.0)


def _sp(adj):
    adj = adj.tocoo()
    return torch.sparse_coo_tensor(np.stack([adj.row, adj.col]), adj.data.astype(np.float64), adj.shape).coalesce()


def _bpr_loss(ue, pe, ne, eps):
    score = (ue * pe).sum(1) - (ue * ne).sum(1)                       # util/loss.py:4
    return -torch.log(torch.sigmoid(score) + eps).sum()               # util/loss.py:5


def _l2n(x):
    return x * torch.rsqrt(torch.clamp((x * x).sum(1, keepdim=True), min=1e-12))   # tf.nn.l2_normalize


def simgcl_loss_and_grad(adj, ego, num_users, u, i, j, n_layers, eps, cl_rate, reg, noise, tau=0.2):
    """model/ranking/SimGCL.py:22-38,60-98.  ego: float64 array [N,d]; noise[e][k]: uniform [N,d]
    arrays for perturbed encoder e (0/1), layer k.  Returns (rec, cl, d total / d ego)."""
    A = _sp(adj)
    E0 = torch.tensor(ego, dtype=torch.float64, requires_grad=True)

    def encoder(pert):
        emb, outs = E0, []
        for k in range(n_layers):
            emb = torch.sparse.mm(A, emb)
            if pert is not None:
                nz = torch.tensor(noise[pert][k], dtype=torch.float64)
                emb = emb + torch.sign(emb) * _l2n(nz) * eps          # SimGCL.py:34-35
            outs.append(emb)
        m = torch.stack(outs).mean(0)                                  # SimGCL.py:27,37
        return m[:num_users], m[num_users:]
    mU, mV = encoder(None)
    p1U, p1V = encoder(0)
    p2U, p2V = encoder(1)
    ut, it, jt = (torch.as_tensor(np.asarray(x), dtype=torch.long) for x in (u, i, j))
    ue, pe, ne = mU[ut], mV[it], mV[jt]
    rec = _bpr_loss(ue, pe, ne, 10e-8) + reg * 0.5 * ((ue ** 2).sum() + (pe ** 2).sum() + (ne ** 2).sum())
    cl = 0
    for t1, t2, idx in ((p1U, p2U, torch.unique(ut)), (p1V, p2V, torch.unique(it))):
        z1, z2 = _l2n(t1[idx]), _l2n(t2[idx])
        pos = torch.exp((z1 * z2).sum(1) / tau)
        ttl = torch.exp(z1 @ z2.t() / tau).sum(1)
        cl = cl - torch.log(pos / ttl).sum()
    total = rec + cl_rate * cl
    total.backward()
    return float(rec), float(cl_rate * cl), E0.grad.numpy()


def ngcf_loss_and_grad(adj, ego, W1, W2, num_users, u, i, j, reg, masks=None, keep=0.9):
    """model/ranking/NGCF.py:19-53.  masks[k]: 0/1 dropout keep masks [N,d] (None = inference).
    Returns (loss, grad_ego, [grad_W1_k], [grad_W2_k], final [N,3d] embeddings)."""
    A = _sp(adj)
    E0 = torch.tensor(ego, dtype=torch.float64, requires_grad=True)
    W1t = [torch.tensor(w, dtype=torch.float64, requires_grad=True) for w in W1]
    W2t = [torch.tensor(w, dtype=torch.float64, requires_grad=True) for w in W2]
    e, outs = E0, [E0]
    for k in range(len(W1)):
        side = torch.sparse.mm(A, e)
        z = (side + e) @ W1t[k] + (e * side) @ W2t[k]
        e = torch.nn.functional.leaky_relu(z, 0.2)
        if masks is not None:
            e = e * torch.tensor(masks[k], dtype=torch.float64) / keep
        outs.append(_l2n(e))
    allE = torch.cat(outs, 1)
    Ue, Ve = allE[:num_users], allE[num_users:]
    ut, it, jt = (torch.as_tensor(np.asarray(x), dtype=torch.long) for x in (u, i, j))
    ue, pe, ne = Ue[ut], Ve[it], Ve[jt]
    loss = _bpr_loss(ue, pe, ne, 10e-8) + reg * 0.5 * ((ue ** 2).sum() + (pe ** 2).sum() + (ne ** 2).sum())
    loss.backward()
    return (float(loss), E0.grad.numpy(), [w.grad.numpy() for w in W1t], [w.grad.numpy() for w in W2t],
            allE.detach().numpy())


def neumf_loss_and_grad(params, mode, u, i, r, reg):
    """model/ranking/NeuMF.py:27-75.  params: dict of float64 arrays (PG,QG,PM,QM,h_mf,h_mlp,W1..b3).
    mode 0 = mf_loss, 1 = mlp_loss, 2 = neu_loss.  Returns (loss, {name: grad}, y)."""
    P = {k: torch.tensor(np.asarray(v, dtype=np.float64), requires_grad=True) for k, v in params.items()}
    ut, it = (torch.as_tensor(np.asarray(x), dtype=torch.long) for x in (u, i))
    rt = torch.tensor(np.asarray(r, dtype=np.float64))
    UG, IG = P['PG'][ut], P['QG'][it]
    gmf = UG * IG
    x = torch.cat([P['PM'][ut], P['QM'][it]], 1)
    h = torch.relu(x @ P['W1'] + P['b1'])
    h = torch.relu(h @ P['W2'] + P['b2'])
    mlp = torch.relu(h @ P['W3'] + P['b3'])
    l2 = lambda t: (t ** 2).sum() / 2                                   # noqa: E731  tf.nn.l2_loss
    mf_reg = reg * (l2(UG) + l2(IG) + l2(P['h_mf']))
    e = 10e-10

    def bce(y):
        return -(rt * torch.log(y + e) + (1 - rt) * torch.log(1 - y + e)).sum()
    if mode == 0:
        y = torch.sigmoid((gmf * P['h_mf']).sum(1))
        loss = bce(y) + mf_reg
    elif mode == 1:
        y = torch.sigmoid((mlp * P['h_mlp']).sum(1))
        loss = bce(y)
    else:
        hn = torch.cat([0.5 * P['h_mf'], 0.5 * P['h_mlp']])
        y = torch.sigmoid((torch.cat([gmf, mlp], 1) * hn).sum(1))
        loss = bce(y) + mf_reg + reg * l2(hn)
    loss.backward()
    return float(loss), {k: (v.grad.numpy() if v.grad is not None else None) for k, v in P.items()}, y.detach().numpy()


def sgl_loss_and_grad(adj_main, views, ego, num_users, u, i, j, n_layers, ssl_reg, temp, reg):
    """model/ranking/SGL.py:56-76 (three LightGCN encoders, mean of E_0..E_n), :206-230 (calc_ssl_loss_v3: users and
    items of the batch merged into one InfoNCE), :232-238 (BPR + batch L2 + ssl).  views[v][k]: scipy matrix of
    view v (0/1), layer k.  Returns (rec, ssl, d total / d ego)."""
    E0 = torch.tensor(ego, dtype=torch.float64, requires_grad=True)

    def encoder(mats):
        emb, outs = E0, [E0]
        for k in range(n_layers):
            emb = torch.sparse.mm(_sp(mats[k]), emb)
            outs.append(emb)
        return torch.stack(outs, 1).mean(1)
    m0 = encoder([adj_main] * n_layers)
    m1, m2 = encoder(views[0]), encoder(views[1])
    ut, it, jt = (torch.as_tensor(np.asarray(x), dtype=torch.long) for x in (u, i, j))
    ue, pe, ne = m0[ut], m0[num_users + it], m0[num_users + jt]
    rec = _bpr_loss(ue, pe, ne, 10e-8) + reg * 0.5 * ((ue ** 2).sum() + (pe ** 2).sum() + (ne ** 2).sum())
    idx = torch.cat([torch.unique(ut), torch.unique(it) + num_users])
    z1, z2 = _l2n(m1[idx]), _l2n(m2[idx])
    pos = torch.exp((z1 * z2).sum(1) / temp)
    ttl = torch.exp(z1 @ z2.t() / temp).sum(1)
    ssl = -torch.log(pos / ttl).sum()
    (rec + ssl_reg * ssl).backward()
    return float(rec), float(ssl_reg * ssl), E0.grad.numpy()


def sbpr_loss_and_grad(U, V, u, i, k, j, weights):
    """model/ranking/SBPR.py:103-115 (trainModel_tf): the minibatch loss over the two embedding tables,

        y_ik = (u.i - u.k) / (weights + 1)                               SBPR.py:110-111
        y_kj = u.k - u.j                                                  SBPR.py:112-113
        loss = -sum( ln(sigmoid(y_ik) + 1e-6) + ln(sigmoid(y_kj) + 1e-6) )   SBPR.py:114

    The `+ self.regU * (l2_loss(U) + l2_loss(V))` on SBPR.py:115 is an expression statement of its own (the previous
    line is complete), so the regulariser never reaches `loss`; it is absent here.  Returns (loss, dL/dU, dL/dV) with
    the gradients dense (TF's IndexedSlices of the three lookups, summed per row)."""
    Ut = torch.tensor(np.asarray(U), dtype=torch.float64, requires_grad=True)
    Vt = torch.tensor(np.asarray(V), dtype=torch.float64, requires_grad=True)
    ul, il, kl, jl = (torch.as_tensor(np.asarray(x, dtype=np.int64)) for x in (u, i, k, j))
    wt = torch.as_tensor(np.asarray(weights, dtype=np.float64))
    ue = Ut[ul]
    y_ik = ((ue * Vt[il]).sum(1) - (ue * Vt[kl]).sum(1)) / (wt + 1)
    y_kj = (ue * Vt[kl]).sum(1) - (ue * Vt[jl]).sum(1)
    loss = -(torch.log(torch.sigmoid(y_ik) + 1e-6) + torch.log(torch.sigmoid(y_kj) + 1e-6)).sum()
    loss.backward()
    return float(loss.detach()), Ut.grad.numpy(), Vt.grad.numpy()
