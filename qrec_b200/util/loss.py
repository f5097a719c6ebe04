"""bpr_loss of the reference (util/loss.py:3-6), executed by the engine.

The reference builds a TF graph node; here the same quantity -- and its gradient, which is what
training actually needs -- comes from the fused K3 kernel (qrec_bpr_grad_scatter_f32)."""
import torch

from .. import engine

BPR_EPS = 10e-8      # the literal in util/loss.py:5


def bpr_loss(user_table, item_table, u_idx, pos_idx, neg_idx, reg=0.0, grad_user=None, grad_item=None):
    """-sum ln(sigmoid(u.p - u.n) + 1e-7) [+ reg * batch L2]; if gradient buffers are given the
    gradient w.r.t. the two tables is scatter-added into them.  Returns a 1-element fp64 tensor."""
    loss = torch.zeros(1, dtype=torch.float64, device=user_table.device)
    gU = grad_user if grad_user is not None else torch.zeros_like(user_table)
    gV = grad_item if grad_item is not None else torch.zeros_like(item_table)
    engine.bpr_grad_scatter(user_table, item_table, u_idx, pos_idx, neg_idx, BPR_EPS, reg, gU, gV, loss)
    return loss
