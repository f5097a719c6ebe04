"""Normalised joint adjacency built on the device.

The reference assembles D^-1/2 (R (+) R^T) D^-1/2 with scipy on the host and then turns it into a
tf.SparseTensor through a Python list of (row, col) pairs (base/graphRecommender.py:10-39) -- minutes
at 100 M non-zeros.  Here the id-mapped interaction arrays go to the device once and the CSR comes out
of a sort + run-length pass there (torch ops: this is data preparation, it runs once per model).

Semantics kept from the reference: duplicate (user, item) lines are SUMMED (an entry of 2.0 before
normalisation, graphRecommender.py:19-20), degrees are row sums of the summed matrix, isolated nodes
get d^-1/2 = 0.
"""
import torch


def norm_adjacency_csr(u_ids, i_ids, num_users, num_items, device=None):
    """u_ids, i_ids: integer tensors (any device) of the training pairs.  Returns (rowptr int64[N+1],
    cols int32[nnz] ascending per row, vals fp32[nnz]) on `device` for the (U+I) x (U+I) matrix."""
    dev = torch.device(device) if device is not None else u_ids.device
    u = u_ids.to(dev).long()
    it = i_ids.to(dev).long()
    n = num_users + num_items
    # unique (user, item) pairs with multiplicities
    key = u * num_items + it
    uniq, counts = torch.unique(key, sorted=True, return_counts=True)          # sorted by (user, item)
    uu = torch.div(uniq, num_items, rounding_mode='floor')
    ii = uniq - uu * num_items
    w = counts.to(torch.float64)
    deg_u = torch.zeros(num_users, dtype=torch.float64, device=dev).index_add_(0, uu, w)
    deg_i = torch.zeros(num_items, dtype=torch.float64, device=dev).index_add_(0, ii, w)
    inv_u = torch.where(deg_u > 0, deg_u.pow(-0.5), torch.zeros_like(deg_u))
    inv_i = torch.where(deg_i > 0, deg_i.pow(-0.5), torch.zeros_like(deg_i))
    # the reference scales in float32: (d_u^-1/2 * a) * d_i^-1/2
    vals_ui = ((inv_u[uu].float() * w.float()) * inv_i[ii].float())
    # user rows: (uu, ii) already sorted by user then item
    cnt_u = torch.bincount(uu, minlength=num_users)
    # item rows: sort the same edges by (item, user)
    order = torch.argsort(ii * num_users + uu)
    cnt_i = torch.bincount(ii, minlength=num_items)
    rowptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    rowptr[1:num_users + 1] = torch.cumsum(cnt_u, 0)
    rowptr[num_users + 1:] = rowptr[num_users] + torch.cumsum(cnt_i, 0)
    cols = torch.cat([(ii + num_users).int(), uu[order].int()]).contiguous()
    # transpose entries: (d_i^-1/2 * a) * d_u^-1/2 -- same product, the reference's operand order
    vals_iu = ((inv_i[ii].float() * w.float()) * inv_u[uu].float())[order]
    vals = torch.cat([vals_ui, vals_iu]).contiguous()
    return rowptr, cols, vals
