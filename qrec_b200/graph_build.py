"""Normalised joint adjacency on the device.

The reference assembles D^-1/2 (R (+) R^T) D^-1/2 with scipy on the host and then turns it into a
tf.SparseTensor through a Python list of (row, col) pairs (base/graphRecommender.py:10-39) -- minutes
at 100 M non-zeros -- and SGL re-does all of it twice per epoch for its edge-dropout views
(model/ranking/SGL.py:113-155, 233-250).  Here:

  structure   once per data set: the id-mapped interaction lines are sorted into the joint CSR (user rows then
              item rows, ascending columns), with `pair` (entry -> undirected edge) and `line_pair` (interaction
              line -> edge) maps.  One-off data preparation with tensor ops (device-agnostic, CPU-testable).
  values      kernels (csrc/adj_kernels.cu): edge multiplicities from the lines (duplicates SUMMED like scipy's
              constructor, graphRecommender.py:19-20), degrees, D^-1/2 scaling in the reference's float32
              operand order; isolated nodes get d^-1/2 = 0.
  sub-graphs  per epoch and view: Philox keep flags per line (or a caller-supplied mask, e.g. from the MT19937
              clone of random.sample) -> multiplicities -> kept counts + degrees -> scan -> ordered compaction.
              Dropping entries keeps a sorted CSR sorted: no sort, one 8-byte host read (the new nnz).
"""
import torch


class JointAdjacency(object):
    def __init__(self, u_ids, i_ids, num_users, num_items, device=None):
        dev = torch.device(device) if device is not None else u_ids.device
        u = u_ids.to(dev).long()
        it = i_ids.to(dev).long()
        self.device, self.num_users, self.num_items = dev, int(num_users), int(num_items)
        self.n_rows = self.num_users + self.num_items
        self.n_lines = int(u.shape[0])
        key = u * num_items + it
        uniq, inverse = torch.unique(key, sorted=True, return_inverse=True)      # edges sorted by (user, item)
        self.n_pairs = int(uniq.shape[0])
        self.line_pair = inverse.int().contiguous()
        uu = torch.div(uniq, num_items, rounding_mode='floor')
        ii = uniq - uu * num_items
        order = torch.argsort(ii * num_users + uu)                               # the same edges by (item, user)
        rowptr = torch.zeros(self.n_rows + 1, dtype=torch.int64, device=dev)
        rowptr[1:num_users + 1] = torch.cumsum(torch.bincount(uu, minlength=num_users), 0)
        rowptr[num_users + 1:] = rowptr[num_users] + torch.cumsum(torch.bincount(ii, minlength=num_items), 0)
        self.rowptr = rowptr
        self.cols = torch.cat([(ii + num_users).int(), uu[order].int()]).contiguous()
        self.pair = torch.cat([torch.arange(self.n_pairs, device=dev, dtype=torch.int32), order.int()]).contiguous()
        self.pair_w = torch.empty(self.n_pairs, dtype=torch.float32, device=dev)

    def full(self):
        """(rowptr, cols, vals) of the whole graph -- create_joint_sparse_adjaceny's matrix."""
        from . import engine as E
        E.adj_line_weights(self.line_pair, None, self.pair_w)
        deg = torch.empty(self.n_rows, dtype=torch.float32, device=self.device)
        vals = torch.empty(self.cols.shape[0], dtype=torch.float32, device=self.device)
        E.adj_normalize(self.rowptr, self.cols, self.pair, self.pair_w, deg, vals)
        return self.rowptr, self.cols, vals

    def edge_dropout(self, drop_rate, seed, tag, epoch, keep=None):
        """One augmented view (SGL aug_type 1): lines kept with probability 1 - drop_rate (Philox stream
        (seed, tag, epoch)) or by the caller's uint8 mask; returns the re-normalised sub-graph CSR."""
        from . import engine as E
        if keep is None:
            keep = E.edge_keep_philox(self.n_lines, drop_rate, seed, tag, epoch, self.device)
        E.adj_line_weights(self.line_pair, keep, self.pair_w)
        return E.adj_subgraph(self.rowptr, self.cols, self.pair, self.pair_w)


def norm_adjacency_csr(u_ids, i_ids, num_users, num_items, device=None):
    """u_ids, i_ids: integer tensors (any device) of the training pairs.  Returns (rowptr int64[N+1],
    cols int32[nnz] ascending per row, vals fp32[nnz]) on `device` for the (U+I) x (U+I) matrix."""
    return JointAdjacency(u_ids, i_ids, num_users, num_items, device).full()
