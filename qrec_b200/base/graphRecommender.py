"""`GraphRecommender`: the adjacency builders of the graph models
(reference: base/graphRecommender.py:10-61).

`create_joint_sparse_adjaceny` returns the same scipy CSR (float32, D^-1/2 (R (+) R^T) D^-1/2 of
the (U+I)x(U+I) bipartite graph, duplicate interactions summed before normalisation);
`create_joint_sparse_adj_tensor` hands it to the device as a `DeviceCSR` (int64 rowptr, int32
cols, fp32 vals) -- the operand of the K2 SpMM kernel -- instead of building a tf.SparseTensor
from a Python list of (row, col) pairs.
"""
import numpy as np
import scipy.sparse as sp

from .deepRecommender import DeepRecommender


class DeviceCSR(object):
    """CSR operand resident in HBM; `matmul` is qrec_spmm_csr_f32."""

    def __init__(self, mat, device):
        import torch
        mat = mat.tocsr()
        mat.sort_indices()
        self._finish(mat.shape, torch.from_numpy(mat.indptr.astype(np.int64)).to(device),
                     torch.from_numpy(mat.indices.astype(np.int32)).to(device),
                     torch.from_numpy(mat.data.astype(np.float32)).to(device))
        self.split_row = None

    @classmethod
    def from_tensors(cls, shape, rowptr, cols, vals, split_row=None):
        """Wraps CSR arrays that already live on the device (graph_build.norm_adjacency_csr).  `split_row`: see
        set_split_row."""
        self = cls.__new__(cls)
        self._finish(tuple(shape), rowptr, cols, vals)
        self.set_split_row(split_row)
        return self

    def set_split_row(self, row):
        """The joint adjacency of a bipartite graph is two very different halves: `row` = num_users short user rows that
        gather item rows (a table that lives in the L2) and long item rows that gather user rows.  One launch over all of
        them measured 2.03-2.11 ms on the 1M x 100K x 50M-edge benchmark graph, the two halves launched one after the
        other 0.78 + 0.92 ms (profiles/r2/s2/bench_spmm_blocks.jsonl): lane groups working on 50-entry and on 500-entry
        rows at the same time keep neither table's rows in the L2.  With a split row `matmul` issues the row-split
        kernel once per half; rows, arithmetic and results are those of the single launch."""
        import os
        n = self.shape[0]
        self.split_row = int(row) if (row is not None and 0 < int(row) < n and os.environ.get('QREC_SPMM_SPLIT', '1') != '0') else None

    def _finish(self, shape, rowptr, cols, vals):
        self.shape = shape
        self.rowptr, self.cols, self.vals = rowptr, cols, vals
        self.nnz = int(cols.shape[0])
        # short, even rows: one lane group per row is fastest (no atomics); a long-tailed degree
        # distribution needs the nnz-balanced kernel or the hot rows serialise the launch
        lengths = rowptr[1:] - rowptr[:-1]
        self.rowsplit = bool(lengths.numel() == 0 or int(lengths.max().item()) <= 4096)

    def matmul_sparse_rows(self, X, src_rows, out, acc=None, acc_scale=0.0):
        """out = A @ X when only the rows `src_rows` of X are non-zero (A symmetric)."""
        from .. import engine
        return engine.spmm_csr_scatter_rows(self.rowptr, self.cols, self.vals, src_rows, X, out, acc=acc, acc_scale=acc_scale)

    def matmul_rows(self, X, rows, out=None, compact=False, acc=None, acc_scale=0.0):
        """The rows `rows` (int32, distinct, -1 = padding) of A @ X, stored into `out` and / or accumulated into acc."""
        from .. import engine
        return engine.spmm_csr_rows(self.rowptr, self.cols, self.vals, rows, X, out, compact=compact, acc=acc, acc_scale=acc_scale)

    def matmul(self, X, out, acc=None, acc_scale=0.0):
        from .. import engine
        r = getattr(self, 'split_row', None)
        if r is not None and self.rowsplit:
            # a row range of a CSR is a CSR over the same cols / vals arrays: rowptr keeps its absolute offsets
            for a, b in ((0, r), (r, self.shape[0])):
                engine.spmm_csr(self.rowptr[a:b + 1], self.cols, self.vals, X, out[a:b],
                                acc=None if acc is None else acc[a:b], acc_scale=acc_scale, rowsplit=True)
            return out
        return engine.spmm_csr(self.rowptr, self.cols, self.vals, X, out, acc=acc, acc_scale=acc_scale,
                               rowsplit=self.rowsplit)


class GraphRecommender(DeepRecommender):
    def __init__(self, conf, trainingSet, testSet, fold='[1]'):
        super(GraphRecommender, self).__init__(conf, trainingSet, testSet, fold)

    def create_joint_sparse_adjaceny(self):
        n_nodes = self.num_users + self.num_items
        u, i, _ = self.data.training_ids()
        ones = np.ones(u.shape[0], dtype=np.float32)
        upper = sp.csr_matrix((ones, (u, i.astype(np.int64) + self.num_users)), shape=(n_nodes, n_nodes))
        adj = upper + upper.T
        deg = np.asarray(adj.sum(1)).ravel()
        with np.errstate(divide='ignore'):
            d_inv_sqrt = np.power(deg, -0.5)
        d_inv_sqrt[np.isinf(d_inv_sqrt)] = 0.
        scale = sp.diags(d_inv_sqrt)
        return scale.dot(adj).dot(scale)

    def create_joint_sparse_adj_tensor(self):
        """The same matrix as create_joint_sparse_adjaceny(), assembled on the device from the
        id-mapped training pairs (sort + run-length; duplicates summed like scipy's constructor)."""
        import torch
        from ..graph_build import norm_adjacency_csr
        dev = self._device()
        u, i, _ = self.data.training_ids()
        rowptr, cols, vals = norm_adjacency_csr(torch.from_numpy(u), torch.from_numpy(i), self.num_users,
                                                self.num_items, device=dev)
        n = self.num_users + self.num_items
        return DeviceCSR.from_tensors((n, n), rowptr, cols, vals, split_row=self.num_users)

    def create_sparse_rating_matrix(self):
        """(U x I) COO float32, entry = 1/|items rated by the user| (graphRecommender.py:41-51)."""
        u, i, _ = self.data.training_ids()
        per_user = np.array([len(self.data.trainSet_u[self.data.id2user[k]]) for k in range(self.num_users)],
                            dtype=np.float64)
        vals = 1.0 / per_user[u]
        return sp.coo_matrix((vals, (u, i)), shape=(self.num_users, self.num_items), dtype=np.float32)

    def create_sparse_adj_tensor(self):
        return DeviceCSR(self.create_sparse_rating_matrix(), self._device())
